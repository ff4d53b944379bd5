"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

numpy restatement of the Nim-side glue of the reference's C backend
(/root/reference/impulse/fft/c_pocketfft/pocketfft.nim, cited NC:line), composed on
top of a 1-D oracle engine (``oracle.Ref`` or ``oracle.Port``).  It is the checker
for impulse_b200's host mirror; it is never imported by the product.

The Nim code cannot run here (no nim toolchain); these few index shuffles are pinned
by the README vectors and tests/test_fft2.nim known answers in tests/test_oracle.py.
"""
from __future__ import annotations

import math

import numpy as np

NK_BACKWARD, NK_ORTHO, NK_FORWARD, NK_CUSTOM = "nkBackward", "nkOrtho", "nkForward", "nkCustom"


def is_odd(i: int) -> bool:  # NC:116
    return (i & 1) == 1


def unpack_fft(data: np.ndarray) -> np.ndarray:
    """NC:126-158 — packed halfcomplex floats -> N/2+1 (even) or (N+1)/2 (odd) complex."""
    n = len(data)
    out_len = (n + 1) // 2 if is_odd(n) else (n + 2) // 2
    out = np.zeros(out_len, dtype=np.complex128)
    k = 0
    re = float(data[0])
    im = 0.0
    out[k] = complex(re, im)
    k += 1
    for i in range(1, n):
        if is_odd(i):
            re = float(data[i])
        else:
            im = float(data[i])
            out[k] = complex(re, im)
            k += 1
    if not is_odd(n):
        out[k] = complex(re, 0.0)
    return out


def symm_target_size(data: np.ndarray) -> int:
    """NC:173-183 — for complex input the parity of the original length is guessed from
    whether the last bin's imaginary part is exactly 0.0 (data dependent)."""
    n = len(data)
    if np.iscomplexobj(data):
        sub = 2 if data[-1].imag == 0.0 else 1
        return n * 2 - sub
    return n


def symmetrize(data: np.ndarray) -> np.ndarray:
    """NC:160-187 — fill bins N/2+1..N-1 with the Hermitian conjugates."""
    out_len = symm_target_size(data)
    res = np.zeros(out_len, dtype=np.complex128)
    if np.iscomplexobj(data):
        res[: len(data)] = data
    else:
        u = unpack_fft(data)
        res[: len(u)] = u
    k = out_len - 1
    for i in range(1, -(-out_len // 2)):  # 1 ..< ceilDiv(outLen, 2)
        res[k] = np.conj(res[i])
        k -= 1
    return res


def init_normalize(kind: str, forward: bool, value: float, length: int) -> float:
    """NC:199-204."""
    if kind == NK_BACKWARD:
        return 1.0 if forward else 1.0 / float(length)
    if kind == NK_FORWARD:
        return 1.0 / float(length) if forward else 1.0
    if kind == NK_ORTHO:
        return 1.0 / math.sqrt(float(length))
    if kind == NK_CUSTOM:
        return value
    raise ValueError(kind)


class NimApi:
    """fft / ifft / rfft / rfft_packed as documented (normalize honoured — the Tensor
    overloads' behaviour, pocketfft_arraymancer.nim:33-94)."""

    def __init__(self, engine):
        self.e = engine

    def rfft_packed(self, data, forward=True, normalize=NK_BACKWARD, norm_value=math.inf):
        a = np.array(data, dtype=np.float64).reshape(1, -1)
        fct = init_normalize(normalize, forward, norm_value, a.shape[1])
        return self.e.rfft_rows(a, forward, fct)[0]

    def fft_inplace(self, data, forward=True, normalize=NK_BACKWARD, norm_value=math.inf):
        """fft(var data): complex -> c2c in place; float -> packed real transform (NC:290-310)."""
        fct = init_normalize(normalize, forward, norm_value, len(data))
        if np.iscomplexobj(data):
            a = np.array(data, dtype=np.complex128).reshape(1, -1)
            return self.e.cfft_rows(a, forward, fct)[0]
        a = np.array(data, dtype=np.float64).reshape(1, -1)
        return self.e.rfft_rows(a, forward, fct)[0]

    def rfft(self, data, forward=True, normalize=NK_BACKWARD, norm_value=math.inf):
        return unpack_fft(self.rfft_packed(data, forward, normalize, norm_value))  # NC:321-332

    def fft(self, data, forward=True, normalize=NK_BACKWARD, norm_value=math.inf):
        data = np.asarray(data)
        if np.iscomplexobj(data):
            return self.fft_inplace(data, forward, normalize, norm_value)
        return symmetrize(self.rfft_packed(data, forward, normalize, norm_value))  # NC:345

    def ifft(self, data, backward=True, normalize=NK_BACKWARD, norm_value=math.inf):
        return self.fft(data, not backward, normalize, norm_value)  # NC:350-360
