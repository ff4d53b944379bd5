"""Generate tests/golden/pocketfft_golden.npz from the COMPILED REFERENCE (oracle/_ref).

Run in the build container (needs /root/reference for `make -C oracle`):
    python tests/golden/make_golden.py
The .npz holds seeded inputs and the reference's outputs for small cases covering each
entry point of the path; tests compare the port, and the CUDA engine, against it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


def main():
    oracle.build()
    ref = oracle.Ref(os.path.join(ROOT, "oracle", "_ref", "libpocketfft_ref.so"))
    rng = np.random.default_rng(1234)
    g = {}

    def u(*shape):
        return rng.uniform(-0.5, 0.5, size=shape)

    # C engine, 1-D: complex and real (packed), fftpack and Bluestein lengths
    for n in (1, 2, 3, 4, 5, 7, 8, 12, 16, 27, 30, 64, 89, 100, 121, 128, 191, 243, 250, 256, 1000, 1024):
        x = (u(3, n) + 1j * u(3, n)).astype(np.complex128)
        g[f"c2c_f64_in_{n}"] = x
        g[f"c2c_f64_fwd_{n}"] = ref.cfft_rows(x.copy(), True, 1.0)
        g[f"c2c_f64_bwd_{n}"] = ref.cfft_rows(x.copy(), False, 1.0 / n)
        r = u(3, n)
        g[f"r_f64_in_{n}"] = r
        p = ref.rfft_rows(r.copy(), True, 1.0)
        g[f"r_f64_packed_fwd_{n}"] = p
        g[f"r_f64_packed_bwd_{n}"] = ref.rfft_rows(p.copy(), False, 1.0 / n)
    # C++ engine: N-D, strided, float32, r2c / c2r
    a = (u(6, 10) + 1j * u(6, 10)).astype(np.complex128)
    g["nd_c2c_f64_in"] = a
    g["nd_c2c_f64_ax01"] = ref.c2c(a, [0, 1], True, 1.0)
    g["nd_c2c_f64_ax0_bwd"] = ref.c2c(a, [0], False, 0.25)
    b = (u(4, 6, 9) + 1j * u(4, 6, 9)).astype(np.complex64)
    g["nd_c2c_f32_in"] = b
    g["nd_c2c_f32_ax12"] = ref.c2c(b, [1, 2], True, 1.0)
    r = u(8, 12).astype(np.float32)
    g["nd_r2c_f32_in"] = r
    g["nd_r2c_f32_ax01"] = ref.r2c(r, [0, 1], True, 1.0)
    g["nd_r2c_f32_ax1_bwd"] = ref.r2c(r, [1], False, 1.0)
    r64 = u(5, 9)
    g["nd_r2c_f64_in"] = r64
    s = ref.r2c(r64, [0, 1], True, 1.0)
    g["nd_r2c_f64_ax01"] = s
    g["nd_c2r_f64_ax01"] = ref.c2r(s, (5, 9), [0, 1], False, 1.0 / 45)
    s1 = ref.r2c(r64, [0], True, 1.0)
    g["nd_r2c_f64_ax0"] = s1
    g["nd_c2r_f64_ax0"] = ref.c2r(s1, (5, 9), [0], False, 0.2)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pocketfft_golden.npz"), **g)
    print("wrote", len(g), "arrays")


if __name__ == "__main__":
    main()
