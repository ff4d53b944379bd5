"""Small cases of every kernel family for compute-sanitizer (memcheck / racecheck) runs:
    compute-sanitizer --tool memcheck python tests/sanitizer_cases.py
Not collected by pytest (no test_ prefix); exits non-zero on a parity failure."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
import impulse_b200 as ib  # noqa: E402
from impulse_b200.filter import FFTFilter2D  # noqa: E402


def run(kind, x, out, axes, fwd=True, fct=1.0):
    ib.FFTDesc.init(axes=axes, forward=fwd, scalingFactor=fct).apply(ib.DataDesc.init(out), ib.DataDesc.init(x))
    return out


def new_kernels_only():
    """`--new`: only the kernel variants added last (paired real rows, register prefetch, chirp table in shared
    memory), so that a memcheck / racecheck pass fits in a couple of minutes."""
    rng = np.random.default_rng(0)
    ok = True
    for dt, cdt, tol in ((np.float64, torch.complex128, 1e-9), (np.float32, torch.complex64, 2e-4)):
        for n in (512, 1000, 2048, 3888, 4096):
            x = torch.from_numpy(rng.standard_normal((9, n)).astype(dt)).cuda()
            s = run("r2c", x, torch.empty((9, n // 2 + 1), dtype=cdt, device="cuda"), [1])
            ok &= bool(torch.allclose(s, torch.fft.rfft(x, dim=1), rtol=tol, atol=tol * n))
            ok &= bool(torch.allclose(run("c2r", s, torch.empty_like(x), [1], False, 1.0 / n), x, atol=tol * 10))
    z = torch.from_numpy(rng.standard_normal((11, 2048)) + 1j * rng.standard_normal((11, 2048))).cuda()
    ok &= bool(torch.allclose(run("c2c", z, torch.empty_like(z), [1]), torch.fft.fft(z, dim=1), rtol=1e-9, atol=1e-8))
    x = torch.from_numpy(rng.standard_normal((5, 4099))).cuda()
    s = run("r2c", x, torch.empty((5, 2050), dtype=torch.complex128, device="cuda"), [1])
    ok &= bool(torch.allclose(s, torch.fft.rfft(x, dim=1), rtol=1e-9, atol=1e-8))
    ok &= bool(torch.allclose(run("c2r", s, torch.empty_like(x), [1], False, 1.0 / 4099), x, atol=1e-9))
    z = torch.from_numpy(rng.standard_normal((3, 4099)) + 1j * rng.standard_normal((3, 4099))).cuda()
    ok &= bool(torch.allclose(run("c2c", z, torch.empty_like(z), [1]), torch.fft.fft(z, dim=1), rtol=1e-9, atol=1e-8))
    # ---- round 2: new two- and three-pass shapes (both precisions), wide 128 / 256-point rows, float32 fused Bluestein
    kernels = set()
    for n in (100, 128, 243, 256, 625, 1536, 2000, 2187, 3000, 4000, 6561):
        for cdt, tol in ((np.complex128, 1e-9), (np.complex64, 2e-4)):
            z = torch.from_numpy((rng.standard_normal((7, n)) + 1j * rng.standard_normal((7, n))).astype(cdt)).cuda()
            ok &= bool(torch.allclose(run("c2c", z, torch.empty_like(z), [1]), torch.fft.fft(z, dim=1), rtol=tol, atol=tol * n))
            ok &= bool(torch.allclose(run("c2c", z, torch.empty_like(z), [1], False, 1.0 / n), torch.fft.ifft(z, dim=1), rtol=tol, atol=tol))
            kernels.add(ib.last_kernel())
    z32 = torch.from_numpy((rng.standard_normal((4, 4099)) + 1j * rng.standard_normal((4, 4099))).astype(np.complex64)).cuda()
    ok &= bool(torch.allclose(run("c2c", z32, torch.empty_like(z32), [1]), torch.fft.fft(z32, dim=1), rtol=2e-3, atol=0.2))
    kernels.add(ib.last_kernel())
    # fused column transform (both launches of the split in one kernel, intermediate in the L2 ring): 64 x 128 and 128 x 128
    for shape, cdt, tol in (((8192, 80), np.complex128, 1e-9), ((3, 16384, 48), np.complex64, 2e-3)):
        z = torch.from_numpy((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cdt)).cuda()
        ax = len(shape) - 2
        ok &= bool(torch.allclose(run("c2c", z, torch.empty_like(z), [ax]), torch.fft.fft(z, dim=ax), rtol=tol, atol=tol * 1e3))
        kernels.add(ib.last_kernel())
    # long even real lines at any stride (gather / scatter passes), DCT of a length beyond one CTA
    y = torch.from_numpy(rng.standard_normal((40000, 3))).cuda()
    s = run("r2c", y, torch.empty((20001, 3), dtype=torch.complex128, device="cuda"), [0])
    ok &= bool(torch.allclose(s, torch.fft.rfft(y, dim=0), rtol=1e-9, atol=1e-6))
    ok &= bool(torch.allclose(run("c2r", s, torch.empty_like(y), [0], False, 1.0 / 40000), y, atol=1e-9))
    d = torch.from_numpy(rng.standard_normal((2, 10000))).cuda()
    o = torch.empty_like(d)
    ib.DCTDesc.init(axes=[1], dctType=2).apply(ib.DataDesc.init(o), ib.DataDesc.init(d))
    back = torch.empty_like(d)
    ib.DCTDesc.init(axes=[1], dctType=3, scalingFactor=1.0 / 20000).apply(ib.DataDesc.init(back), ib.DataDesc.init(o))
    ok &= bool(torch.allclose(back, d, atol=1e-9))
    torch.cuda.synchronize()
    print("sanitizer cases (new kernels):", "ok" if ok else "PARITY FAILURE", "kernels", sorted(kernels))
    sys.exit(0 if ok else 1)


def whole_axis_only():
    """`--whole`: every instance of the whole-axis kernel (colconvw_kernel): convolution and plain transform, vector and
    per-line global access, ragged groups, several tiles per CTA — the in-place exchange scheme under racecheck."""
    import ctypes as C
    from impulse_b200 import _lib
    os.environ["IMPULSE_FFT_CONV_WHOLE"] = "2"
    os.environ["IMPULSE_FFT_COL_WHOLE"] = "2"
    L = _lib.lib()
    rng = np.random.default_rng(3)
    ok = True
    kernels = set()
    for n in (512, 1024, 2048, 4096):
        for cdt, tol in ((np.complex64, 2e-3), (np.complex128, 1e-9)):
            for cols in (10, 37):     # even: vector access, ragged last group; odd: per-line access
                shape = (320 if n <= 1024 else 40, n, cols)   # more tiles than CTAs at the small sizes
                z = torch.from_numpy((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cdt)).cuda()
                m = torch.from_numpy((rng.standard_normal(n * cols) + 1j * rng.standard_normal(n * cols)).astype(cdt)).cuda()
                want = torch.fft.ifft(torch.fft.fft(z, dim=1) * m.view(n, cols), dim=1)
                esz = z.element_size()
                st = (C.c_ssize_t * 3)(n * cols * esz, cols * esz, esz)
                _lib.check(L.impulse_fft_convolve_axis(_lib.F32 if cdt == np.complex64 else _lib.F64, 3, (C.c_size_t * 3)(*shape), st, st, 1,
                                                       z.data_ptr(), z.data_ptr(), 1.0 / n, m.data_ptr(), n * cols, None))
                kernels.add(ib.last_kernel())
                ok &= bool(torch.allclose(z, want, rtol=tol, atol=tol * 10))
    for n in (1024, 2048):
        for cdt, tol in ((np.complex64, 2e-3), (np.complex128, 1e-9)):
            for cols in (12, 21):
                shape = (200 if n == 1024 else 30, n, cols)
                z = torch.from_numpy((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(cdt)).cuda()
                for fwd in (True, False):
                    got = run("c2c", z, torch.empty_like(z), [1], fwd, 1.0)
                    kernels.add(ib.last_kernel())
                    ref = torch.fft.fft(z, dim=1) if fwd else torch.fft.ifft(z, dim=1) * n
                    ok &= bool(torch.allclose(got, ref, rtol=tol, atol=tol * n))
    torch.cuda.synchronize()
    ok &= all(k.startswith("colconvw_kernel") for k in kernels)
    print("sanitizer cases (whole-axis kernel):", "ok" if ok else "PARITY FAILURE", sorted(kernels))
    sys.exit(0 if ok else 1)


def main():
    if "--whole" in sys.argv:
        whole_axis_only()
    if "--new" in sys.argv:
        new_kernels_only()
    rng = np.random.default_rng(0)
    ok = True

    def c(shape, dt=np.complex128):
        return torch.from_numpy((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)).cuda()

    def r(shape, dt=np.float64):
        return torch.from_numpy(rng.standard_normal(shape).astype(dt)).cuda()

    # generic engine (mixed radix, Bluestein, r2c/c2r odd+even, packed), fast2/fast2p, fast3, colfast2, fastblue, four-step, DCT
    for n in (1, 6, 16, 30, 32, 64, 97, 128, 1000, 256, 512, 1024, 2048, 4096, 8192, 4099, 500, 1944):
        for rows in (1, 5):
            x = c((rows, n))
            y = run("c2c", x, torch.empty_like(x), [1])
            back = run("c2c", y, torch.empty_like(x), [1], False, 1.0 / n)
            ok &= bool(torch.allclose(back, x, atol=1e-9))
            ref = torch.fft.fft(x, dim=1)   # second opinion only; parity proper is in tests/
            ok &= bool(torch.allclose(y, ref, rtol=1e-9, atol=1e-9 * n))
    for n in (5, 8, 32, 64, 128, 256, 512, 1000, 1024, 2048, 3888, 4096, 4099, 16384, 32768, 29999):
        x = r((3, n))
        s = run("r2c", x, torch.empty((3, n // 2 + 1), dtype=torch.complex128, device="cuda"), [1])
        ok &= bool(torch.allclose(s, torch.fft.rfft(x, dim=1), rtol=1e-9, atol=1e-9 * n))
        b = run("c2r", s, torch.empty_like(x), [1], False, 1.0 / n)
        ok &= bool(torch.allclose(b, x, atol=1e-9))
        p = ib.rfft_packed(x)
        ok &= bool(torch.allclose(ib.rfft_packed(p, forward=False), x, atol=1e-9))
    for shape, axes in (((64, 24), [0]), ((4096, 16), [0]), ((300, 520), [0, 1]), ((3, 40, 36), [1, 2])):
        x = c(shape)
        y = run("c2c", x, torch.empty_like(x), axes)
        ok &= bool(torch.allclose(y, torch.fft.fftn(x, dim=axes), rtol=1e-9, atol=1e-8 * max(shape)))
    x = c((2, 32768))
    ok &= bool(torch.allclose(run("c2c", x, torch.empty_like(x), [1]), torch.fft.fft(x, dim=1), rtol=1e-9, atol=1e-6))
    for n in (16, 64, 256, 512, 1024, 8192, 16384):                   # fp32 row kernels and the fp32 split
        x32 = c((7, n), np.complex64)
        ok &= bool(torch.allclose(run("c2c", x32, torch.empty_like(x32), [1]), torch.fft.fft(x32, dim=1), rtol=1e-4, atol=2e-2))
    r32 = r((5, 2048), np.float32)                                    # real rows on the three-pass kernels, fp32
    ok &= bool(torch.allclose(run("r2c", r32, torch.empty((5, 1025), dtype=torch.complex64, device="cuda"), [1]),
                              torch.fft.rfft(r32, dim=1), rtol=1e-4, atol=2e-2))
    for n in (1000, 3888, 4096):                                      # paired real kernels (r2c in pass 3, c2r in pass 1), fp32
        r32 = r((7, n), np.float32)
        s32 = run("r2c", r32, torch.empty((7, n // 2 + 1), dtype=torch.complex64, device="cuda"), [1])
        ok &= bool(torch.allclose(s32, torch.fft.rfft(r32, dim=1), rtol=1e-4, atol=2e-2))
        ok &= bool(torch.allclose(run("c2r", s32, torch.empty_like(r32), [1], False, 1.0 / n), r32, atol=1e-4))
    x = c((1, 1 << 19))                                                # peeled three-pass split, pipelined column kernel
    ok &= bool(torch.allclose(run("c2c", x, torch.empty_like(x), [1]), torch.fft.fft(x, dim=1), rtol=1e-9, atol=1e-5))
    x = c((8192, 160))                                                 # strided 8192-point columns: 64 x 128, 16 lines per CTA
    ok &= bool(torch.allclose(run("c2c", x, torch.empty_like(x), [0]), torch.fft.fft(x, dim=0), rtol=1e-9, atol=1e-6))
    x = c((2, 20011))                                                  # long Bluestein line: elementwise chirp passes
    ok &= bool(torch.allclose(run("c2c", x, torch.empty_like(x), [1]), torch.fft.fft(x, dim=1), rtol=1e-9, atol=1e-6))
    # fused multiply, axis convolution (fused middle pass), Hartley fold, FFTPACK
    import ctypes as C
    from impulse_b200 import _lib
    L = _lib.lib()
    z = c((2, 4096, 24), np.complex64)
    m = c((4096 * 24,), np.complex64)
    want = torch.fft.ifft(torch.fft.fft(z, dim=1) * m.view(4096, 24), dim=1)
    st = (C.c_ssize_t * 3)(4096 * 24 * 8, 24 * 8, 8)
    _lib.check(L.impulse_fft_convolve_axis(_lib.F32, 3, (C.c_size_t * 3)(2, 4096, 24), st, st, 1, z.data_ptr(), z.data_ptr(),
                                           1.0 / 4096, m.data_ptr(), 4096 * 24, None))
    ok &= bool(torch.allclose(z, want, rtol=1e-3, atol=1e-3))
    hsrc = r((6, 10, 12))
    hout = torch.empty_like(hsrc)
    ib.r2r_genuine_hartley(ib.DataDesc.init(hout), ib.DataDesc.init(hsrc), [0, 2])
    f = torch.fft.fftn(hsrc, dim=[0, 2])
    ok &= bool(torch.allclose(hout, f.real + f.imag, atol=1e-9))
    ib.r2r_fftpack(ib.DataDesc.init(hout), ib.DataDesc.init(hsrc), [1, 2], True, True)
    from impulse_b200 import signal as isig
    sig = isig.fftconvolve(r((3, 5000)), r((65,)))
    ok &= sig.shape == (3, 5064)
    img = torch.rand((2, 64, 96), device="cuda", dtype=torch.float32)
    ker = torch.rand((5, 5), device="cuda", dtype=torch.float32)
    FFTFilter2D(ker / ker.sum(), 64, 96).apply(img)
    d = r((4, 100))
    o = torch.empty_like(d)
    ib.DCTDesc.init(axes=[1], dctType=2).apply(ib.DataDesc.init(o), ib.DataDesc.init(d))
    h = np.random.default_rng(1).standard_normal((600, 1024)) + 0j     # host staging path
    ok &= bool(np.allclose(ib.fft(h), np.fft.fft(h, axis=1), rtol=1e-9, atol=1e-8))
    torch.cuda.synchronize()
    print("sanitizer cases:", "ok" if ok else "PARITY FAILURE", "last kernel", ib.last_kernel())
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
