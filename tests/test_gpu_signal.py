"""impulse_b200.signal on the GPU: the reference's golden vectors for upfirdn / resample
(tests/test_signal.nim:67-147) through the product engine, and fftconvolve against direct convolution."""
import numpy as np
import pytest

from tests import signal_cases as cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from impulse_b200 import signal
    return signal


def mae(a, b):
    a, b = np.asarray(a, dtype=np.complex128), np.asarray(b, dtype=np.complex128)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.mean(np.abs(a - b)))


def test_upfirdn_golden(sg):
    for args, kw, want in cases.UPFIRDN:
        got = sg.upfirdn(np.array(args[0]), np.array(args[1]), **kw)
        assert mae(got, want) < 1e-12, (args, kw)
        assert got.dtype.kind == np.array(args[0]).dtype.kind


def test_resample_golden(sg):
    import torch
    for kw, want in cases.RESAMPLE:
        assert mae(sg.resample(cases.T5, **kw), want) < 1e-8, kw
    up, down = sg.reduce_resampling_rates(6, 4)
    h = sg.generate_resampling_filter(up, down)
    assert mae(sg.resample(cases.T5, h, up=up, down=down), cases.RESAMPLE[2][1]) < 1e-8
    tc = cases.T5 + 2.0j * cases.T5
    assert mae(sg.resample(tc, up=3, down=2), np.array(cases.RESAMPLE[2][1]) + 1j * np.array(cases.RESAMPLE_COMPLEX_IM)) < 1e-8
    t = np.arange(-50.0, 100.0)
    for n in (100, 107, 119):
        for up in (2, 3, 7, 9):
            for down in (2, 5, 9):
                assert len(sg.resample(t[:n], up=up, down=down)) == cases.expected_resample_len(n, up, down)
    td = torch.from_numpy(cases.T5).cuda()                 # device tensors stay on the device
    out = sg.resample(td, up=3, down=2)
    assert out.is_cuda and mae(out.cpu().numpy(), cases.RESAMPLE[2][1]) < 1e-8


def test_fftconvolve_matches_direct(sg):
    import torch
    import impulse_b200 as ib
    rng = np.random.default_rng(41)
    before = ib.launch_count()
    for dt, tol in ((np.float64, 1e-13), (np.float32, 2e-5), (np.complex128, 1e-13), (np.complex64, 2e-5)):
        for (b, n, m) in ((1, 1, 1), (3, 50, 7), (4, 1000, 129), (2, 4099, 31), (8, 16384, 257), (2, 7, 12)):
            def draw(shape):
                v = rng.standard_normal(shape)
                if np.issubdtype(dt, np.complexfloating):
                    v = v + 1j * rng.standard_normal(shape)
                return v.astype(dt)
            x, h = draw((b, n)), draw((m,))
            want = np.stack([np.convolve(r.astype(np.complex128), h.astype(np.complex128)) for r in x])
            got = sg.fftconvolve(x, h)
            assert got.dtype == dt and got.shape == (b, n + m - 1)
            scale = np.abs(want).max() + 1e-30
            assert np.abs(got - want).max() / scale < tol * max(1.0, np.log2(n + m)), (dt, b, n, m)
            if m <= n:
                for mode in ("same", "valid"):
                    g2 = sg.fftconvolve(torch.from_numpy(x).cuda(), torch.from_numpy(h).cuda(), mode)
                    w2 = np.stack([np.convolve(r.astype(np.complex128), h.astype(np.complex128), mode) for r in x])
                    assert g2.is_cuda and np.abs(g2.cpu().numpy() - w2).max() / scale < tol * max(1.0, np.log2(n + m))
    assert ib.launch_count() > before          # the library's kernels did the work
    # rank-1 signal, linearity and commutativity as size-independent properties at a long length
    x = rng.standard_normal(1 << 20)
    h = rng.standard_normal(1001)
    y = sg.fftconvolve(x, h)
    y2 = sg.fftconvolve(2.0 * x, h)
    assert y.shape == (x.size + h.size - 1,) and np.abs(y2 - 2.0 * y).max() < 1e-9
    idx = rng.integers(0, y.size, 16)
    for i in idx:                               # spot-check against the direct sum
        lo, hi = max(0, i - h.size + 1), min(x.size - 1, i)
        direct = float(np.dot(x[lo:hi + 1], h[i - np.arange(lo, hi + 1)]))
        assert abs(y[i] - direct) < 1e-9
