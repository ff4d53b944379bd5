"""Host logic of impulse_b200.signal on CPU: filter design against the reference's golden vectors, and the
upfirdn / resample bookkeeping with a direct-convolution engine standing in for the GPU (the product engine
is exercised by tests/test_gpu_signal.py)."""
import numpy as np
import pytest

from impulse_b200 import signal as sg
from tests import signal_cases as cases


class DirectEngine:
    """np.convolve per row: the semantics of arraymancer's convolve(mode = full)."""
    host = True

    def full(self, x, h):
        import torch
        xa, ha = x.numpy(), h.numpy()
        return torch.from_numpy(np.stack([np.convolve(r, ha) for r in xa]))


def mae(a, b):
    a, b = np.asarray(a, dtype=np.complex128), np.asarray(b, dtype=np.complex128)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.mean(np.abs(a - b)))


def test_kaiser_golden():  # tests/test_signal.nim:5-16
    assert mae(sg.kaiser(4), [0.03671089, 0.7753221, 0.7753221, 0.03671089]) < 1e-8
    assert mae(sg.kaiser(4, 8.0), [0.00233883, 0.65247867, 0.65247867, 0.00233883]) < 1e-8
    assert sg.kaiser(1)[0] == 1.0 and sg.kaiser(0).size == 0


def test_firls_golden():  # tests/test_signal.nim:18-65
    e = [0.1264964, 0.278552488, 0.34506779765, 0.278552488, 0.1264964]
    assert mae(sg.firls(4, [[0.0, 0.3], [0.4, 1.0]], [[1.0, 1.0], [0.0, 0.0]]), e) < 1e-8
    assert mae(sg.firls(4, [0.0, 0.3, 0.4, 1.0], [1.0, 1.0, 0.0, 0.0]), e) < 1e-8
    assert mae(sg.firls(4, [[0.0, 1.5], [2.0, 5.0]], [[1.0, 1.0], [0.0, 0.0]], fs=10.0), e) < 1e-8
    assert mae(sg.firls(4, [[0.0, 0.3], [0.4, 1.0]], [[1.0, 1.0], [0.0, 0.0]], weights=[1.0, 0.5]),
               [0.110353444, 0.28447325, 0.36086805, 0.28447325, 0.110353444]) < 1e-8
    assert mae(sg.firls(5, [[0.0, 0.3], [0.3, 0.6], [0.6, 1.0]], [[1.0, 1.0], [1.0, 0.2], [0.0, 0.0]]),
               [-0.05603328, 0.146107441, 0.43071645, 0.43071645, 0.146107441, -0.05603328]) < 1e-8
    assert mae(sg.firls(6, [[0.0, 0.4], [0.6, 1.0]], [[0.0, 0.0], [0.9, 1.0]], symmetric=False),
               [-0.13944975, 0.2851858, -0.25859575, 0.0, 0.25859575, -0.2851858, 0.13944975]) < 1e-8
    with pytest.raises(ValueError):
        sg.firls(5, [0.0, 0.5, 0.6, 1.0], [1.0, 1.0, 0.0, 1.0])
    with pytest.raises(ValueError):
        sg.firls(4, [0.0, 0.5, 0.4, 1.0], [1.0, 1.0, 0.0, 0.0])


def test_reduce_rates_and_fast_len():
    assert sg.reduce_resampling_rates(6, 4) == (3, 2)
    for n in (1, 2, 7, 17, 1000, 1025, 4099, 8191, 100003):
        p = sg.next_fast_len(n)
        q = p
        for f in (2, 3, 5):
            while q % f == 0:
                q //= f
        assert p >= n and q == 1 and p < 2 * n


@pytest.mark.parametrize("args,kw,want", cases.UPFIRDN)
def test_upfirdn_golden_host_logic(args, kw, want):
    got = sg.upfirdn(np.array(args[0]), np.array(args[1]), engine=DirectEngine(), **kw)
    assert mae(got, want) < 1e-12
    assert got.dtype.kind == np.array(args[0]).dtype.kind      # integer signals stay integer


@pytest.mark.parametrize("kw,want", cases.RESAMPLE)
def test_resample_golden_host_logic(kw, want):
    assert mae(sg.resample(cases.T5, engine=DirectEngine(), **kw), want) < 1e-8


def test_resample_variants_host_logic():
    eng = DirectEngine()
    up, down = sg.reduce_resampling_rates(6, 4)
    h = sg.generate_resampling_filter(up, down)
    assert mae(sg.resample(cases.T5, h, up=up, down=down, engine=eng), cases.RESAMPLE[2][1]) < 1e-8
    tc = cases.T5 + 2.0j * cases.T5
    got = sg.resample(tc, up=3, down=2, engine=eng)
    assert mae(got, np.array(cases.RESAMPLE[2][1]) + 1j * np.array(cases.RESAMPLE_COMPLEX_IM)) < 1e-8
    assert np.array_equal(sg.resample(cases.T5, up=4, down=4, engine=eng), cases.T5)
    t = np.arange(-50.0, 100.0)                           # tests/test_signal.nim:150-250 (lengths table)
    for n in range(100, 120, 3):
        for up in range(2, 10):
            for down in range(2, 10):
                assert len(sg.resample(t[:n], up=up, down=down, engine=eng)) == cases.expected_resample_len(n, up, down)


def test_fftconvolve_modes_host_logic():
    rng = np.random.default_rng(3)
    x, h = rng.standard_normal((3, 50)), rng.standard_normal(7)
    eng = DirectEngine()
    for mode in ("full", "same", "valid"):
        got = sg.fftconvolve(x, h, mode, engine=eng)
        want = np.stack([np.convolve(r, h, mode) for r in x])
        assert mae(got, want) < 1e-13
    assert mae(sg.fftconvolve(x[0], h, engine=eng), np.convolve(x[0], h)) < 1e-13
    with pytest.raises(ValueError):
        sg.fftconvolve(x, h, "circular", engine=eng)


def test_product_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from impulse_b200 import FFTError
    with pytest.raises(FFTError):
        sg.fftconvolve(np.ones(8), np.ones(3))
