"""CPU-only check of the fused Bluestein kernel through the thread-level host emulation (tests/emu/emu_fastblue.cpp):
the kernel body the GPU runs plus the planner's own tables.  All variants here are parity-green on the B200 as well;
the emulation exists so that the next restructuring of this kernel can be checked before GPU time is spent."""
import numpy as np
import pytest

from tests.emu import harness_fastblue as fb


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


VARIANTS = [dict(), dict(bk_smem=True), dict(bk_smem=True, bf_early=True),
            dict(bk_smem=True, four_pass=True), dict(bk_smem=True, bf_early=True, four_pass=True),
            dict(bk_smem=True, bf_early=True, four_pass=True, bf_tmem=True)]   # FFT(b)/M in tensor memory


@pytest.mark.parametrize("var", VARIANTS)
def test_config3c_length_real_pairs(var):
    L = 4099                                   # BASELINE config 3c: prime, work length 8192 with 5 aliased lags
    rng = np.random.default_rng(1)
    x = rng.random((5, L)) - 0.5               # odd row count: the last unit holds a single row
    for fwd in (True, False):
        got = fb.run("r2c", x, L, fwd, 0.5, **var)
        want = np.fft.rfft(x, axis=1) * 0.5
        assert rel(got, want if fwd else np.conj(want)) < 1e-14
    X = np.fft.rfft(x, axis=1)
    X[:, 0] += 0.125j                          # ignored by pocketfft's c2r
    back = fb.run("c2r", X, L, False, 1.0 / L, **var)
    assert rel(back, x) < 1e-14
    back = fb.run("c2r", np.conj(X), L, True, 1.0 / L, **var)
    assert rel(back, x) < 1e-14


@pytest.mark.parametrize("var", VARIANTS)
def test_complex_lines(var):
    rng = np.random.default_rng(2)
    for L in (4099, 4001, 4100):               # d = 5, exact fit, d = 7
        z = rng.random((3, L)) - 0.5 + 1j * (rng.random((3, L)) - 0.5)
        assert rel(fb.run("c2c", z, L, True, 1.0, **var), np.fft.fft(z, axis=1)) < 1e-14
        assert rel(fb.run("c2c", z, L, False, 1.0 / L, **var), np.fft.ifft(z, axis=1)) < 1e-14


def test_shorter_work_lengths():
    rng = np.random.default_rng(3)
    for L in (1021, 2053 - 4, 1031, 521):       # work lengths 2048 / 4096 / 4096 / 2048
        x = rng.random((4, L)) - 0.5
        assert rel(fb.run("r2c", x, L, True, 1.0), np.fft.rfft(x, axis=1)) < 1e-14
        z = rng.random((2, L)) - 0.5 + 1j * (rng.random((2, L)) - 0.5)
        assert rel(fb.run("c2c", z, L, True, 1.0), np.fft.fft(z, axis=1)) < 1e-14


def test_c2c_8192_on_the_four_pass_core():
    rng = np.random.default_rng(4)
    z = rng.random((5, 8192)) - 0.5 + 1j * (rng.random((5, 8192)) - 0.5)   # 5 rows over 2 CTAs
    assert rel(fb.run_fast4(z, True, 0.7), np.fft.fft(z, axis=1) * 0.7) < 2e-15 * 13
    assert rel(fb.run_fast4(z, False, 1.0 / 8192), np.fft.ifft(z, axis=1)) < 2e-15 * 13


def test_float32():
    rng = np.random.default_rng(6)
    for L, var in ((4099, dict(bk_smem=True, bf_early=True)), (4099, dict()), (2051, dict()), (1021, dict())):
        x = (rng.random((5, L)) - 0.5).astype(np.float32)
        got = fb.run("r2c", x, L, True, 1.0, **var)
        assert got.dtype == np.complex64
        want = np.fft.rfft(x.astype(np.float64), axis=1)
        assert rel(got.astype(np.complex128), want) < 1e-5 * np.log2(L) / 10
        back = fb.run("c2r", want.astype(np.complex64), L, False, 1.0 / L, **var)
        assert rel(back.astype(np.float64), x.astype(np.float64)) < 1e-5 * np.log2(L) / 10
        z = ((rng.random((3, L)) - 0.5) + 1j * (rng.random((3, L)) - 0.5)).astype(np.complex64)
        assert rel(fb.run("c2c", z, L, True, 1.0, **var).astype(np.complex128), np.fft.fft(z.astype(np.complex128), axis=1)) < 1e-5 * np.log2(L) / 10
