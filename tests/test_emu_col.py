"""CPU-only check of the column register kernels (colfast2 / colpipe2 / colconv2) through their thread-level host
emulation (tests/emu/emu_col.cpp), driven by the jobs the product's planner builds: strided axes of N-D transforms,
both launches of the four-step split, the fused convolution middle pass.  All of these are parity-green on the B200;
the emulation exists so that the next change to these kernels can be checked before GPU time is spent."""
import numpy as np
import pytest

from tests.emu import harness as emu

COL_IDS = set(range(13, 23))         # COL2_* (fft_types.h)
CONV_IDS = set(range(26, 32))        # COLCONV_*


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


@pytest.fixture(autouse=True)
def _fast_cols_off_afterwards():
    yield
    emu.set_fast_cols(None)


def rnd(rng, shape, dt):
    return ((rng.random(shape) - 0.5) + 1j * (rng.random(shape) - 0.5)).astype(dt)


@pytest.mark.parametrize("groups", [2, 0])          # 2 = colpipe2 where it applies (product default), 0 = plain colfast2
@pytest.mark.parametrize("dt,shape,axis", [(np.complex128, (128, 64), 0), (np.complex128, (64, 40), 0), (np.complex64, (256, 48), 0),
                                            (np.complex128, (3, 64, 32), 1), (np.complex64, (2, 32, 16), 1), (np.complex128, (512, 8), 0)])
def test_strided_axis(groups, dt, shape, axis):
    rng = np.random.default_rng(sum(shape))
    x = rnd(rng, shape, dt)
    ids = emu.nd_fast_ids("c2c", x, x, shape, [axis], True)
    assert len(ids) == 1 and ids[0] in COL_IDS, ids
    emu.set_fast_cols(groups)
    tol = 2e-15 * 12 if dt == np.complex128 else 2e-6
    for fwd in (True, False):
        got = emu.nd("c2c", x, np.empty_like(x), shape, [axis], fwd, 0.5)
        want = (np.fft.fft(x.astype(np.complex128), axis=axis) if fwd else np.fft.ifft(x.astype(np.complex128), axis=axis) * shape[axis]) * 0.5
        assert rel(got, want) < tol, (fwd, rel(got, want))
    assert emu.col_job_count() == 2


@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
@pytest.mark.parametrize("n,rows", [(16384, 1), (16384, 3), (32768, 2)])
def test_four_step_split_on_the_column_kernels(dt, n, rows):
    rng = np.random.default_rng(n + rows)
    x = rnd(rng, (rows, n), dt)
    ids = emu.nd_fast_ids("c2c", x, x, x.shape, [1], True)
    assert len(ids) == 2 and all(i in COL_IDS for i in ids), ids      # A: adjacent lines + W_N^(k n2), B: rows in, lines out
    emu.set_fast_cols(2)
    got = emu.nd("c2c", x, np.empty_like(x), x.shape, [1], True, 1.0)
    assert rel(got, np.fft.fft(x.astype(np.complex128), axis=1)) < (2e-15 * 16 if dt == np.complex128 else 3e-6)
    assert emu.col_job_count() == 2


@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
def test_fused_convolution_middle_pass(dt, monkeypatch):
    monkeypatch.setenv("IMPULSE_FFT_CONV_WHOLE", "0")   # the three-launch scheme (axes beyond 4096 points on the GPU)
    rng = np.random.default_rng(9)
    shape = (2, 1024, 16)                       # axis of 1024 = 32 x 32: pass A, colconv2, pass B
    x = rnd(rng, shape, dt)
    m = rnd(rng, (1024, 16), dt)
    want = np.fft.ifft(np.fft.fft(x.astype(np.complex128), axis=1) * m.astype(np.complex128), axis=1)
    emu.set_fast_cols(2)
    got = emu.convolve_axis(x, np.empty_like(x), 1, m.reshape(-1), 1.0 / 1024)
    assert rel(got, want) < (1e-13 if dt == np.complex128 else 5e-6)
    assert emu.col_job_count() == 3             # pass A, the fused middle pass, pass B


# whole-axis convolution (colconvw_kernel): one launch, the axis resident in shared memory
@pytest.mark.parametrize("dt,shape", [(np.complex64, (3, 512, 19)), (np.complex64, (2, 1024, 9)), (np.complex64, (2, 2048, 10)),
                                       (np.complex64, (3, 4096, 5)), (np.complex128, (3, 512, 7)), (np.complex128, (2, 1024, 5)),
                                       (np.complex128, (2, 2048, 5)), (np.complex128, (2, 4096, 3))])
def test_whole_axis_convolution(dt, shape, monkeypatch):
    monkeypatch.setenv("IMPULSE_FFT_CONV_WHOLE", "2")   # 2048 / 4096 points too (by default they take the three-launch scheme)
    rng = np.random.default_rng(shape[1] + shape[2])
    x = rnd(rng, shape, dt)
    m = rnd(rng, shape[1:], dt)
    want = np.fft.ifft(np.fft.fft(x.astype(np.complex128), axis=1) * m.astype(np.complex128), axis=1) * 0.5
    emu.set_fast_cols(2)
    got = emu.convolve_axis(x, np.full_like(x, np.nan), 1, m.reshape(-1), 0.5 / shape[1])
    tol = 1e-13 if dt == np.complex128 else 5e-6
    assert rel(got, want) < tol, rel(got, want)
    assert emu.col_job_count() == 1             # one launch
    # in place, as FFTFilter2D calls it (every tile reads its whole lines before it writes them)
    y = x.copy()
    emu.convolve_axis(y, y, 1, m.reshape(-1), 0.5 / shape[1])
    assert rel(y, want) < tol


def test_whole_axis_convolution_padded_pitch():
    """the layout FFTFilter2D uses: half spectra in rows padded to a multiple of four bins, multiplier padded alike"""
    rng = np.random.default_rng(5)
    b, h, wc, pitch = 2, 512, 9, 12
    buf = rnd(rng, (b, h, pitch), np.complex64)
    mbuf = rnd(rng, (h, pitch), np.complex64)
    x, m = buf[:, :, :wc], mbuf[:, :wc]
    want = np.fft.ifft(np.fft.fft(x.astype(np.complex128), axis=1) * m.astype(np.complex128), axis=1)
    keep = buf.copy()
    emu.set_fast_cols(2)
    emu.convolve_axis(x, x, 1, mbuf.reshape(-1), 1.0 / h)
    assert emu.col_job_count() == 1
    assert rel(buf[:, :, :wc], want) < 5e-6
    assert np.array_equal(buf[:, :, wc:], keep[:, :, wc:])   # the padding is not touched


def test_three_launch_convolution_padded_pitch(monkeypatch):
    """the same layout on the three-launch scheme (what a 4096-row image takes by default)"""
    monkeypatch.setenv("IMPULSE_FFT_CONV_WHOLE", "0")
    rng = np.random.default_rng(6)
    b, h, wc, pitch = 2, 1024, 17, 20
    buf = rnd(rng, (b, h, pitch), np.complex64)
    mbuf = rnd(rng, (h, pitch), np.complex64)
    x, m = buf[:, :, :wc], mbuf[:, :wc]
    want = np.fft.ifft(np.fft.fft(x.astype(np.complex128), axis=1) * m.astype(np.complex128), axis=1)
    keep = buf.copy()
    emu.set_fast_cols(2)
    emu.convolve_axis(x, x, 1, mbuf.reshape(-1), 1.0 / h)
    assert emu.col_job_count() == 3
    assert rel(buf[:, :, :wc], want) < 5e-6
    assert np.array_equal(buf[:, :, wc:], keep[:, :, wc:])


COLW_IDS = set(range(99, 103))       # COLW_* (fft_types.h)


# plain c2c of a strided 1024 / 2048-point axis on the whole-axis kernel: one launch instead of the two of the split
@pytest.mark.parametrize("dt,shape,axis", [(np.complex64, (1024, 21), 0), (np.complex64, (2, 2048, 9), 1), (np.complex128, (3, 1024, 10), 1),
                                            (np.complex128, (2048, 5), 0), (np.complex64, (2, 3, 1024, 18), 2)])
def test_strided_axis_whole(dt, shape, axis, monkeypatch):
    monkeypatch.setenv("IMPULSE_FFT_COL_WHOLE", "2")    # every instance (the default is complex128 at 1024 points only)
    rng = np.random.default_rng(sum(shape))
    x = rnd(rng, shape, dt)
    emu.set_fast_cols(2)
    ids = emu.nd_fast_ids("c2c", x, x, shape, [axis], True)
    assert len(ids) == 1 and ids[0] in COLW_IDS, ids
    tol = 2e-15 * 12 if dt == np.complex128 else 2e-6
    for fwd in (True, False):
        got = emu.nd("c2c", x, np.full_like(x, np.nan), shape, [axis], fwd, 0.5)
        want = (np.fft.fft(x.astype(np.complex128), axis=axis) if fwd else np.fft.ifft(x.astype(np.complex128), axis=axis) * shape[axis]) * 0.5
        assert rel(got, want) < tol, (fwd, rel(got, want))
    y = x.copy()
    emu.nd("c2c", y, y, shape, [axis], True, 1.0)      # in place
    assert rel(y, np.fft.fft(x.astype(np.complex128), axis=axis)) < tol
    assert emu.col_job_count() == 3


def test_fft2_of_1024_square_is_two_launches():
    x = np.zeros((2, 1024, 1024), np.complex128)
    emu.set_fast_cols(2)
    assert len(emu.nd_fast_ids("c2c", x, x, x.shape, [1, 2], True)) == 2
