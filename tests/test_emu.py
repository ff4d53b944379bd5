"""CPU-only checks of the CUDA engine's logic through the host emulation (tests/emu):
the same phase functions and the same planner the kernels use, against the oracle."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests.emu import harness as emu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "pocketfft_golden.npz"))


def tol(n, dtype=np.float64):
    """north_star bound: 1e-12*log2(N) fp64, 1e-5*log2(N) fp32."""
    base = 1e-12 if dtype in (np.float64, np.complex128) else 1e-5
    return base * max(1.0, np.log2(max(n, 2)))


def rnd(rng, shape, dtype):
    if dtype in (np.complex128, np.complex64):
        return (rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)).astype(dtype)
    return rng.uniform(-0.5, 0.5, shape).astype(dtype)


def test_planner_schedules():
    assert emu.plan_info(1024) == (1024, False, [8, 8, 4, 4])
    assert emu.plan_info(4096) == (4096, False, [8, 8, 8, 8])
    assert emu.plan_info(8192) == (8192, False, [8, 8, 8, 4, 4])
    assert emu.plan_info(1000)[2] == [5, 5, 5, 8]
    assert emu.plan_info(1944)[2] == [9, 9, 3, 8]
    n_fft, blue, rad = emu.plan_info(4099)
    assert blue and n_fft >= 2 * 4099 - 1 and np.prod(rad) == n_fft
    assert emu.plan_info(1) == (1, False, [])
    assert emu.plan_info(37)[1] is True   # prime > 31 -> Bluestein
    assert emu.plan_info(31)[1] is False


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_c2c_lengths(checker, dtype):
    rng = np.random.default_rng(5)
    lengths = list(range(1, 131)) + [144, 169, 187, 191, 243, 256, 289, 343, 360, 361, 500, 512, 529, 625,
                                     729, 841, 961, 1000, 1024, 1331, 2048, 2187, 3125, 3888, 4096, 4099]
    if dtype == np.complex64:
        lengths = lengths[::3]
    for n in lengths:
        x = rnd(rng, (3, n), dtype)
        for fwd in (True, False):
            fct = 1.0 if fwd else 1.0 / n
            want = checker.c2c(x, [1], fwd, fct) if dtype == np.complex64 else checker.cfft_rows(x.copy(), fwd, fct)
            got = emu.nd("c2c", x, np.empty_like(x), x.shape, [1], fwd, fct)
            assert oracle.max_row_rel_l2(got, want) <= tol(n, dtype), (n, fwd)


def test_c2c_inplace_and_golden(checker):
    for key in GOLD.files:
        if key.startswith("c2c_f64_in_"):
            n = int(key.rsplit("_", 1)[1])
            x = GOLD[key].copy()
            emu.nd("c2c", x, x, x.shape, [1], True, 1.0)
            assert oracle.max_row_rel_l2(x, GOLD[f"c2c_f64_fwd_{n}"]) <= tol(n), n


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_r2c_c2r_hermitian(checker, dtype):
    rng = np.random.default_rng(6)
    cdt = np.complex128 if dtype == np.float64 else np.complex64
    lengths = list(range(1, 70)) + [74, 89, 97, 100, 121, 128, 191, 192, 243, 250, 382, 500, 1000, 1024, 2000, 3888, 4096, 4099, 4126]
    if dtype == np.float32:
        lengths = lengths[::2]
    for n in lengths:
        x = rnd(rng, (2, n), dtype)
        for fwd in (True, False):
            want = checker.r2c(x, [1], fwd, 0.5)
            got = emu.nd("r2c", x, np.zeros((2, n // 2 + 1), cdt), x.shape, [1], fwd, 0.5)
            assert oracle.max_row_rel_l2(got, want) <= tol(n, dtype), (n, fwd)
        spec = checker.r2c(x, [1], True, 1.0)
        for fwd in (False, True):
            want = checker.c2r(spec, x.shape, [1], fwd, 1.0 / n)
            got = emu.nd("c2r", spec, np.zeros_like(x), x.shape, [1], fwd, 1.0 / n)
            assert oracle.max_row_rel_l2(got, want) <= tol(n, dtype), (n, fwd)
        # c2r ignores imag of bin 0 (and N/2): hdronly.h:3196-3215
        spec2 = spec.copy()
        spec2[:, 0] += 0.25j
        if n % 2 == 0:
            spec2[:, -1] -= 0.5j
        got2 = emu.nd("c2r", spec2, np.zeros_like(x), x.shape, [1], False, 1.0 / n)
        assert oracle.max_row_rel_l2(got2, checker.c2r(spec, x.shape, [1], False, 1.0 / n)) <= tol(n, dtype)


def test_halfcomplex_packed_roundtrip_all_classes(checker):
    """The reference's own sweep (tests/test_fft.nim:29-49) on a subsample of lengths, through the
    FFTPACK-packed in-place layout the C ABI uses, plus forward parity against the oracle."""
    rng = np.random.default_rng(7)
    odata = rng.uniform(-0.5, 0.5, 8192)
    odata[0] = 0.340188
    lengths = sorted(set(list(range(1, 200)) + list(range(200, 7000, 131)) + [1000, 3888, 4096, 4099]))
    for n in lengths:
        d = odata[:n].copy().reshape(1, n)
        emu.nd("r2c", d, d, d.shape, [1], True, 1.0, layout="halfcomplex")
        want = checker.rfft_rows(odata[:n].copy().reshape(1, n), True, 1.0)
        assert oracle.rel_l2(d, want) <= tol(n), n
        emu.nd("c2r", d, d, d.shape, [1], False, 1.0 / n, layout="halfcomplex")
        assert oracle.rel_l2(d[0], odata[:n]) <= 2e-15 * max(1, np.log2(n)), n


def test_real_layouts_times_forward_flag(checker, monkeypatch):
    """Every r2c output layout and c2r input layout with BOTH values of `forward` (include/impulse_fft_b200.h: forward = 0
    is exp(+...)): r2c(forward=False) stores the conjugate spectrum in the Hermitian, halfcomplex AND full-symmetric
    layouts, c2r(forward=True) conjugates its input in both input layouts, so r2c(fwd=f) and c2r(fwd=not f) round-trip.
    Short lines (one launch) and the long-line aux path (IMPULSE_FFT_FORCE_BIGREAL)."""
    rng = np.random.default_rng(81)

    def pack(spec, n):          # N/2+1 complex bins -> FFTPACK halfcomplex reals
        out = np.empty(spec.shape[:-1] + (n,))
        out[..., 0] = spec[..., 0].real
        k = (n - 1) // 2
        out[..., 1:2 * k:2] = spec[..., 1:k + 1].real
        out[..., 2:2 * k + 1:2] = spec[..., 1:k + 1].imag
        if n % 2 == 0:
            out[..., n - 1] = spec[..., n // 2].real
        return out

    for force in ("0", "1"):
        monkeypatch.setenv("IMPULSE_FFT_FORCE_BIGREAL", force)
        for n in (8, 9, 30, 64, 89, 100, 191, 256):
            x = rnd(rng, (3, n), np.float64)
            for fwd in (True, False):
                want = checker.r2c(x, [1], fwd, 0.5)          # conjugate spectrum for forward=False (hdronly.h:3152-3155)
                got = emu.nd("r2c", x, np.zeros((3, n // 2 + 1), np.complex128), x.shape, [1], fwd, 0.5)
                assert oracle.max_row_rel_l2(got, want) <= tol(n), (n, fwd, "hermitian")
                hc = emu.nd("r2c", x, np.zeros((3, n)), x.shape, [1], fwd, 0.5, layout="halfcomplex")
                assert oracle.max_row_rel_l2(hc, pack(want, n)) <= tol(n), (n, fwd, "halfcomplex")
                full = np.concatenate([want, np.conj(want[:, 1:(n + 1) // 2][:, ::-1])], axis=1)
                fs = emu.nd("r2c", x, np.zeros((3, n), np.complex128), x.shape, [1], fwd, 0.5, layout="fullsym")
                assert oracle.max_row_rel_l2(fs, full) <= tol(n), (n, fwd, "fullsym")
                # the matching inverse: c2r with the opposite flag undoes it in either input layout
                back = emu.nd("c2r", hc, np.zeros((3, n)), x.shape, [1], not fwd, 2.0 / n, layout="halfcomplex")
                assert oracle.max_row_rel_l2(back, x) <= tol(n), (n, fwd, "halfcomplex round trip")
                back = emu.nd("c2r", got, np.zeros((3, n)), x.shape, [1], not fwd, 2.0 / n)
                assert oracle.max_row_rel_l2(back, x) <= tol(n), (n, fwd, "hermitian round trip")


def test_fullsym_layout(checker):
    rng = np.random.default_rng(8)
    for n in (1, 2, 3, 4, 5, 8, 9, 16, 30, 89, 191):
        x = rnd(rng, (2, n), np.float64)
        want = checker.cfft_rows(x.astype(np.complex128), True, 1.0)
        got = emu.nd("r2c", x, np.zeros((2, n), np.complex128), x.shape, [1], True, 1.0, layout="fullsym")
        assert oracle.max_row_rel_l2(got, want) <= tol(n), n


def test_nd_and_strided(checker):
    rng = np.random.default_rng(9)
    # 2-D c2c, both axis orders, in place and out of place
    a = rnd(rng, (12, 20), np.complex128)
    for axes in ([0, 1], [1, 0], [0], [1]):
        want = checker.c2c(a, axes, True, 0.5)
        got = emu.nd("c2c", a, np.empty_like(a), a.shape, axes, True, 0.5)
        assert oracle.rel_l2(got, want) <= tol(20)
    b = a.copy()
    emu.nd("c2c", b, b, b.shape, [0, 1], False, 1.0)
    assert oracle.rel_l2(b, checker.c2c(a, [0, 1], False, 1.0)) <= tol(20)
    # 3-D, strided views, negative stride, fp32
    c = rnd(rng, (5, 6, 14), np.complex64)
    v = c[::-1, :, ::2]
    want = checker.c2c(v, [0, 2], True, 1.0)
    got = emu.nd("c2c", v, np.empty(v.shape, np.complex64), v.shape, [0, 2], True, 1.0)
    assert oracle.rel_l2(got, want) <= tol(7, np.float32)
    # transposed (column-major) input: axis 1 is the strided one
    t = rnd(rng, (16, 9), np.complex128).T
    assert oracle.rel_l2(emu.nd("c2c", t, np.empty(t.shape, np.complex128), t.shape, [0, 1], True, 1.0),
                         checker.c2c(t, [0, 1], True, 1.0)) <= tol(16)
    # 5-D: more batch dims than the kernel takes -> host loop
    e = rnd(rng, (2, 3, 2, 4, 6), np.complex128)[:, :, ::-1]
    assert oracle.rel_l2(emu.nd("c2c", e, np.empty(e.shape, np.complex128), e.shape, [4], True, 1.0),
                         checker.c2c(e, [4], True, 1.0)) <= tol(6)
    assert oracle.rel_l2(emu.nd("c2c", e, np.empty(e.shape, np.complex128), e.shape, [1, 3], True, 1.0),
                         checker.c2c(e, [1, 3], True, 1.0)) <= tol(6)
    # N-D r2c / c2r (hdronly.h:3334-3390)
    for shape, axes in (((8, 12), [0, 1]), ((5, 9), [0, 1]), ((5, 9), [1, 0]), ((4, 6, 10), [0, 2]), ((3, 7, 5), [0, 1, 2])):
        r = rnd(rng, shape, np.float64)
        want = checker.r2c(r, axes, True, 1.0)
        got = emu.nd("r2c", r, np.zeros(want.shape, np.complex128), shape, axes, True, 1.0)
        assert oracle.rel_l2(got, want) <= tol(12), (shape, axes)
        back = emu.nd("c2r", want, np.zeros(shape), shape, axes, False, 1.0 / np.prod([shape[a] for a in axes]))
        assert oracle.rel_l2(back, r) <= tol(12), (shape, axes)
        wantb = checker.c2r(want, shape, axes, False, 1.0)
        assert oracle.rel_l2(emu.nd("c2r", want, np.zeros(shape), shape, axes, False, 1.0), wantb) <= tol(12)
    # golden N-D fixtures
    a = GOLD["nd_c2c_f64_in"]
    assert oracle.rel_l2(emu.nd("c2c", a, np.empty_like(a), a.shape, [0, 1], True, 1.0), GOLD["nd_c2c_f64_ax01"]) <= tol(10)
    r = GOLD["nd_r2c_f32_in"]
    assert oracle.rel_l2(emu.nd("r2c", r, np.zeros((8, 7), np.complex64), r.shape, [0, 1], True, 1.0),
                         GOLD["nd_r2c_f32_ax01"]) <= tol(12, np.float32)


def test_errors_like_sanity_check():
    a = np.zeros((4, 4), np.complex128)
    with pytest.raises(emu.EmuError):
        emu.nd("c2c", a, a.copy(), a.shape, [2], True, 1.0)       # bad axis
    with pytest.raises(emu.EmuError):
        emu.nd("c2c", a, a.copy(), a.shape, [0, 0], True, 1.0)    # repeated axis
    z = np.zeros((0, 4), np.complex128)
    emu.nd("c2c", z, z.copy(), z.shape, [1], True, 1.0)            # empty: no-op (hdronly.h:3277)


def test_four_step_split(checker, monkeypatch):
    """Lines too long for one CTA (or strided and too long for a coalesced tile) are split N = N1*N2 over two
    launches with a twiddle in between; forced here at small sizes, and natural at 16384 / strided 4096."""
    rng = np.random.default_rng(12)
    # natural: contiguous 32768-point complex128 line does not fit 227 KB
    x = rnd(rng, (2, 32768), np.complex128)
    assert emu.nd_steps("c2c", x, x, x.shape, [1]) == 2
    for fwd in (True, False):
        got = emu.nd("c2c", x, np.empty_like(x), x.shape, [1], fwd, 0.5)
        assert oracle.max_row_rel_l2(got, checker.c2c(x, [1], fwd, 0.5)) <= tol(32768), fwd
    # natural: strided axis of 4096 points (8 adjacent columns of 4096 complex128 do not fit)
    y = rnd(rng, (4096, 12), np.complex128)
    assert emu.nd_steps("c2c", y, y, y.shape, [0]) == 2
    got = emu.nd("c2c", y, np.empty_like(y), y.shape, [0], True, 1.0)
    assert oracle.rel_l2(got, checker.c2c(y, [0], True, 1.0)) <= tol(4096)
    yi = y.copy()
    emu.nd("c2c", yi, yi, yi.shape, [0], False, 1.0)        # in place through the scratch buffer
    assert oracle.rel_l2(yi, checker.c2c(y, [0], False, 1.0)) <= tol(4096)
    # forced at small sizes: every layout class
    monkeypatch.setenv("IMPULSE_FFT_FORCE_FOURSTEP", "1")
    for n in (64, 72, 100, 128, 243, 360, 1000, 1024):
        z = rnd(rng, (3, n), np.complex128)
        assert emu.nd_steps("c2c", z, z, z.shape, [1]) == 2, n
        for fwd in (True, False):
            got = emu.nd("c2c", z, np.empty_like(z), z.shape, [1], fwd, 1.0)
            assert oracle.max_row_rel_l2(got, checker.c2c(z, [1], fwd, 1.0)) <= tol(n), (n, fwd)
    a = rnd(rng, (5, 96, 7), np.complex64)[::-1]
    got = emu.nd("c2c", a, np.empty(a.shape, np.complex64), a.shape, [1], True, 1.0)
    assert oracle.rel_l2(got, checker.c2c(a, [1], True, 1.0)) <= tol(96, np.float32)
    b = rnd(rng, (2, 3, 2, 64, 3), np.complex128)
    got = emu.nd("c2c", b, np.empty_like(b), b.shape, [3, 1], True, 1.0)
    assert oracle.rel_l2(got, checker.c2c(b, [3, 1], True, 1.0)) <= tol(64)
    r = rnd(rng, (80, 66), np.float64)
    spec = emu.nd("r2c", r, np.zeros((80, 34), np.complex128), r.shape, [0, 1], True, 1.0)
    assert oracle.rel_l2(spec, checker.r2c(r, [0, 1], True, 1.0)) <= tol(80)
    back = emu.nd("c2r", spec, np.zeros_like(r), r.shape, [0, 1], False, 1.0 / r.size)
    assert oracle.rel_l2(back, r) <= tol(80)


def test_multi_launch_bluestein(checker, monkeypatch):
    """Bluestein lengths whose n2 exceeds one CTA's shared memory (odd N > 7204 with a large prime
    factor, complex primes > 7232) run as stage-in / FFT(n2) x bkf / IFFT(n2) / stage-out launches."""
    rng = np.random.default_rng(13)
    # natural triggers at the top of the reference's 1..8191 sweep
    for n in (7211, 7919, 8191):
        r = rnd(rng, (2, n), np.float64)
        assert emu.nd_steps("r2c", r, np.zeros((2, n // 2 + 1), np.complex128), r.shape, [1]) == 6, n
        got = emu.nd("r2c", r, np.zeros((2, n // 2 + 1), np.complex128), r.shape, [1], True, 1.0)
        want = checker.r2c(r, [1], True, 1.0)
        assert oracle.max_row_rel_l2(got, want) <= tol(n), n
        back = emu.nd("c2r", want, np.zeros_like(r), r.shape, [1], False, 1.0 / n)
        assert oracle.max_row_rel_l2(back, r) <= tol(n), n
        d = r[:1].copy()
        emu.nd("r2c", d, d, d.shape, [1], True, 1.0, layout="halfcomplex")
        assert oracle.rel_l2(d, checker.rfft_rows(r[:1].copy(), True, 1.0)) <= tol(n), n
        emu.nd("c2r", d, d, d.shape, [1], False, 1.0 / n, layout="halfcomplex")
        assert oracle.rel_l2(d, r[:1]) <= 2e-15 * np.log2(n), n
    x = rnd(rng, (2, 7879), np.complex128)
    for fwd in (True, False):
        got = emu.nd("c2c", x, np.empty_like(x), x.shape, [1], fwd, 0.5)
        assert oracle.max_row_rel_l2(got, checker.c2c(x, [1], fwd, 0.5)) <= tol(7879), fwd
    # forced on small Bluestein lengths: every kind / layout / parity, batches and strides
    monkeypatch.setenv("IMPULSE_FFT_FORCE_BIGBLUE", "1")
    for n in (37, 74, 97, 134, 191, 382, 1009, 2018):
        c = rnd(rng, (3, n), np.complex128)
        for fwd in (True, False):
            got = emu.nd("c2c", c, np.empty_like(c), c.shape, [1], fwd, 1.0)
            assert oracle.max_row_rel_l2(got, checker.c2c(c, [1], fwd, 1.0)) <= tol(n), (n, fwd)
        r = rnd(rng, (3, n), np.float64)
        for fwd in (True, False):
            got = emu.nd("r2c", r, np.zeros((3, n // 2 + 1), np.complex128), r.shape, [1], fwd, 1.0)
            assert oracle.max_row_rel_l2(got, checker.r2c(r, [1], fwd, 1.0)) <= tol(n), (n, fwd)
        spec = checker.r2c(r, [1], True, 1.0)
        for fwd in (False, True):
            got = emu.nd("c2r", spec, np.zeros_like(r), r.shape, [1], fwd, 1.0 / n)
            assert oracle.max_row_rel_l2(got, checker.c2r(spec, r.shape, [1], fwd, 1.0 / n)) <= tol(n), (n, fwd)
        full = emu.nd("r2c", r, np.zeros((3, n), np.complex128), r.shape, [1], True, 1.0, layout="fullsym")
        assert oracle.max_row_rel_l2(full, checker.c2c(r.astype(np.complex128), [1], True, 1.0)) <= tol(n), n
    a = rnd(rng, (6, 37, 5), np.complex64)[:, :, ::-1]
    got = emu.nd("c2c", a, np.empty(a.shape, np.complex64), a.shape, [1], True, 1.0)
    assert oracle.rel_l2(got, checker.c2c(a, [1], True, 1.0)) <= tol(37, np.float32)


def test_dct_dst_of_any_length(checker, monkeypatch):
    """DCT / DST lines whose embedding does not fit one CTA run as embed -> complex transform -> extract (VERDICT r1:
    DST-I of 4096, DCT-II of 8192 and N = 10000 returned ERR_UNSUPPORTED; pocketfft_hdronly.h:2424-2648 takes any N).
    The same three-step plan forced on short lines, where every type / ortho combination is cheap to compare."""
    rng = np.random.default_rng(44)

    def want(cosine, t, x, fct, ortho):
        return checker.r2r(cosine, t, x, [x.ndim - 1], fct, ortho)

    for n, dt, rt in ((4096, np.float64, 1e-12), (8192, np.float64, 1e-12), (10000, np.float64, 1e-12), (10000, np.float32, 1e-5)):
        x = rnd(rng, (2, n), dt)
        for cosine, t in ((True, 1), (False, 1), (True, 2), (False, 3), (True, 4)):
            got = emu.r2r(cosine, t, x, np.empty_like(x), [1], 0.5, False)
            assert oracle.max_row_rel_l2(got, want(cosine, t, x, 0.5, False)) <= rt * np.log2(n), (n, cosine, t, dt)
    monkeypatch.setenv("IMPULSE_FFT_FORCE_BIGR2R", "1")
    for n in (2, 3, 8, 37, 100):
        x = rnd(rng, (3, n), np.float64)
        for cosine in (True, False):
            for t in (1, 2, 3, 4):
                for ortho in (False, True):
                    got = emu.r2r(cosine, t, x, np.empty_like(x), [1], 0.5, ortho)
                    assert oracle.max_row_rel_l2(got, want(cosine, t, x, 0.5, ortho)) <= 1e-12 * max(1, np.log2(n)), (n, cosine, t, ortho)
    a = rnd(rng, (5, 12, 6), np.float64)                 # strided axis, N-D, in place
    got = emu.r2r(True, 2, a, np.empty_like(a), [1, 0], 1.0, True)
    assert oracle.rel_l2(got, checker.r2r(True, 2, a, [1, 0], 1.0, True)) <= 1e-12 * 4
    b = a.copy()
    emu.r2r(False, 4, b, b, [1], 1.0, False)
    assert oracle.rel_l2(b, checker.r2r(False, 4, a, [1], 1.0, False)) <= 1e-12 * 4


def test_dct_dst_all_types(checker):
    """DCTDesc path: DCT/DST I-IV, ortho, fp64/fp32, N-D and strided, through the
    emulated engine against the compiled reference (or the O(N^2) definitions when it is absent)."""
    rng = np.random.default_rng(18)

    def want(cosine, t, x, axes, fct, ortho):
        if hasattr(checker, "r2r"):
            return checker.r2r(cosine, t, x, axes, fct, ortho)
        y = x.astype(np.float64)
        for i, ax in enumerate(axes):
            y = np.moveaxis(oracle.r2r_direct(cosine, t, np.moveaxis(y, ax, -1), fct if i == 0 else 1.0, ortho), -1, ax)
        return y

    for dt, rt in ((np.float64, 1e-12), (np.float32, 1e-5)):
        for n in (2, 3, 5, 8, 16, 37, 100, 128, 1000):
            x = rnd(rng, (3, n), dt)
            for cosine in (True, False):
                for t in (1, 2, 3, 4):
                    for ortho in (False, True):
                        got = emu.r2r(cosine, t, x, np.empty_like(x), [1], 0.5, ortho)
                        assert oracle.max_row_rel_l2(got, want(cosine, t, x, [1], 0.5, ortho)) <= rt * max(1, np.log2(n)), \
                            (n, cosine, t, ortho, dt)
    a = rnd(rng, (6, 10, 7), np.float64)
    for axes in ([0, 1, 2], [2, 0], [1]):
        got = emu.r2r(True, 2, a, np.empty_like(a), axes, 1.0, True)
        assert oracle.rel_l2(got, want(True, 2, a, axes, 1.0, True)) <= 1e-12 * 4
    v = rnd(rng, (8, 24), np.float64)[:, ::2]
    got = emu.r2r(False, 3, v, np.empty(v.shape), [0, 1], 1.0, False)
    assert oracle.rel_l2(got, want(False, 3, np.ascontiguousarray(v), [0, 1], 1.0, False)) <= 1e-12 * 4
    b = a.copy()
    emu.r2r(True, 4, b, b, [1], 1.0, False)  # in place
    assert oracle.rel_l2(b, want(True, 4, a, [1], 1.0, False)) <= 1e-12 * 4
    # DCT-II then DCT-III with 1/(2N) is the identity (orthogonality of the pair)
    x = rnd(rng, (4, 60), np.float64)
    y = emu.r2r(True, 3, emu.r2r(True, 2, x, np.empty_like(x), [1]), np.empty_like(x), [1], 1.0 / 120)
    assert oracle.max_row_rel_l2(y, x) <= 1e-14
    with pytest.raises(emu.EmuError):
        emu.r2r(True, 1, np.zeros((2, 1)), np.zeros((2, 1)), [1])   # DCT-I of one point (pocketfft throws too)


@pytest.mark.parametrize("shape,axes,period", [
    ((6, 64), [1], 64),          # one row of multipliers broadcast over the batch
    ((3, 40, 24), [1], 40 * 24),  # strided axis, one image of multipliers
    ((3, 40, 24), [1, 2], 40 * 24),
    ((2, 4099), [1], 2 * 4099),  # Bluestein line, no broadcast
    ((5, 16384), [1], 16384),    # split into two launches: the multiply rides on the second
])
@pytest.mark.parametrize("forward", [True, False])
def test_c2c_fused_multiply(shape, axes, period, forward):
    """impulse_fft_c2c_mul: out = c2c(in) * mul[offset % period]; checked against numpy."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    m = rng.standard_normal(period) + 1j * rng.standard_normal(period)
    want = (np.fft.fftn(x, axes=axes) if forward else np.fft.ifftn(x, axes=axes, norm="forward"))
    want = (want.reshape(-1, period) * m).reshape(shape)
    got = emu.c2c_mul(x, np.empty_like(x), axes, m, forward=forward)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13


R2R_REAL_CASES = [((7,), [0]), ((8,), [0]), ((1,), [0]), ((2,), [0]), ((3, 10), [1]), ((6, 9), [0, 1]), ((4, 5, 6), [2, 0]),
                  ((2, 191), [1]), ((3, 1000), [1]), ((2, 4099), [1]), ((40, 24), [0]), ((12, 16, 10), [0, 1, 2])]


@pytest.mark.parametrize("shape,axes", R2R_REAL_CASES)
def test_r2r_fftpack(shape, axes):
    """impulse_fft_r2r_fftpack against the (reference-pinned) restatement, all four flag combinations."""
    rng = np.random.default_rng(23)
    a = rng.standard_normal(shape)
    for r2h in (True, False):
        for fwd in (True, False):
            want = oracle.fftpack_numpy(a, axes, r2h, fwd, 0.25)
            got = emu.r2r_real("fftpack", a, np.empty_like(a), axes, r2h, fwd, 0.25)
            assert oracle.rel_l2(got, want) < 2e-14, (r2h, fwd)
    b = a.copy()                                    # in place
    emu.r2r_real("fftpack", b, b, axes, True, True, 1.0)
    assert oracle.rel_l2(b, oracle.fftpack_numpy(a, axes, True, True, 1.0)) < 2e-14


@pytest.mark.parametrize("shape,axes", R2R_REAL_CASES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_hartley(shape, axes, dtype):
    rng = np.random.default_rng(29)
    a = rng.standard_normal(shape).astype(dtype)
    tol = 2e-14 if dtype == np.float64 else 2e-5
    for which, genuine in (("separable_hartley", False), ("genuine_hartley", True)):
        want = oracle.hartley_numpy(a, axes, genuine, 0.5)
        got = emu.r2r_real(which, a, np.empty_like(a), axes, fct=0.5)
        assert oracle.rel_l2(got.astype(np.float64), want) < tol, which
    # strided output, non-contiguous input
    big = rng.standard_normal(tuple(2 * s for s in shape)).astype(dtype)
    view = big[tuple(slice(None, None, 2) for _ in shape)]
    out = np.zeros_like(big)
    oview = out[tuple(slice(None, None, 2) for _ in shape)]
    emu.r2r_real("genuine_hartley", view, oview, axes)
    assert oracle.rel_l2(oview.astype(np.float64), oracle.hartley_numpy(view, axes, True)) < tol


@pytest.mark.parametrize("shape,axis", [((4, 64), 1), ((3, 40, 24), 1), ((2, 1024, 12), 1), ((5, 4099), 1), ((3, 16384), 1)])
def test_convolve_axis(shape, axis):
    """impulse_fft_convolve_axis (plain plan): IFFT(FFT(x) * m) along one axis, m broadcast over the batch."""
    rng = np.random.default_rng(37)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    period = int(np.prod(shape[1:]))
    m = rng.standard_normal(period) + 1j * rng.standard_normal(period)
    n = shape[axis]
    want = np.fft.ifft(np.fft.fft(x, axis=axis) * m.reshape(shape[1:]), axis=axis)
    got = emu.convolve_axis(x, np.empty_like(x), axis, m, fct=1.0 / n)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13
    y = x.copy()
    emu.convolve_axis(y, y, axis, m, fct=1.0 / n)   # in place
    assert np.linalg.norm(y - want) / np.linalg.norm(want) < 1e-13


def _real_roundtrip_case(n, rows, dtype, layout="hermitian"):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((rows, n)).astype(dtype)
    cdt = np.complex128 if dtype == np.float64 else np.complex64
    tol = 1e-12 * np.log2(max(n, 2)) if dtype == np.float64 else 1e-5 * np.log2(max(n, 2))
    want = np.fft.rfft(x.astype(np.float64), axis=1)
    if layout == "hermitian":
        spec = emu.nd("r2c", x, np.empty((rows, n // 2 + 1), cdt), [rows, n], [1], True, 0.5)
        assert oracle.rel_l2(spec, 0.5 * want) < tol, (n, "r2c")
        back = emu.nd("c2r", spec, np.empty_like(x), [rows, n], [1], False, 2.0 / n)
        assert oracle.rel_l2(back, x) < tol, (n, "c2r")
        conj = emu.nd("r2c", x, np.empty((rows, n // 2 + 1), cdt), [rows, n], [1], False, 1.0)
        assert oracle.rel_l2(conj, np.conj(want)) < tol, (n, "r2c backward")
    else:
        packed = emu.nd("r2c", x, np.empty_like(x), [rows, n], [1], True, 1.0, layout="halfcomplex")
        assert oracle.rel_l2(packed, oracle._halfcomplex_pack(want, n)) < tol, (n, "packed")
        back = emu.nd("c2r", packed, np.empty_like(x), [rows, n], [1], False, 1.0 / n, layout="halfcomplex")
        assert oracle.rel_l2(back, x) < tol, (n, "unpacked")


@pytest.mark.parametrize("n", [32768, 65536, 30000, 28930, 29999, 40001])
def test_long_real_lines(n):
    """Real transforms whose complex line does not fit one CTA (N/2 > 14464, odd N > 14464): complex
    transform on a work array (split / Bluestein) with elementwise conversion passes around it."""
    _real_roundtrip_case(n, 2, np.float64)
    if n in (32768, 29999):
        _real_roundtrip_case(n, 2, np.float64, "halfcomplex")
        _real_roundtrip_case(n, 1, np.float32)


@pytest.mark.parametrize("n", [6, 7, 64, 100, 191, 1000, 4099])
def test_long_real_path_forced_on_short_lines(n, monkeypatch):
    """The same plan shape forced on short lines, where every layout can be compared cheaply."""
    monkeypatch.setenv("IMPULSE_FFT_FORCE_BIGREAL", "1")
    _real_roundtrip_case(n, 3, np.float64)
    _real_roundtrip_case(n, 3, np.float64, "halfcomplex")
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, 3, n))
    full = emu.nd("r2c", x, np.empty((2, 3, n), np.complex128), [2, 3, n], [2], True, 1.0, layout="fullsym")
    assert oracle.rel_l2(full, np.fft.fft(x, axis=2)) < 1e-12
    hart = emu.r2r_real("separable_hartley", x, np.empty_like(x), [2])
    assert oracle.rel_l2(hart, oracle.hartley_numpy(x, [2])) < 1e-12
    # strided axis, and rows an odd number of elements apart: even N takes the gather / scatter passes (the packed
    # complex view of the real side needs contiguous lines and even row strides), as general_r2c / general_c2r accept
    # any byte stride (pocketfft_hdronly.h:3125-3250)
    y = rng.standard_normal((n, 4))
    spec = emu.nd("r2c", y, np.empty((n // 2 + 1, 4), np.complex128), [n, 4], [0], True, 1.0)
    assert oracle.rel_l2(spec, np.fft.rfft(y, axis=0)) < 1e-12
    back = emu.nd("c2r", spec, np.empty_like(y), [n, 4], [0], False, 1.0 / n)
    assert oracle.rel_l2(back, y) < 1e-12
    wide = rng.standard_normal((3, n + 1))
    yv = wide[:, :n]                                      # row stride n + 1: odd for even n
    spec = emu.nd("r2c", yv, np.empty((3, n // 2 + 1), np.complex128), [3, n], [1], False, 2.0)
    assert oracle.rel_l2(spec, 2.0 * np.conj(np.fft.rfft(yv, axis=1))) < 1e-12
    outw = np.zeros((3, n + 1))
    back = emu.nd("c2r", np.conj(spec), outw[:, :n], [3, n], [1], False, 0.5 / n)
    assert oracle.rel_l2(back, yv) < 1e-12 and not outw[:, n].any()
    # r2r_fftpack with real2hermitian != forward (negated elements 2, 4, ... of the real side) on this path
    a = rng.standard_normal((2, n))
    for r2h in (True, False):
        for fwd in (True, False):
            got = emu.r2r_real("fftpack", a, np.empty_like(a), [1], r2h, fwd, 0.25)
            assert oracle.rel_l2(got, oracle.fftpack_numpy(a, [1], r2h, fwd, 0.25)) < 2e-13, (r2h, fwd)


def test_long_even_real_lines_at_any_stride():
    """r2c over axis 0 of a [40000, 4] array and a [3, 40000] array with an odd row stride used to
    return ERR_UNSUPPORTED; pocketfft takes any byte stride."""
    rng = np.random.default_rng(40000)
    y = rng.standard_normal((40000, 4))
    spec = emu.nd("r2c", y, np.empty((20001, 4), np.complex128), [40000, 4], [0], True, 1.0)
    assert oracle.rel_l2(spec, np.fft.rfft(y, axis=0)) < 1e-12 * 16
    back = emu.nd("c2r", spec, np.empty_like(y), [40000, 4], [0], False, 1.0 / 40000)
    assert oracle.rel_l2(back, y) < 1e-12 * 16
    wide = rng.standard_normal((3, 40001))
    yv = wide[:, :40000]
    spec = emu.nd("r2c", yv, np.empty((3, 20001), np.complex128), [3, 40000], [1], True, 1.0)
    assert oracle.rel_l2(spec, np.fft.rfft(yv, axis=1)) < 1e-12 * 16
    a = rng.standard_normal((1, 32768))                  # r2r_fftpack, real2hermitian != forward, beyond one CTA
    got = emu.r2r_real("fftpack", a, np.empty_like(a), [1], True, False, 1.0)
    assert oracle.rel_l2(got, oracle.fftpack_numpy(a, [1], True, False, 1.0)) < 1e-12 * 16


@pytest.mark.parametrize("n", [89, 191, 4099, 100003, 20011])
def test_long_bluestein_lines(n, monkeypatch):
    """Complex Bluestein lengths whose LINE does not fit one CTA (elementwise chirp passes), and the same plan
    forced on short prime lengths."""
    if n < 20000:
        monkeypatch.setenv("IMPULSE_FFT_FORCE_BIGBLUE", "1")
        monkeypatch.setenv("IMPULSE_FFT_FORCE_AUXBLUE", "1")
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
    tol = 1e-12 * np.log2(n)
    for fwd in (True, False):
        got = emu.nd("c2c", x, np.empty_like(x), [2, n], [1], fwd, 0.5)
        want = 0.5 * (np.fft.fft(x, axis=1) if fwd else np.fft.ifft(x, axis=1) * n)
        assert oracle.rel_l2(got, want) < tol, (n, fwd)
    if n == 20011:                                        # odd real line on top of it
        _real_roundtrip_case(n, 1, np.float64)


def test_split_beyond_four_million_points():
    """The two-launch split has no table of the whole length: 2^23 points (sub-transforms 2048 x 4096)."""
    n = 1 << 23
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).reshape(1, n)
    got = emu.nd("c2c", x, np.empty_like(x), [1, n], [1], True, 1.0)
    assert oracle.rel_l2(got, np.fft.fft(x, axis=1)) < 1e-14
