"""CPU-only check of fast3_kernel's index algebra through the thread-per-thread host emulation
(tests/emu/emu_fast3.cpp): the SAME kernel body the GPU runs, compiled with g++.  The paths that were already
parity-green on the B200 (c2c, r2c/c2r through shared memory) pin the emulation itself; the in-register pair
post-twiddle (PAIR) is then checked the same way.  Semantics: pocketfft r2c/c2r with `forward`
(pocketfft_hdronly.h:3125-3250) restated with numpy."""
import numpy as np
import pytest

from tests.emu import harness_fast3 as f3

SHAPES_PAIR = [(16, 16, 8, 16), (16, 8, 8, 16), (8, 8, 4, 8), (10, 10, 5, 10), (18, 18, 6, 18), (8, 8, 8, 16)]
SHAPES_OLD = [(8, 8, 8, 8), (5, 10, 10, 10), (6, 18, 18, 18), (16, 16, 16, 16), (10, 10, 10, 10)]


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def ref_r2c(x, forward, fct):
    X = np.fft.rfft(x.astype(np.float64), axis=1)
    return (X if forward else np.conj(X)) * fct


def ref_c2r(X, n, forward, fct):
    X = X.astype(np.complex128).copy()
    if forward:
        X = np.conj(X)
    return np.fft.irfft(X, n=n, axis=1) * n * fct


@pytest.mark.parametrize("shape", SHAPES_PAIR + SHAPES_OLD)
def test_emulation_matches_known_good_paths(shape):
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(n)
    x = rng.random((3, 2 * n)) - 0.5
    assert rel(f3.run(shape, "r2c", x, True, 0.5), ref_r2c(x, True, 0.5)) < 2e-15 * np.log2(n) * 4
    X = ref_r2c(x, True, 1.0)
    X[:, 0] += 0.25j   # the imaginary parts of bins 0 and N must be ignored
    assert rel(f3.run(shape, "c2r", X, False, 1.0 / (2 * n)), x) < 2e-15 * np.log2(n) * 4
    if shape in ((16, 16, 8, 16), (16, 16, 16, 16), (5, 10, 10, 10)):
        z = rng.random((2, n)) - 0.5 + 1j * (rng.random((2, n)) - 0.5)
        assert rel(f3.run(shape, "c2c", z, True, 2.0), np.fft.fft(z, axis=1) * 2.0) < 2e-15 * np.log2(n) * 4
        assert rel(f3.run(shape, "c2c", z, False, 1.0), np.fft.ifft(z, axis=1) * n) < 2e-15 * np.log2(n) * 4


@pytest.mark.parametrize("shape", SHAPES_PAIR)
@pytest.mark.parametrize("forward", [True, False])
def test_pair_post_twiddle_f64(shape, forward):
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(7 * n + forward)
    x = rng.random((5, 2 * n)) - 0.5     # 5 rows over 2 CTAs: exercises the dynamic row claims
    got = f3.run(shape, "r2c", x, forward, 0.75, pair=True, ctas=2)
    assert not np.isnan(got.view(np.float64)).any()      # every bin 0..N written
    want = ref_r2c(x, forward, 0.75)
    assert rel(got, want) < 2e-15 * np.log2(n) * 4
    for r in range(x.shape[0]):
        assert rel(got[r], want[r]) < 2e-15 * np.log2(n) * 4
    # same numbers as the shared-memory post-twiddle to rounding
    old = f3.run(shape, "r2c", x, forward, 0.75, pair=False, ctas=1)
    assert rel(got, old) < 1e-15


@pytest.mark.parametrize("shape", [(16, 16, 8, 16), (16, 8, 8, 16), (8, 8, 4, 8)])
def test_pair_post_twiddle_f32(shape):
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(n + 1)
    x = (rng.random((3, 2 * n)) - 0.5).astype(np.float32)
    got = f3.run(shape, "r2c", x, True, 1.0, pair=True)
    assert got.dtype == np.complex64 and not np.isnan(got.view(np.float32)).any()
    assert rel(got.astype(np.complex128), ref_r2c(x, True, 1.0)) < 1e-6 * np.log2(n)


SHAPES_C2R_PAIR = [(5, 10, 10, 10), (6, 18, 18, 18), (8, 16, 16, 16), (8, 8, 16, 16), (8, 8, 8, 16), (4, 8, 8, 8)]


@pytest.mark.parametrize("shape", SHAPES_C2R_PAIR)
@pytest.mark.parametrize("forward", [False, True])
def test_pair_pre_twiddle_f64(shape, forward):
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(11 * n + forward)
    X = rng.random((5, n + 1)) - 0.5 + 1j * (rng.random((5, n + 1)) - 0.5)   # bins 0 and N carry junk imaginary parts
    got = f3.run(shape, "c2r", X, forward, 0.5 / n, pair=True, ctas=2)
    assert not np.isnan(got).any()
    Xc = X.copy()
    Xc[:, 0] = Xc[:, 0].real
    Xc[:, n] = Xc[:, n].real
    want = ref_c2r(Xc, 2 * n, forward, 0.5 / n)
    for r in range(X.shape[0]):
        assert rel(got[r], want[r]) < 2e-15 * np.log2(n) * 4
    old = f3.run(shape, "c2r", X, forward, 0.5 / n, pair=False, ctas=1)
    assert rel(got, old) < 1e-15


@pytest.mark.parametrize("shape", [(8, 16, 16, 16), (8, 8, 16, 16), (4, 8, 8, 8)])
def test_pair_pre_twiddle_f32(shape):
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(n + 2)
    x = rng.random((3, 2 * n)) - 0.5
    X = np.fft.rfft(x, axis=1).astype(np.complex64)
    got = f3.run(shape, "c2r", X, False, 0.5 / n, pair=True)
    assert got.dtype == np.float32 and not np.isnan(got).any()
    assert rel(got.astype(np.float64), x) < 1e-6 * np.log2(n)


@pytest.mark.parametrize("shape,kind", [((16, 16, 8, 16), "r2c"), ((18, 18, 6, 18), "r2c"), ((10, 10, 5, 10), "r2c"),
                                        ((16, 16, 16, 16), "c2c"), ((16, 16, 8, 16), "c2c"), ((10, 10, 10, 10), "c2c")])
@pytest.mark.parametrize("forward", [True, False])
def test_register_prefetch_of_next_row(shape, kind, forward):
    """PF variant: the next claimed row is loaded into registers during pass 3.  7 rows over 2 CTAs, so that every
    CTA runs several rows back to back (first row loaded at the top, later rows prefetched, last one without a successor)."""
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(3 * n + forward)
    if kind == "r2c":
        x = rng.random((7, 2 * n)) - 0.5
        want = ref_r2c(x, forward, 1.25)
    elif kind == "c2r":
        x = rng.random((7, n + 1)) - 0.5 + 1j * (rng.random((7, n + 1)) - 0.5)
        xc = x.copy()
        xc[:, 0] = xc[:, 0].real
        xc[:, n] = xc[:, n].real
        want = ref_c2r(xc, 2 * n, forward, 1.25)
    else:
        x = rng.random((7, n)) - 0.5 + 1j * (rng.random((7, n)) - 0.5)
        want = (np.fft.fft(x, axis=1) if forward else np.fft.ifft(x, axis=1) * n) * 1.25
    got = f3.run(shape, kind, x, forward, 1.25, pair=kind != "c2c", ctas=2, prefetch=True)
    for r in range(7):
        assert rel(got[r], want[r]) < 2e-15 * np.log2(n) * 4, r


@pytest.mark.parametrize("shape", [(5, 10, 10, 10), (10, 10, 5, 10), (18, 18, 6, 18), (6, 18, 18, 18), (10, 10, 10, 10)])
def test_non_power_of_two_shapes_f32(shape):
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(n + 5)
    x = (rng.random((3, 2 * n)) - 0.5).astype(np.float32)
    pr = (r1 * r2) % 2 == 0 and e // r3 >= 2
    got = f3.run(shape, "r2c", x, True, 1.0, pair=pr)
    assert rel(got.astype(np.complex128), ref_r2c(x, True, 1.0)) < 1e-6 * np.log2(n)
    X = np.fft.rfft(x.astype(np.float64), axis=1).astype(np.complex64)
    pc = (r2 * r3) % 2 == 0 and e // r1 >= 2
    back = f3.run(shape, "c2r", X, False, 0.5 / n, pair=pc)
    assert rel(back.astype(np.float64), x.astype(np.float64)) < 1e-6 * np.log2(n)
    z = (rng.random((2, n)) - 0.5 + 1j * (rng.random((2, n)) - 0.5)).astype(np.complex64)
    assert rel(f3.run(shape, "c2c", z, True, 1.0).astype(np.complex128), np.fft.fft(z.astype(np.complex128), axis=1)) < 1e-6 * np.log2(n)


@pytest.mark.parametrize("shape,kind", [((16, 16, 8, 16), "r2c"), ((18, 18, 6, 18), "r2c"), ((10, 10, 5, 10), "r2c"),
                                        ((8, 16, 16, 16), "c2r"), ((6, 18, 18, 18), "c2r"), ((5, 10, 10, 10), "c2r"),
                                        ((16, 16, 8, 16), "c2c")])
def test_second_exchange_buffer(shape, kind):
    """DB variant (two barriers per row): 9 rows over 2 CTAs, repeated, since a missing barrier shows up as a race."""
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(5 * n)
    for rep in range(3):
        if kind == "r2c":
            x = rng.random((9, 2 * n)) - 0.5
            want = ref_r2c(x, True, 1.0)
        elif kind == "c2r":
            x = np.fft.rfft(rng.random((9, 2 * n)) - 0.5, axis=1)
            want = ref_c2r(x, 2 * n, False, 1.0)
        else:
            x = rng.random((9, n)) - 0.5 + 1j * (rng.random((9, n)) - 0.5)
            want = np.fft.fft(x, axis=1)
        got = f3.run(shape, kind, x, True if kind != "c2r" else False, 1.0, pair=kind != "c2c", ctas=2, double_buffer=True)
        for r in range(9):
            assert rel(got[r], want[r]) < 2e-15 * np.log2(n) * 4, (rep, r)


@pytest.mark.parametrize("shape", [(8, 24, 8, 24), (10, 20, 10, 20), (10, 20, 20, 20)])
def test_more_mixed_radix_shapes(shape):
    """1536 = 8*24*8, 2000 = 10*20*10, 4000 = 10*20*20 (and the real rows of twice those lengths)."""
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(n + 9)
    z = rng.random((3, n)) - 0.5 + 1j * (rng.random((3, n)) - 0.5)
    assert rel(f3.run(shape, "c2c", z, True, 1.0), np.fft.fft(z, axis=1)) < 2e-15 * np.log2(n) * 4
    assert rel(f3.run(shape, "c2c", z, False, 1.0 / n), np.fft.ifft(z, axis=1)) < 2e-15 * np.log2(n) * 4
    x = rng.random((5, 2 * n)) - 0.5
    pr = e // r3 >= 2
    assert rel(f3.run(shape, "r2c", x, True, 1.0, pair=pr), ref_r2c(x, True, 1.0)) < 2e-15 * np.log2(n) * 4
    X = np.fft.rfft(x, axis=1)
    assert rel(f3.run(shape, "c2r", X, False, 0.5 / n, pair=True), x) < 2e-15 * np.log2(n) * 4


@pytest.mark.parametrize("db", [False, True])
@pytest.mark.parametrize("shape,kind", [((16, 16, 8, 16), "r2c"), ((18, 18, 6, 18), "r2c"), ((10, 10, 5, 10), "r2c"),
                                        ((8, 16, 16, 16), "c2r"), ((6, 18, 18, 18), "c2r"), ((5, 10, 10, 10), "c2r"),
                                        ((16, 16, 8, 16), "c2c"), ((16, 16, 16, 16), "c2c")])
def test_rows_staged_by_bulk_copy(shape, kind, db):
    """TMA variant: the next claimed row is copied into a shared staging buffer (cp.async.bulk + mbarrier on the GPU, a
    memcpy + phase word here) while the current row is transformed.  11 rows over 3 CTAs, both directions, repeated:
    a copy issued before every reader has left the staging buffer, or a wrong phase, shows up as a wrong row."""
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    if db and kind == "c2c" and shape == (16, 16, 16, 16):
        pytest.skip("no second exchange buffer for 4096-point complex rows")
    rng = np.random.default_rng(3 * n + db)
    for rep in range(2):
        for fwd in (True, False):
            if kind == "r2c":
                x = rng.random((11, 2 * n)) - 0.5
                want = ref_r2c(x, fwd, 0.5)
            elif kind == "c2r":
                x = np.fft.rfft(rng.random((11, 2 * n)) - 0.5, axis=1)
                x[:, 0] += 0.125j                       # ignored imaginary parts
                want = ref_c2r(np.where(np.arange(n + 1) % n == 0, x.real, x), 2 * n, fwd, 0.5)
            else:
                x = rng.random((11, n)) - 0.5 + 1j * (rng.random((11, n)) - 0.5)
                want = (np.fft.fft(x, axis=1) if fwd else np.fft.ifft(x, axis=1) * n) * 0.5
            got = f3.run(shape, kind, x, fwd, 0.5, pair=kind != "c2c", ctas=3, double_buffer=db, staged=True)
            assert not np.isnan(got.view(np.float64)).any()
            for r in range(11):
                assert rel(got[r], want[r]) < 2e-15 * np.log2(n) * 4, (rep, fwd, r)


def test_rows_staged_by_bulk_copy_f32():
    rng = np.random.default_rng(99)
    n = 2048
    x = (rng.random((7, 2 * n)) - 0.5).astype(np.float32)
    got = f3.run((16, 16, 8, 16), "r2c", x, True, 1.0, pair=True, ctas=2, staged=True)
    assert rel(got.astype(np.complex128), ref_r2c(x, True, 1.0)) < 1e-6 * np.log2(n)
    X = np.fft.rfft(x.astype(np.float64), axis=1).astype(np.complex64)
    got = f3.run((8, 16, 16, 16), "c2r", X, False, 0.5 / n, pair=True, ctas=2, staged=True, double_buffer=True)
    assert rel(got.astype(np.float64), x.astype(np.float64)) < 1e-6 * np.log2(n)


@pytest.mark.parametrize("shape", [(27, 9, 9, 27), (10, 30, 10, 30), (27, 27, 9, 27)])
def test_round2_complex_shapes(shape):
    """3^7 = 27*9*9, 3000 = 10*30*10, 3^8 = 27*27*9: complex rows (radix-27 = 3 x 9 and radix-30 = 5 x 6 composites)."""
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    rng = np.random.default_rng(n + 1)
    z = rng.random((4, n)) - 0.5 + 1j * (rng.random((4, n)) - 0.5)
    got = f3.run(shape, "c2c", z, True, 0.5, ctas=2)
    want = np.fft.fft(z, axis=1) * 0.5
    for r in range(4):
        assert rel(got[r], want[r]) < 2e-15 * np.log2(n) * 4, r
    assert rel(f3.run(shape, "c2c", z, False, 1.0 / n, ctas=2), np.fft.ifft(z, axis=1)) < 2e-15 * np.log2(n) * 4
