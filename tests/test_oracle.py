"""Pin the oracle (CPU, no GPU): golden vectors of the reference's own tests/README,
the compiled reference (oracle/_ref) and this repo's restatement (oracle/pocketfft_port.c)
must all agree.  Citations: README.md, tests/test_fft.nim (T1), tests/test_fft2.nim (T2)."""
import os

import numpy as np
import pytest

from oracle import nim_helpers as nh
from oracle import oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "pocketfft_golden.npz"))

README_IN = np.array([1.0, 2.0, 1.0, -1.0, 1.5])
README_PACKED = np.array([4.5, 2.081559480312316, -1.651098762732523, -1.831559480312316,
                          1.608220406444071])  # README.md:41
README_FULL = np.array([4.5 + 0j, 2.081559480312316 - 1.651098762732523j,
                        -1.831559480312316 + 1.608220406444071j,
                        -1.831559480312316 - 1.608220406444071j,
                        2.081559480312316 + 1.651098762732523j])  # README.md:68


def engines(request):
    out = [request.getfixturevalue("port")]
    try:
        out.append(oracle.Ref())
    except (FileNotFoundError, OSError):
        pass
    return out


def test_readme_vectors(request):
    for e in engines(request):
        api = nh.NimApi(e)
        np.testing.assert_allclose(api.rfft_packed(README_IN), README_PACKED, rtol=0, atol=2e-15)
        np.testing.assert_allclose(api.fft(README_IN), README_FULL, rtol=0, atol=2e-15)
        np.testing.assert_allclose(api.fft(README_IN.astype(np.complex128)), README_FULL, rtol=0, atol=2e-15)
        back = api.fft(api.fft(README_IN), forward=False)
        np.testing.assert_allclose(back.real, README_IN, atol=1e-10)  # README.md:39
        # C++ r2c into a 5-slot buffer: 3 bins written, 2 untouched (README.md:98)
        out = np.zeros(5, dtype=np.complex128)
        e.r2c(README_IN, [0], True, 1.0, out=out[:3])
        np.testing.assert_allclose(out[:3], README_FULL[:3], rtol=0, atol=2e-15)
        assert out[3] == 0 and out[4] == 0


def test_t2_known_answers(request):
    """tests/test_fft2.nim:5-15."""
    expected = np.array([6, -2 + 2j, -2, -2 - 2j], dtype=np.complex128)
    x = np.arange(4.0)
    for e in engines(request):
        api = nh.NimApi(e)
        assert np.array_equal(api.fft(x.astype(np.complex128)), expected)
        assert np.array_equal(api.fft(x), expected)
        assert np.array_equal(api.fft(x, forward=True), expected)
        rt = api.ifft(api.fft(x))
        assert np.array_equal(rt.real, x)
        assert np.array_equal(rt.imag, np.zeros(4))
        assert np.array_equal(api.fft(x, normalize=nh.NK_FORWARD), expected / 4.0)
        assert np.array_equal(api.fft(x, normalize=nh.NK_ORTHO), expected / 2.0)
        assert np.array_equal(api.ifft(x), api.fft(x, forward=False))


def test_unpack_symmetrize_layout():
    # even: [r0, r1,i1, r2]; odd: [r0, r1,i1, r2,i2]   (pocketfft.nim:228-238)
    np.testing.assert_array_equal(nh.unpack_fft(np.array([1.0, 2, 3, 4])), [1, 2 + 3j, 4])
    np.testing.assert_array_equal(nh.unpack_fft(np.array([1.0, 2, 3, 4, 5])), [1, 2 + 3j, 4 + 5j])
    np.testing.assert_array_equal(nh.symmetrize(np.array([1.0, 2, 3, 4])), [1, 2 + 3j, 4, 2 - 3j])
    np.testing.assert_array_equal(nh.symmetrize(np.array([1.0, 2, 3, 4, 5])),
                                  [1, 2 + 3j, 4 + 5j, 4 - 5j, 2 - 3j])
    # complex input: parity guessed from imag of last bin (pocketfft.nim:173-180)
    assert nh.symm_target_size(np.array([1, 2 + 3j, 4 + 0j])) == 4
    assert nh.symm_target_size(np.array([1, 2 + 3j, 4 + 5j])) == 5


def test_normalize_table():
    assert nh.init_normalize(nh.NK_BACKWARD, True, np.inf, 8) == 1.0
    assert nh.init_normalize(nh.NK_BACKWARD, False, np.inf, 8) == 0.125
    assert nh.init_normalize(nh.NK_FORWARD, True, np.inf, 8) == 0.125
    assert nh.init_normalize(nh.NK_FORWARD, False, np.inf, 8) == 1.0
    assert nh.init_normalize(nh.NK_ORTHO, False, np.inf, 16) == 0.25
    assert nh.init_normalize(nh.NK_CUSTOM, True, 3.0, 16) == 3.0


def test_port_matches_golden(port):
    """The restatement against outputs of the compiled reference (committed fixture)."""
    for key in GOLD.files:
        if key.startswith("c2c_f64_in_"):
            n = int(key.rsplit("_", 1)[1])
            x = GOLD[key]
            got = port.cfft_rows(x.copy(), True, 1.0)
            assert oracle.max_row_rel_l2(got, GOLD[f"c2c_f64_fwd_{n}"]) <= 4e-15, n
            got = port.cfft_rows(x.copy(), False, 1.0 / n)
            assert oracle.max_row_rel_l2(got, GOLD[f"c2c_f64_bwd_{n}"]) <= 4e-15, n
        if key.startswith("r_f64_in_"):
            n = int(key.rsplit("_", 1)[1])
            p = port.rfft_rows(GOLD[key].copy(), True, 1.0)
            assert oracle.max_row_rel_l2(p, GOLD[f"r_f64_packed_fwd_{n}"]) <= 4e-15, n
            b = port.rfft_rows(GOLD[f"r_f64_packed_fwd_{n}"].copy(), False, 1.0 / n)
            assert oracle.max_row_rel_l2(b, GOLD[f"r_f64_packed_bwd_{n}"]) <= 4e-15, n
    a = GOLD["nd_c2c_f64_in"]
    assert oracle.rel_l2(port.c2c(a, [0, 1], True, 1.0), GOLD["nd_c2c_f64_ax01"]) <= 4e-15
    assert oracle.rel_l2(port.c2c(a, [0], False, 0.25), GOLD["nd_c2c_f64_ax0_bwd"]) <= 4e-15
    b = GOLD["nd_c2c_f32_in"]
    assert oracle.rel_l2(port.c2c(b, [1, 2], True, 1.0), GOLD["nd_c2c_f32_ax12"]) <= 2e-6
    r = GOLD["nd_r2c_f32_in"]
    assert oracle.rel_l2(port.r2c(r, [0, 1], True, 1.0), GOLD["nd_r2c_f32_ax01"]) <= 2e-6
    assert oracle.rel_l2(port.r2c(r, [1], False, 1.0), GOLD["nd_r2c_f32_ax1_bwd"]) <= 2e-6
    r64 = GOLD["nd_r2c_f64_in"]
    assert oracle.rel_l2(port.r2c(r64, [0, 1], True, 1.0), GOLD["nd_r2c_f64_ax01"]) <= 4e-15
    assert oracle.rel_l2(port.c2r(GOLD["nd_r2c_f64_ax01"], (5, 9), [0, 1], False, 1.0 / 45),
                         GOLD["nd_c2r_f64_ax01"]) <= 4e-15
    assert oracle.rel_l2(port.c2r(GOLD["nd_r2c_f64_ax0"], (5, 9), [0], False, 0.2),
                         GOLD["nd_c2r_f64_ax0"]) <= 4e-15


def test_golden_roundtrip_consistency():
    """The fixture itself: backward(forward(x))/n == x as the reference's tests require."""
    for key in GOLD.files:
        if key.startswith("r_f64_in_"):
            n = int(key.rsplit("_", 1)[1])
            assert oracle.max_row_rel_l2(GOLD[f"r_f64_packed_bwd_{n}"], GOLD[key]) <= 2e-15, n


def test_ref_roundtrip_all_lengths(ref):
    """tests/test_fft.nim:29-49 (test_real) and ffttest.c:69-93 (complex) on the compiled
    reference: every length 1..8191, rel-L2 <= 2e-15."""
    rng = np.random.default_rng(7)
    odata = rng.uniform(-0.5, 0.5, 8192)
    odata[0] = 0.340188
    worst = 0.0
    for n in range(1, 8192):
        d = odata[:n].copy().reshape(1, n)
        ref.rfft_rows(d, True, 1.0)
        ref.rfft_rows(d, False, 1.0 / n)
        worst = max(worst, oracle.rel_l2(d[0], odata[:n]))
    assert worst <= 2e-15
    cdata = rng.uniform(-0.5, 0.5, 4096) + 1j * rng.uniform(-0.5, 0.5, 4096)
    for n in list(range(1, 600)) + [1000, 2048, 3888, 4093, 4096]:
        d = cdata[:n].copy().reshape(1, n)
        ref.cfft_rows(d, True, 1.0)
        ref.cfft_rows(d, False, 1.0 / n)
        assert oracle.rel_l2(d[0], cdata[:n]) <= 2e-15, n


def test_port_vs_ref_sweep(port, ref):
    """Restatement vs compiled reference, forward results, real and complex, across every
    factorisation class and the fftpack<->Bluestein switch (first at 89 complex / 191 real)."""
    rng = np.random.default_rng(11)
    lengths = sorted(set(list(range(1, 260)) + list(range(260, 8192, 97)) +
                         [1000, 1024, 3888, 4096, 4099, 4126, 8191]))
    worst_c = worst_r = 0.0
    for n in lengths:
        x = (rng.uniform(-0.5, 0.5, (1, n)) + 1j * rng.uniform(-0.5, 0.5, (1, n)))
        a = ref.cfft_rows(x.copy(), True, 1.0)
        b = port.cfft_rows(x.copy(), True, 1.0)
        worst_c = max(worst_c, oracle.rel_l2(b, a))
        r = rng.uniform(-0.5, 0.5, (1, n))
        a = ref.rfft_rows(r.copy(), True, 1.0)
        b = port.rfft_rows(r.copy(), True, 1.0)
        worst_r = max(worst_r, oracle.rel_l2(b, a))
        a2 = ref.rfft_rows(a.copy(), False, 1.0 / n)
        b2 = port.rfft_rows(a.copy(), False, 1.0 / n)
        worst_r = max(worst_r, oracle.rel_l2(b2, a2))
    assert worst_c <= 5e-15 and worst_r <= 5e-15, (worst_c, worst_r)


def test_port_plan_decisions_match_survey(port):
    """SURVEY A.2 (probe of the compiled reference): factor lists and Bluestein switch."""
    assert port.factors(1024) == [4, 4, 4, 4, 4]
    assert port.factors(8192) == [2, 4, 4, 4, 4, 4, 4]
    assert port.factors(1000) == [2, 4, 5, 5, 5]
    assert port.factors(3888) == [4, 4, 3, 3, 3, 3, 3]
    assert port.uses_bluestein(4099) and port.good_size(2 * 4099 - 1) == 8232
    assert port.good_size(2 * 4126 - 1) == 8316
    assert min(n for n in range(1, 300) if port.uses_bluestein(n, False)) == 89
    assert min(n for n in range(1, 300) if port.uses_bluestein(n, True)) == 191
    assert sum(port.uses_bluestein(n, False) for n in range(1, 8192)) == 4294
    assert sum(port.uses_bluestein(n, True) for n in range(1, 8192)) == 3189


def test_ref_cpp_matches_c_engine(ref):
    """The reference's two engines agree (2-7e-16)."""
    rng = np.random.default_rng(3)
    for n in (1000, 1024, 3888, 4096, 4099):
        x = rng.uniform(-0.5, 0.5, (2, n)) + 1j * rng.uniform(-0.5, 0.5, (2, n))
        a = ref.cfft_rows(x.copy(), True, 1.0)
        b = ref.c2c(x, [1], True, 1.0)
        assert oracle.max_row_rel_l2(b, a) <= 2e-15


def test_length_one_scaling_quirk(port, ref):
    """The C engine returns before scaling for length-1 plans (pocketfft.c:874,1703,1739); the C++
    engine scales (pocketfft_hdronly.h:1309,2114).  Both behaviours are mirrored downstream."""
    c = np.array([[2.0 + 1.0j]])
    for e in (ref, port):
        assert e.cfft_rows(c.copy(), True, 0.37)[0, 0] == 2.0 + 1.0j
        assert e.rfft_rows(np.array([[3.0]]), True, 0.37)[0, 0] == 3.0
        assert np.allclose(e.c2c(c, [1], True, 0.5), c * 0.5)


def test_dct_dst_conventions(ref):
    """pocketfft::dct / dst (pocketfft_hdronly.h:3284-3318) against the O(N^2) definitions
    (FFTW REDFT/RODFT kinds) incl. the `ortho` rules of README_pocketfft.md:220-241, and the one
    value the reference prints: DCT-II of [4,3,5,10] (cpp_pocketfft/pocketfft.nim:322-339)."""
    rng = np.random.default_rng(17)
    for n in (2, 3, 4, 5, 8, 9, 16, 31, 64):
        x = rng.uniform(-0.5, 0.5, (2, n))
        for cosine in (True, False):
            for t in (1, 2, 3, 4):
                for ortho in (False, True):
                    a = ref.r2r(cosine, t, x, [1], 0.7, ortho)
                    assert oracle.rel_l2(oracle.r2r_direct(cosine, t, x, 0.7, ortho), a) <= 1e-13, (n, cosine, t, ortho)
    d = ref.r2r(True, 2, np.array([[4.0, 3.0, 5.0, 10.0]]), [1], 1.0, False)[0]
    np.testing.assert_allclose(d, oracle.r2r_direct(True, 2, np.array([4.0, 3.0, 5.0, 10.0])), rtol=1e-14)
    assert abs(d[0] - 44.0) < 1e-13   # 2 * sum(x)


def test_fftpack_and_hartley_conventions(ref):
    """Pin the numpy restatements of r2r_fftpack / r2r_separable_hartley / r2r_genuine_hartley against the
    compiled reference (pocketfft_hdronly.h:3392-3445), including the mixed real2hermitian/forward cases
    that the vendored header computes in its own way."""
    rng = np.random.default_rng(17)
    for shape, axes in (((7,), [0]), ((8,), [0]), ((3, 10), [1]), ((6, 9), [0, 1]), ((4, 5, 6), [2, 0]), ((2, 191), [1])):
        a = rng.standard_normal(shape)
        for r2h in (True, False):
            for fwd in (True, False):
                got = oracle.fftpack_numpy(a, axes, r2h, fwd, 0.5)
                want = ref.r2r_real("fftpack", a, axes, r2h, fwd, 0.5)
                assert oracle.rel_l2(got, want) < 1e-14, (shape, axes, r2h, fwd)
        for which, genuine in (("separable_hartley", False), ("genuine_hartley", True)):
            got = oracle.hartley_numpy(a, axes, genuine, 2.0)
            want = ref.r2r_real(which, a, axes, fct=2.0)
            assert oracle.rel_l2(got, want) < 1e-14, (shape, axes, which)
    # a Hartley transform is its own inverse up to 1/N
    a = rng.standard_normal((5, 12))
    back = ref.r2r_real("genuine_hartley", ref.r2r_real("genuine_hartley", a, [0, 1]), [0, 1], fct=1.0 / 60)
    assert oracle.rel_l2(back, a) < 1e-14
