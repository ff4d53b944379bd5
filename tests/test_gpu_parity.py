"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI via the
Python host mirror, against the oracle (compiled reference when its prebuilt .so travelled, else
the port) and the committed golden fixtures.

Tolerance (north_star): rel-L2 <= 1e-12*log2(N) for float64, 1e-5*log2(N) for float32; the
reference's own round-trip bar of 2e-15 (tests/test_fft.nim:32) is additionally asserted, scaled
by log2(N), where the reference tests it.
"""
import os

import numpy as np
import pytest

from oracle import nim_helpers as nh
from oracle import oracle

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "pocketfft_golden.npz"))


def tol(n, dtype=np.float64):
    base = 1e-12 if np.dtype(dtype) in (np.dtype(np.float64), np.dtype(np.complex128)) else 1e-5
    return base * max(1.0, np.log2(max(n, 2)))


def rnd(rng, shape, dtype):
    if np.dtype(dtype).kind == "c":
        return (rng.uniform(-0.5, 0.5, shape) + 1j * rng.uniform(-0.5, 0.5, shape)).astype(dtype)
    return rng.uniform(-0.5, 0.5, shape).astype(dtype)


@pytest.fixture(scope="module")
def ib():
    import torch
    assert torch.cuda.is_available()
    import impulse_b200
    return impulse_b200


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


def apply_nd(ib, kind, a_in, out, axes, forward=True, fct=1.0):
    ib.FFTDesc.init(axes=axes, forward=forward, scalingFactor=fct).apply(ib.DataDesc.init(out), ib.DataDesc.init(a_in))
    return out


# ---- reference known answers ---------------------------------------------------------------
def test_t2_known_answers(ib, torch_mod):
    """tests/test_fft2.nim:5-15, exact comparisons, host arrays and device tensors."""
    expected = np.array([6, -2 + 2j, -2, -2 - 2j], dtype=np.complex128)
    x = np.arange(4.0)
    for conv in (lambda a: a, lambda a: torch_mod.from_numpy(np.ascontiguousarray(a)).cuda()):
        back = (lambda r: r.cpu().numpy() if hasattr(r, "cpu") else r)
        assert np.array_equal(back(ib.fft(conv(x.astype(np.complex128)))), expected)
        assert np.array_equal(back(ib.fft(conv(x))), expected)
        assert np.array_equal(back(ib.fft(conv(x), forward=True)), expected)
        rt = back(ib.ifft(ib.fft(conv(x))))
        assert np.array_equal(rt.real, x) and np.array_equal(rt.imag, np.zeros(4))
        assert np.array_equal(back(ib.fft(conv(x), normalize=ib.nkForward)), expected / 4.0)
        assert np.array_equal(back(ib.fft(conv(x), normalize=ib.nkOrtho)), expected / 2.0)
        assert np.array_equal(back(ib.ifft(conv(x))), back(ib.fft(conv(x), forward=False)))


def test_readme_vectors(ib):
    din = np.array([1.0, 2.0, 1.0, -1.0, 1.5])
    packed = np.array([4.5, 2.081559480312316, -1.651098762732523, -1.831559480312316, 1.608220406444071])
    full = np.array([4.5 + 0j, 2.081559480312316 - 1.651098762732523j, -1.831559480312316 + 1.608220406444071j,
                     -1.831559480312316 - 1.608220406444071j, 2.081559480312316 + 1.651098762732523j])
    np.testing.assert_allclose(ib.rfft_packed(din), packed, rtol=0, atol=4e-15)   # README.md:41
    np.testing.assert_allclose(ib.fft(din), full, rtol=0, atol=4e-15)             # README.md:68
    np.testing.assert_allclose(ib.fft(din.astype(np.complex128)), full, rtol=0, atol=4e-15)
    np.testing.assert_allclose(ib.fft(ib.fft(din), forward=False).real, din, atol=1e-10)
    # C++ example (README.md:75-99): r2c into a 5-slot buffer leaves the last two slots untouched
    out = np.zeros(5, dtype=np.complex128)
    ib.FFTDesc.init(axes=[0], forward=True).apply(ib.DataDesc.init(out, [3]), ib.DataDesc.init(din, [5]))
    np.testing.assert_allclose(out[:3], full[:3], rtol=0, atol=4e-15)
    assert out[3] == 0 and out[4] == 0


def test_pocketfft_c_symbols(ib):
    """The ten drop-in symbols (include/pocketfft.h) exactly as tests/test_fft.nim:39-42 calls them."""
    import ctypes as C
    from impulse_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(2)
    for n in (1, 2, 5, 64, 1000, 4099):
        d = rng.uniform(-0.5, 0.5, n)
        o = d.copy()
        plan = L.make_rfft_plan(n)
        assert plan and L.rfft_length(plan) == n
        assert L.rfft_forward(plan, d.ctypes.data, 1.0) == 0
        want = oracle.load().rfft_rows(o.copy().reshape(1, n), True, 1.0)[0]
        assert oracle.rel_l2(d, want) <= tol(n)
        assert L.rfft_backward(plan, d.ctypes.data, 1.0 / n) == 0
        L.destroy_rfft_plan(plan)
        assert oracle.rel_l2(d, o) <= 2e-15 * max(1, np.log2(n))
        c = rnd(rng, n, np.complex128)
        oc = c.copy()
        plan = L.make_cfft_plan(n)
        assert plan and L.cfft_length(plan) == n
        assert L.cfft_forward(plan, c.ctypes.data, 1.0) == 0
        want = oracle.load().cfft_rows(oc.copy().reshape(1, n), True, 1.0)[0]
        assert oracle.rel_l2(c, want) <= tol(n)
        assert L.cfft_backward(plan, c.ctypes.data, 1.0 / n) == 0
        L.destroy_cfft_plan(plan)
        assert oracle.rel_l2(c, oc) <= 2e-15 * max(1, np.log2(n))
    assert L.make_cfft_plan(0) is None and L.make_rfft_plan(0) is None


# ---- 1-D parity sweeps -----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_c2c_lengths(ib, torch_mod, checker, dtype):
    rng = np.random.default_rng(5)
    lengths = list(range(1, 131)) + [144, 169, 187, 191, 192, 200, 243, 256, 289, 343, 360, 361, 400, 500, 512, 529, 576, 625, 900,
                                     729, 841, 961, 1000, 1024, 1331, 2048, 2187, 3125, 3888, 4096, 4099, 6561, 7000]
    if dtype == np.complex64:
        lengths = lengths[::3]
    for n in lengths:
        x = rnd(rng, (5, n), dtype)
        xd = torch_mod.from_numpy(x).cuda()
        for fwd in (True, False):
            fct = 1.0 if fwd else 1.0 / n
            want = checker.c2c(x, [1], fwd, fct)
            got = apply_nd(ib, "c2c", xd, torch_mod.empty_like(xd), [1], fwd, fct).cpu().numpy()
            assert oracle.max_row_rel_l2(got, want) <= tol(n, dtype), (n, fwd)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tiny_rows_many(ib, torch_mod, checker, dtype):
    """rows of 4 / 8 complex and 8 / 16 real points on the warp kernels (16 rows per warp): more rows than one sweep of
    the grid, a row count that leaves the last warp ragged, both directions"""
    rng = np.random.default_rng(21)
    cdt = np.complex128 if dtype == np.float64 else np.complex64
    used = set()
    for n in (4, 8):
        x = rnd(rng, (100003, n), cdt)
        xd = torch_mod.from_numpy(x).cuda()
        for fwd in (True, False):
            got = apply_nd(ib, "c2c", xd, torch_mod.empty_like(xd), [1], fwd, 0.5).cpu().numpy()
            used.add(ib.last_kernel())
            assert oracle.max_row_rel_l2(got, checker.c2c(x, [1], fwd, 0.5)) <= tol(n, dtype), (n, fwd)
    for n in (8, 16):
        x = rnd(rng, (70001, n), dtype)
        xd = torch_mod.from_numpy(x).cuda()
        for fwd in (True, False):
            want = checker.r2c(x, [1], fwd, 0.5)
            sd = apply_nd(ib, "r2c", xd, torch_mod.zeros((70001, n // 2 + 1), dtype=getattr(torch_mod, np.dtype(cdt).name), device="cuda"), [1], fwd, 0.5)
            used.add(ib.last_kernel())
            assert oracle.max_row_rel_l2(sd.cpu().numpy(), want) <= tol(n, dtype), (n, fwd)
        spec = checker.r2c(x, [1], True, 1.0).astype(cdt)
        back = apply_nd(ib, "c2r", torch_mod.from_numpy(spec).cuda(), torch_mod.empty_like(xd), [1], False, 1.0 / n).cpu().numpy()
        used.add(ib.last_kernel())
        assert oracle.max_row_rel_l2(back, x) <= 4 * tol(n, dtype), n
    assert all(k.startswith("fast2") for k in used), used


def test_real_roundtrip_all_lengths(ib, torch_mod, checker):
    """tests/test_fft.nim:29-109 — every length 1..8191, forward then backward(1/N) through the
    packed in-place layout recovers the input; forward parity vs the oracle on every length too.
    Every length must be supported: Bluestein sizes beyond one CTA's shared memory (odd N > 7204 with
    a large prime factor) take the multi-launch path."""
    rng = np.random.default_rng(7)
    odata = rng.uniform(-0.5, 0.5, 8192)
    odata[0] = 0.340188
    od = torch_mod.from_numpy(odata).cuda()
    unsupported = []
    worst_rt = worst_fw = 0.0
    for n in range(1, 8192):
        d = od[:n].clone()
        try:
            ib.fft_inplace(d, forward=True)
        except ib.FFTError as e:
            assert e.code == -3, (n, str(e))
            unsupported.append(n)
            continue
        if n % 7 == 0 or n < 300:
            want = checker.rfft_rows(odata[:n].copy().reshape(1, n), True, 1.0)[0]
            worst_fw = max(worst_fw, oracle.rel_l2(d.cpu().numpy(), want) / max(1.0, np.log2(n)))
        ib.fft_inplace(d, forward=False)
        worst_rt = max(worst_rt, oracle.rel_l2(d.cpu().numpy(), odata[:n]) / max(1.0, np.log2(n)))
    assert worst_fw <= 1e-12, worst_fw
    assert worst_rt <= 2e-15, worst_rt
    assert not unsupported, unsupported[:10]
    print(f"unsupported lengths: {len(unsupported)}; worst forward {worst_fw:.2e}, round trip {worst_rt:.2e} (per log2 N)")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_r2c_c2r_hermitian(ib, torch_mod, checker, dtype):
    rng = np.random.default_rng(6)
    cdt = np.complex128 if dtype == np.float64 else np.complex64
    lengths = list(range(1, 70)) + [74, 89, 97, 100, 121, 128, 191, 192, 243, 250, 382, 500, 1000, 1024, 2000, 3888,
                                    4096, 4099, 4126, 3072, 4000, 4374, 6000, 8000, 13122]   # (the last six: real rows on the
    # three-pass shapes 1536 / 2000 / 2187 / 3000 / 4000 / 6561 through the shared-memory Hermitian twiddle)
    for n in lengths:
        x = rnd(rng, (3, n), dtype)
        xd = torch_mod.from_numpy(x).cuda()
        for fwd in (True, False):
            want = checker.r2c(x, [1], fwd, 0.5)
            got = apply_nd(ib, "r2c", xd, torch_mod.zeros((3, n // 2 + 1), dtype=getattr(torch_mod, np.dtype(cdt).name),
                                                          device="cuda"), [1], fwd, 0.5).cpu().numpy()
            assert oracle.max_row_rel_l2(got, want) <= tol(n, dtype), (n, fwd)
        spec = checker.r2c(x, [1], True, 1.0)
        sd = torch_mod.from_numpy(spec).cuda()
        for fwd in (False, True):
            want = checker.c2r(spec, x.shape, [1], fwd, 1.0 / n)
            got = apply_nd(ib, "c2r", sd, torch_mod.zeros_like(xd), [1], fwd, 1.0 / n).cpu().numpy()
            assert oracle.max_row_rel_l2(got, want) <= tol(n, dtype), (n, fwd)


def test_highlevel_api_matches_nim_semantics(ib, checker):
    """fft/ifft/rfft/rfft_packed/fft_inplace against the restated Nim glue on top of the oracle."""
    api = nh.NimApi(checker)
    rng = np.random.default_rng(8)
    for n in (1, 2, 3, 4, 5, 8, 9, 30, 89, 191, 256, 1000):
        x = rng.uniform(-0.5, 0.5, n)
        c = rnd(rng, n, np.complex128)
        for kind, val in ((ib.nkBackward, np.inf), (ib.nkForward, np.inf), (ib.nkOrtho, np.inf), (ib.nkCustom, 0.37)):
            for fwd in (True, False):
                assert oracle.rel_l2(ib.fft(x, fwd, kind, val), api.fft(x, fwd, kind, val)) <= tol(n), (n, kind, fwd)
                assert oracle.rel_l2(ib.fft(c, fwd, kind, val), api.fft(c, fwd, kind, val)) <= tol(n)
                assert oracle.rel_l2(ib.rfft_packed(x, fwd, kind, val), api.rfft_packed(x, fwd, kind, val)) <= tol(n)
                assert oracle.rel_l2(ib.rfft(x, fwd, kind, val), api.rfft(x, fwd, kind, val)) <= tol(n)
        assert oracle.rel_l2(ib.ifft(c), api.ifft(c)) <= tol(n)
        d = x.copy()
        ib.fft_inplace(d)
        assert oracle.rel_l2(d, api.fft_inplace(x)) <= tol(n)
    with pytest.raises(ib.FFTError):
        ib.fft(np.zeros(0))
    # batched extension: last axis, leading axes batched
    xb = rng.uniform(-0.5, 0.5, (4, 3, 100))
    got = ib.fft(xb)
    for i in range(4):
        for j in range(3):
            assert oracle.rel_l2(got[i, j], api.fft(xb[i, j])) <= tol(100)


# ---- N-D / strided / golden --------------------------------------------------------------------
def test_nd_and_strided(ib, torch_mod, checker):
    rng = np.random.default_rng(9)
    a = rnd(rng, (12, 20), np.complex128)
    for axes in ([0, 1], [1, 0], [0], [1]):
        want = checker.c2c(a, axes, True, 0.5)
        assert oracle.rel_l2(apply_nd(ib, "c2c", a, np.empty_like(a), axes, True, 0.5), want) <= tol(20)
    b = a.copy()
    apply_nd(ib, "c2c", b, b, [0, 1], False, 1.0)
    assert oracle.rel_l2(b, checker.c2c(a, [0, 1], False, 1.0)) <= tol(20)
    c = rnd(rng, (5, 6, 14), np.complex64)
    v = c[::-1, :, ::2]
    assert oracle.rel_l2(apply_nd(ib, "c2c", v, np.empty(v.shape, np.complex64), [0, 2]), checker.c2c(v, [0, 2])) <= tol(7, np.float32)
    t = rnd(rng, (16, 9), np.complex128).T
    assert oracle.rel_l2(apply_nd(ib, "c2c", t, np.empty(t.shape, np.complex128), [0, 1]), checker.c2c(t, [0, 1])) <= tol(16)
    e = rnd(rng, (2, 3, 2, 4, 6), np.complex128)[:, :, ::-1]
    assert oracle.rel_l2(apply_nd(ib, "c2c", e, np.empty(e.shape, np.complex128), [1, 3]), checker.c2c(e, [1, 3])) <= tol(6)
    # strided output with gaps must leave the gaps alone (host staging path)
    big = np.full((6, 20), 7.0 + 7.0j)
    view = big[:, ::2]
    src = rnd(rng, (6, 10), np.complex128)
    apply_nd(ib, "c2c", src, view, [1])
    assert oracle.rel_l2(view, checker.c2c(src, [1])) <= tol(10)
    assert np.all(big[:, 1::2] == 7.0 + 7.0j)
    for shape, axes in (((8, 12), [0, 1]), ((5, 9), [0, 1]), ((5, 9), [1, 0]), ((4, 6, 10), [0, 2]), ((3, 7, 5), [0, 1, 2])):
        r = rnd(rng, shape, np.float64)
        want = checker.r2c(r, axes, True, 1.0)
        got = apply_nd(ib, "r2c", r, np.zeros(want.shape, np.complex128), axes)
        assert oracle.rel_l2(got, want) <= tol(12), (shape, axes)
        back = apply_nd(ib, "c2r", want, np.zeros(shape), axes, False, 1.0 / np.prod([shape[k] for k in axes]))
        assert oracle.rel_l2(back, r) <= tol(12), (shape, axes)
    # device-resident 2-D of a size where the strided axis spans many tiles
    m = rnd(rng, (300, 520), np.complex128)
    md = torch_mod.from_numpy(m).cuda()
    got = apply_nd(ib, "c2c", md, torch_mod.empty_like(md), [0, 1]).cpu().numpy()
    assert oracle.rel_l2(got, checker.c2c(m, [0, 1])) <= tol(520)
    r32 = rnd(rng, (3, 64, 96), np.float32)
    rd = torch_mod.from_numpy(r32).cuda()
    spec = apply_nd(ib, "r2c", rd, torch_mod.zeros((3, 64, 49), dtype=torch_mod.complex64, device="cuda"), [1, 2])
    assert oracle.rel_l2(spec.cpu().numpy(), checker.r2c(r32, [1, 2])) <= tol(96, np.float32)
    back = apply_nd(ib, "c2r", spec, torch_mod.zeros_like(rd), [1, 2], False, 1.0 / (64 * 96)).cpu().numpy()
    assert oracle.rel_l2(back, r32) <= tol(96, np.float32)


def test_golden_fixtures(ib):
    for key in GOLD.files:
        if key.startswith("c2c_f64_in_"):
            n = int(key.rsplit("_", 1)[1])
            x = GOLD[key]
            assert oracle.max_row_rel_l2(ib.fft(x), GOLD[f"c2c_f64_fwd_{n}"]) <= tol(n), n
            assert oracle.max_row_rel_l2(ib.ifft(x), GOLD[f"c2c_f64_bwd_{n}"]) <= tol(n), n
        if key.startswith("r_f64_in_"):
            n = int(key.rsplit("_", 1)[1])
            p = ib.rfft_packed(GOLD[key])
            assert oracle.max_row_rel_l2(p, GOLD[f"r_f64_packed_fwd_{n}"]) <= tol(n), n
            b = ib.rfft_packed(GOLD[f"r_f64_packed_fwd_{n}"], forward=False)
            assert oracle.max_row_rel_l2(b, GOLD[f"r_f64_packed_bwd_{n}"]) <= tol(n), n
    a = GOLD["nd_c2c_f64_in"]
    assert oracle.rel_l2(apply_nd(ib, "c2c", a, np.empty_like(a), [0, 1]), GOLD["nd_c2c_f64_ax01"]) <= tol(10)
    assert oracle.rel_l2(apply_nd(ib, "c2c", a, np.empty_like(a), [0], False, 0.25), GOLD["nd_c2c_f64_ax0_bwd"]) <= tol(10)
    b = GOLD["nd_c2c_f32_in"]
    assert oracle.rel_l2(apply_nd(ib, "c2c", b, np.empty_like(b), [1, 2]), GOLD["nd_c2c_f32_ax12"]) <= tol(9, np.float32)
    r = GOLD["nd_r2c_f32_in"]
    assert oracle.rel_l2(apply_nd(ib, "r2c", r, np.zeros((8, 7), np.complex64), [0, 1]), GOLD["nd_r2c_f32_ax01"]) <= tol(12, np.float32)
    assert oracle.rel_l2(apply_nd(ib, "r2c", r, np.zeros((8, 7), np.complex64), [1], False), GOLD["nd_r2c_f32_ax1_bwd"]) <= tol(12, np.float32)
    s = GOLD["nd_r2c_f64_ax01"]
    assert oracle.rel_l2(apply_nd(ib, "c2r", s, np.zeros((5, 9)), [0, 1], False, 1.0 / 45), GOLD["nd_c2r_f64_ax01"]) <= tol(9)
    s1 = GOLD["nd_r2c_f64_ax0"]
    assert oracle.rel_l2(apply_nd(ib, "c2r", s1, np.zeros((5, 9)), [0], False, 0.2), GOLD["nd_c2r_f64_ax0"]) <= tol(9)


def test_errors(ib):
    a = np.zeros((4, 4), np.complex128)
    for axes in ([2], [0, 0]):
        with pytest.raises(ib.FFTError):
            apply_nd(ib, "c2c", a, a.copy(), axes)
    z = np.zeros((0, 4), np.complex128)
    apply_nd(ib, "c2c", z, z.copy(), [1])  # empty array: no-op (hdronly.h:3277)
    with pytest.raises(TypeError):
        apply_nd(ib, "r2r", np.zeros(4), np.zeros(4), [0])
    with pytest.raises(ib.FFTError):  # r2c in place (Hermitian layout): rows would overlap their neighbours' output
        buf = np.zeros((4, 10))
        ib.FFTDesc.init(axes=[1], forward=True).apply(ib.DataDesc.init(buf.view(np.complex128)), ib.DataDesc.init(buf[:, :8]))
    with pytest.raises(ib.FFTError):  # in place with different strides (hdronly.h:455)
        m = np.zeros((4, 4), np.complex128)
        ib.FFTDesc.init(axes=[0], forward=True).apply(ib.DataDesc.init(m.T), ib.DataDesc.init(m))


# ---- BASELINE configs at full size: properties that need no full-size oracle ------------------
def _all_threads(checker):
    return max(1, checker.hardware_threads())


def test_config2_full_size_properties(ib, torch_mod, checker):
    """65536 x 1024 complex128: EVERY row against the oracle (all host threads: the dynamic row scheduler
    could skip or repeat a row that a sampled comparison never sees), round trip on all rows, Parseval."""
    g = torch_mod.Generator(device="cuda").manual_seed(1234)
    x = torch_mod.rand((65536, 1024, 2), generator=g, device="cuda", dtype=torch_mod.float64) - 0.5
    x = torch_mod.view_as_complex(x)
    y = torch_mod.full_like(x, float("nan"))          # a row that is never written stays NaN
    apply_nd(ib, "c2c", x, y, [1])
    want = checker.cfft_rows(x.cpu().numpy().copy(), True, 1.0, nthreads=_all_threads(checker))
    assert oracle.max_row_rel_l2(y.cpu().numpy(), want) <= tol(1024)
    del want
    back = ib.ifft(y)
    num = torch_mod.linalg.vector_norm(back - x, dim=1)
    den = torch_mod.linalg.vector_norm(x, dim=1)
    assert float((num / den).max()) <= 2e-15 * 10
    # Parseval: sum|X|^2 = N sum|x|^2 per row
    e_in = (x.abs() ** 2).sum(dim=1)
    e_out = (y.abs() ** 2).sum(dim=1) / 1024
    assert float(((e_in - e_out).abs() / e_in).max()) <= 1e-13


def test_config1_and_3_shapes(ib, torch_mod, checker):
    """r2c 1024x4096 (config 1) and r2c/c2r at 16384 x {1000, 3888, 4099} (config 3): every row of both directions
    against the oracle, and the r2c -> c2r round trip."""
    g = torch_mod.Generator(device="cuda").manual_seed(1234)
    nt = _all_threads(checker)
    for rows, n in ((1024, 4096), (16384, 1000), (16384, 3888), (16384, 4099)):
        x = torch_mod.rand((rows, n), generator=g, device="cuda", dtype=torch_mod.float64) - 0.5
        spec = torch_mod.full((rows, n // 2 + 1), float("nan"), dtype=torch_mod.complex128, device="cuda")
        apply_nd(ib, "r2c", x, spec, [1])
        xh = x.cpu().numpy()
        want = checker.r2c(xh, [1], True, 1.0, nthreads=nt)
        assert oracle.max_row_rel_l2(spec.cpu().numpy(), want) <= tol(n), n
        back = torch_mod.full_like(x, float("nan"))
        apply_nd(ib, "c2r", spec, back, [1], False, 1.0 / n)
        wantb = checker.c2r(want, xh.shape, [1], False, 1.0 / n, nthreads=nt)   # the oracle's own inverse of its spectrum
        assert oracle.max_row_rel_l2(back.cpu().numpy(), wantb) <= tol(n), n
        num = torch_mod.linalg.vector_norm(back - x, dim=1)
        den = torch_mod.linalg.vector_norm(x, dim=1)
        assert float((num / den).max()) <= 2e-15 * np.log2(n), n


def test_four_step_long_and_strided_lines(ib, torch_mod, checker):
    """Lines that do not fit one CTA (contiguous 32768 points) or cannot be tiled 8 columns wide
    (strided 4096 / 8192 points) run as two launches with a twiddle in between."""
    rng = np.random.default_rng(12)
    x = rnd(rng, (3, 32768), np.complex128)
    xd = torch_mod.from_numpy(x).cuda()
    for fwd in (True, False):
        got = apply_nd(ib, "c2c", xd, torch_mod.empty_like(xd), [1], fwd, 0.5).cpu().numpy()
        assert oracle.max_row_rel_l2(got, checker.c2c(x, [1], fwd, 0.5)) <= tol(32768), fwd
    y = rnd(rng, (4096, 40), np.complex128)
    yd = torch_mod.from_numpy(y).cuda()
    got = apply_nd(ib, "c2c", yd, torch_mod.empty_like(yd), [0]).cpu().numpy()
    assert oracle.rel_l2(got, checker.c2c(y, [0])) <= tol(4096)
    apply_nd(ib, "c2c", yd, yd, [0], False, 1.0 / 4096)   # in place
    assert oracle.rel_l2(yd.cpu().numpy(), checker.c2c(y, [0], False, 1.0 / 4096)) <= tol(4096)
    # fft2 (FFTDesc axes=[0,1], the path of BASELINE config 4) at 2048^2 against the oracle
    m = rnd(rng, (2048, 2048), np.complex128)
    md = torch_mod.from_numpy(m).cuda()
    got = apply_nd(ib, "c2c", md, torch_mod.empty_like(md), [0, 1]).cpu().numpy()
    assert oracle.rel_l2(got, checker.c2c(m, [0, 1], nthreads=0)) <= tol(2048)
    f32 = rnd(rng, (4096, 24), np.complex64)
    fd = torch_mod.from_numpy(f32).cuda()
    got = apply_nd(ib, "c2c", fd, torch_mod.empty_like(fd), [0]).cpu().numpy()
    assert oracle.rel_l2(got, checker.c2c(f32, [0])) <= tol(4096, np.float32)


def test_config4_full_size_properties(ib, torch_mod, checker):
    """fft2 8192 x 8192 complex128 on one GPU: the WHOLE array against the oracle (all host threads), then round
    trip, Parseval and the DC bin."""
    g = torch_mod.Generator(device="cuda").manual_seed(1234)
    x = torch_mod.view_as_complex(torch_mod.rand((8192, 8192, 2), generator=g, device="cuda", dtype=torch_mod.float64) - 0.5)
    y = torch_mod.full_like(x, float("nan"))
    apply_nd(ib, "c2c", x, y, [0, 1])
    want = checker.c2c(x.cpu().numpy(), [0, 1], True, 1.0, nthreads=0)
    got = y.cpu().numpy()
    assert oracle.rel_l2(got, want) <= tol(8192)
    assert oracle.max_row_rel_l2(got, want) <= tol(8192)
    del want, got
    e_in = float((x.abs() ** 2).sum())
    e_out = float((y.abs() ** 2).sum()) / (8192.0 * 8192.0)
    assert abs(e_in - e_out) / e_in <= 1e-13
    assert abs(complex(y[0, 0]) - complex(x.sum())) / abs(complex(x.sum())) <= 1e-11
    apply_nd(ib, "c2c", y, y, [0, 1], False, 1.0 / (8192.0 * 8192.0))
    err = float(torch_mod.linalg.vector_norm(y - x) / torch_mod.linalg.vector_norm(x))
    assert err <= 2e-15 * 26, err


def test_fft_filter2d(ib, torch_mod, checker):
    """BASELINE config 5 path (r2c -> multiply -> c2r, float32) at a reduced size, against the same
    composition on the oracle and against a direct circular convolution."""
    from impulse_b200.filter import FFTFilter2D
    rng = np.random.default_rng(15)
    for (b, h, w, kh, kw, dt, rt) in ((3, 64, 96, 5, 7, np.float32, 2e-5), (2, 40, 36, 31, 31, np.float64, 1e-12),
                                      (2, 512, 512, 31, 31, np.float32, 2e-5)):
        img = rng.uniform(0, 1, (b, h, w)).astype(dt)
        ker = rng.uniform(0, 1, (kh, kw)).astype(dt)
        ker /= ker.sum()
        f = FFTFilter2D(torch_mod.from_numpy(ker).cuda(), h, w)
        got = f.apply(torch_mod.from_numpy(img).cuda()).cpu().numpy()
        pad = np.zeros((h, w), dt)
        ii = (np.arange(kh) - kh // 2) % h
        jj = (np.arange(kw) - kw // 2) % w
        pad[np.ix_(ii, jj)] = ker
        kspec = checker.r2c(pad, [0, 1], True, 1.0)
        spec = checker.r2c(img, [1, 2], True, 1.0) * kspec[None]
        want = checker.c2r(np.ascontiguousarray(spec), img.shape, [1, 2], False, 1.0 / (h * w))
        assert oracle.rel_l2(got, want) <= rt * np.log2(max(h, w)), (h, w)
        if h <= 64:  # direct circular convolution, one image
            direct = np.zeros((h, w))
            for i in range(kh):
                for j in range(kw):
                    direct += ker[i, j] * np.roll(np.roll(img[0].astype(np.float64), i - kh // 2, axis=0), j - kw // 2, axis=1)
            assert oracle.rel_l2(got[0], direct) <= rt * 10
        # mean preserved by a unit-sum kernel (the image_filters benchmark's sanity metric)
        assert abs(got.mean() - img.mean()) <= 1e-4


def test_register_kernels_all_kinds(ib, torch_mod, checker):
    """The specialised register kernels (two-pass 16...1024, three-pass 2048/4096/8192) in every
    kind they serve: c2c both directions, r2c/c2r with both `forward` flags, fp64 and fp32, odd batch
    sizes (dynamic row claiming with a ragged tail), plus the generic engine on the same inputs
    (IMPULSE_FFT_NO_FAST) as a second opinion."""
    rng = np.random.default_rng(21)
    used = set()
    for dt, cdt in ((np.float64, np.complex128), (np.float32, np.complex64)):
        more = (1536, 2000, 4000, 2187, 3000, 6561) if os.environ.get("IMPULSE_FFT_MORE_SHAPES", "1") == "1" else ()
        for n in (16, 32, 64, 100, 128, 243, 256, 500, 512, 625, 1000, 1024, 1944, 2048, 4096, 8192) + more:
            for rows in (1, 37, 301) + ((5000,) if n <= 128 else ()):
                x = rnd(rng, (rows, n), cdt)
                xd = torch_mod.from_numpy(x).cuda()
                for fwd in (True, False):
                    got = apply_nd(ib, "c2c", xd, torch_mod.empty_like(xd), [1], fwd, 0.7).cpu().numpy()
                    used.add(ib.last_kernel())
                    assert oracle.max_row_rel_l2(got, checker.c2c(x, [1], fwd, 0.7)) <= tol(n, dt), (n, rows, fwd, dt)
        for n in (32, 64, 128, 256, 512, 1024, 2048, 1000, 3888, 4096, 8192, 16384) + tuple(2 * m for m in more):
            for rows in (1, 53) + ((1001,) if n <= 2048 else ()):
                r = rnd(rng, (rows, n), dt)
                rd = torch_mod.from_numpy(r).cuda()
                for fwd in (True, False):
                    spec = apply_nd(ib, "r2c", rd, torch_mod.empty((rows, n // 2 + 1), dtype=getattr(torch_mod, np.dtype(cdt).name),
                                                                   device="cuda"), [1], fwd, 1.0)
                    used.add(ib.last_kernel())
                    want = checker.r2c(r, [1], fwd, 1.0)
                    assert oracle.max_row_rel_l2(spec.cpu().numpy(), want) <= tol(n, dt), (n, rows, fwd, dt)
                sp = checker.r2c(r, [1], True, 1.0)
                sd = torch_mod.from_numpy(sp).cuda()
                for fwd in (False, True):
                    back = apply_nd(ib, "c2r", sd, torch_mod.empty_like(rd), [1], fwd, 1.0 / n).cpu().numpy()
                    used.add(ib.last_kernel())
                    assert oracle.max_row_rel_l2(back, checker.c2r(sp, r.shape, [1], fwd, 1.0 / n)) <= tol(n, dt), (n, rows, fwd, dt)
                # pocketfft's c2r never reads the imaginary parts of bins 0 and N/2 (pocketfft_hdronly.h:3196-3210)
                sj = sp.copy()
                sj[:, 0] += 0.375j
                sj[:, -1] -= 0.25j
                back = apply_nd(ib, "c2r", torch_mod.from_numpy(sj).cuda(), torch_mod.empty_like(rd), [1], False, 1.0 / n).cpu().numpy()
                assert oracle.max_row_rel_l2(back, checker.c2r(sp, r.shape, [1], False, 1.0 / n)) <= tol(n, dt), (n, rows, "junk imag", dt)
    print(sorted(used))
    assert any(k.startswith("fast3_kernel") for k in used) and any(k.startswith("fast2") for k in used)
    # the in-register pair variants (r2c post-twiddle in pass 3, c2r pre-twiddle in pass 1) are the default
    if os.environ.get("IMPULSE_FFT_R2C_PAIR", "1") != "0" and os.environ.get("IMPULSE_FFT_C2R_PAIR", "1") != "0":
        pair = sorted(k for k in used if "+pair" in k)
        assert any("10,10,5" in k for k in pair) and any("18,18,6" in k for k in pair) and any("8,16,16" in k for k in pair), pair


def test_column_kernels(ib, torch_mod, checker):
    """Strided-axis register kernels (colfast2) directly (axis 0 of 32..512 points over >= 8/16 adjacent
    columns, both directions, fp64/fp32) and inside the four-step split (1024..16384-point columns)."""
    rng = np.random.default_rng(31)
    used = set()
    for dt, cdt, cols in ((np.float64, np.complex128, 24), (np.float32, np.complex64, 48), (np.float64, np.complex128, 29),
                          (np.float32, np.complex64, 2049)):
        for n in ((32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384) if cols < 100 else (64, 4096)):
            a = rnd(rng, (n, cols), cdt)
            ad = torch_mod.from_numpy(a).cuda()
            for fwd in (True, False):
                got = apply_nd(ib, "c2c", ad, torch_mod.empty_like(ad), [0], fwd, 0.5).cpu().numpy()
                used.add(ib.last_kernel())
                assert oracle.rel_l2(got, checker.c2c(a, [0], fwd, 0.5)) <= tol(n, dt), (n, fwd, dt)
        b = rnd(rng, (3, 128, 5, cols), cdt)       # two more batch dims around the strided axis
        bd = torch_mod.from_numpy(b).cuda()
        got = apply_nd(ib, "c2c", bd, torch_mod.empty_like(bd), [1], True, 1.0).cpu().numpy()
        assert oracle.rel_l2(got, checker.c2c(b, [1], True, 1.0)) <= tol(128, dt)
    print(sorted(used))
    assert any(k.startswith("colfast2") for k in used)


def test_host_staging_pipeline(ib, torch_mod, checker):
    """Host arrays large enough to take the chunked H2D -> kernel -> D2H pipeline must give exactly what
    the device-resident call gives (same kernels, same arithmetic), in and out of place, for the
    register kernels and the generic engine; plus parity of sampled rows against the oracle."""
    rng = np.random.default_rng(41)
    x = rnd(rng, (6000, 1024), np.complex128)                       # 94 MiB each way: fast2p, several chunks
    want_dev = ib.fft(torch_mod.from_numpy(x).cuda()).cpu().numpy()
    got = ib.fft(x)
    assert np.array_equal(got, want_dev)
    assert oracle.max_row_rel_l2(got[::997], checker.c2c(x[::997], [1])) <= tol(1024)
    y = x.copy()
    ib.fft_inplace(y)
    assert np.array_equal(y, want_dev)
    r = rnd(rng, (3000, 4096), np.float64)                          # r2c through fast3 2048, out of place
    spec = np.empty((3000, 2049), np.complex128)
    apply_nd(ib, "r2c", r, spec, [1])
    sd = torch_mod.empty((3000, 2049), dtype=torch_mod.complex128, device="cuda")
    apply_nd(ib, "r2c", torch_mod.from_numpy(r).cuda(), sd, [1])
    assert np.array_equal(spec, sd.cpu().numpy())
    p = rnd(rng, (5000, 1000), np.float64)                          # packed in-place layout of the C API, generic engine
    q = ib.rfft_packed(p)
    assert oracle.max_row_rel_l2(q[::499], checker.rfft_rows(p[::499].copy(), True, 1.0)) <= tol(1000)
    assert oracle.max_row_rel_l2(ib.rfft_packed(q, forward=False), p) <= 2e-15 * 10
    v = rnd(rng, (4000, 600), np.complex128)[:, ::2]                # strided view: gaps in the input rows
    out = np.empty((4000, 300), np.complex128)
    apply_nd(ib, "c2c", v, out, [1])
    assert oracle.max_row_rel_l2(out[::333], checker.c2c(np.ascontiguousarray(v[::333]), [1])) <= tol(300)


def test_fused_bluestein(ib, torch_mod, checker):
    """Fused register Bluestein (fastblue): complex prime lengths across the three work sizes, including
    work lengths SHORTER than 2L-1 (alias corrections: 1031->2048 is too short by 13 and goes to 4096,
    2051->4096 by 5, 4099->8192 by 5, 4100 even -> generic), and odd real lengths with two rows packed per
    complex line (odd and even row counts, both `forward` flags)."""
    rng = np.random.default_rng(51)
    used = set()
    for n in (257, 521, 1021, 1031, 2039, 2051, 2053, 3001, 4093, 4099, 4101):
        for rows in (1, 2, 7, 64):
            x = rnd(rng, (rows, n), np.complex128)
            xd = torch_mod.from_numpy(x).cuda()
            for fwd in (True, False):
                got = apply_nd(ib, "c2c", xd, torch_mod.empty_like(xd), [1], fwd, 0.9).cpu().numpy()
                used.add(ib.last_kernel())
                assert oracle.max_row_rel_l2(got, checker.c2c(x, [1], fwd, 0.9)) <= tol(n), (n, rows, fwd)
    for n in (263, 1019, 2043, 4099):   # 2043 = 3^2 * 227
        for rows in (1, 2, 5, 32):
            r = rnd(rng, (rows, n), np.float64)
            rd = torch_mod.from_numpy(r).cuda()
            for fwd in (True, False):
                spec = apply_nd(ib, "r2c", rd, torch_mod.empty((rows, n // 2 + 1), dtype=torch_mod.complex128, device="cuda"),
                                [1], fwd, 1.0)
                used.add(ib.last_kernel())
                assert oracle.max_row_rel_l2(spec.cpu().numpy(), checker.r2c(r, [1], fwd, 1.0)) <= tol(n), (n, rows, fwd)
            sp = checker.r2c(r, [1], True, 1.0)
            sd = torch_mod.from_numpy(sp).cuda()
            for fwd in (False, True):
                back = apply_nd(ib, "c2r", sd, torch_mod.empty_like(rd), [1], fwd, 1.0 / n).cpu().numpy()
                used.add(ib.last_kernel())
                assert oracle.max_row_rel_l2(back, checker.c2r(sp, r.shape, [1], fwd, 1.0 / n)) <= tol(n), (n, rows, fwd)
            rt = apply_nd(ib, "c2r", apply_nd(ib, "r2c", rd, torch_mod.empty((rows, n // 2 + 1), dtype=torch_mod.complex128,
                                                                            device="cuda"), [1]),
                          torch_mod.empty_like(rd), [1], False, 1.0 / n).cpu().numpy()
            assert oracle.max_row_rel_l2(rt, r) <= 2e-15 * np.log2(n), (n, rows)
    # float32 instances of the fused kernel: selected (and checked here) with IMPULSE_FFT_BLUE_F32=1 until they are
    # measured and become the default
    used32 = set()
    for n in ((1021, 2051, 4099) if os.environ.get("IMPULSE_FFT_BLUE_F32", "1") == "1" else ()):
        for rows in (1, 5, 32):
            x = rnd(rng, (rows, n), np.complex64)
            xd = torch_mod.from_numpy(x).cuda()
            got = apply_nd(ib, "c2c", xd, torch_mod.empty_like(xd), [1], True, 1.0).cpu().numpy()
            used32.add(ib.last_kernel())
            assert oracle.max_row_rel_l2(got, checker.c2c(x, [1], True, 1.0)) <= tol(n, np.float32), (n, rows)
            r = rnd(rng, (rows, n), np.float32)
            rd = torch_mod.from_numpy(r).cuda()
            spec = apply_nd(ib, "r2c", rd, torch_mod.empty((rows, n // 2 + 1), dtype=torch_mod.complex64, device="cuda"), [1], True, 1.0)
            used32.add(ib.last_kernel())
            want = checker.r2c(r, [1], True, 1.0)
            assert oracle.max_row_rel_l2(spec.cpu().numpy(), want) <= tol(n, np.float32), (n, rows)
            back = apply_nd(ib, "c2r", torch_mod.from_numpy(want).cuda(), torch_mod.empty_like(rd), [1], False, 1.0 / n).cpu().numpy()
            assert oracle.max_row_rel_l2(back, r) <= tol(n, np.float32), (n, rows)
    print(sorted(used), sorted(used32))
    assert any(k.startswith("fastblue") for k in used)
    if os.environ.get("IMPULSE_FFT_BLUE_F32", "1") == "1":
        assert any(k.startswith("fastblue_kernel<float") for k in used32), used32


def test_dct_dst(ib, torch_mod, checker):
    """DCTDesc.init(axes, dctType, ortho, scalingFactor).apply (cpp_pocketfft/pocketfft.nim:217-233,279-295):
    every type, both families, host and device buffers, N-D."""
    rng = np.random.default_rng(61)
    for dt, rt in ((np.float64, 1e-12), (np.float32, 1e-5)):
        # 4096 (type I: 8190 / 8194-point embeddings), 8192 and 10000 run as embed -> complex transform -> extract: any N,
        # as pocketfft's T_dct1 / T_dcst23 / T_dcst4 (pocketfft_hdronly.h:2424-2648)
        for n in (2, 4, 9, 64, 100, 1000, 4096, 8192, 10000):
            x = rnd(rng, (5, n), dt)
            xd = torch_mod.from_numpy(x).cuda()
            for sine in (False, True):
                for t in (1, 2, 3, 4):
                    for ortho in (False, True):
                        want = checker.r2r(not sine, t, x, [1], 0.5, ortho)
                        out = torch_mod.empty_like(xd)
                        ib.DCTDesc.init(axes=[1], dctType=t, ortho=ortho, scalingFactor=0.5, sine=sine).apply(
                            ib.DataDesc.init(out), ib.DataDesc.init(xd))
                        assert oracle.max_row_rel_l2(out.cpu().numpy(), want) <= rt * np.log2(max(n, 2)), (n, sine, t, ortho, dt)
    # the reference's own example (pocketfft.nim:322-339), host buffers
    d_in = np.array([4.0, 3.0, 5.0, 10.0])
    d_out = np.zeros(4)
    ib.DCTDesc.init(axes=[0], dctType=2).apply(ib.DataDesc.init(d_out), ib.DataDesc.init(d_in))
    np.testing.assert_allclose(d_out, checker.r2r(True, 2, d_in.reshape(1, 4), [1])[0], rtol=1e-14)
    a = rnd(rng, (12, 20, 6), np.float64)
    out = np.empty_like(a)
    ib.DCTDesc.init(axes=[0, 1], dctType=2, ortho=True).apply(ib.DataDesc.init(out), ib.DataDesc.init(a))
    assert oracle.rel_l2(out, checker.r2r(True, 2, a, [0, 1], 1.0, True)) <= 1e-12 * 5
    big = rnd(rng, (40000, 3), np.float64)          # a long strided axis, in place
    want = checker.r2r(True, 2, big, [0], 1.0, False)
    ib.DCTDesc.init(axes=[0], dctType=2).apply(ib.DataDesc.init(big), ib.DataDesc.init(big))
    assert oracle.rel_l2(big, want) <= 1e-12 * 16
    with pytest.raises(ValueError):
        ib.DCTDesc.init(axes=[0], dctType=5)
    with pytest.raises(ib.FFTError):
        ib.DCTDesc.init(axes=[0], dctType=1).apply(ib.DataDesc.init(np.zeros(1)), ib.DataDesc.init(np.zeros(1)))


def test_cols_from_parts_single_gpu(ib, torch_mod, checker):
    """impulse_fft_cols_from_parts with the parts in separate allocations of ONE GPU: the segmented-load
    logic of the fused multi-GPU column pass, isolated from CUDA IPC (tests/test_gpu_dist.py covers that)."""
    import ctypes as C
    from impulse_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(71)
    for (nparts, rpp, cols, col0, ncols) in ((2, 32, 96, 48, 48), (2, 32, 96, 0, 48), (4, 256, 64, 16, 24), (8, 1024, 40, 8, 32),
                                             (2, 4096, 24, 0, 24), (1, 64, 16, 0, 16)):
        full = rnd(rng, (nparts * rpp, cols), np.complex128)
        parts = [torch_mod.from_numpy(full[q * rpp:(q + 1) * rpp].copy()).cuda() for q in range(nparts)]
        out = torch_mod.empty((nparts * rpp, ncols), dtype=torch_mod.complex128, device="cuda")
        ptrs = (C.c_void_p * nparts)(*[p.data_ptr() for p in parts])
        for fwd in (True, False):
            _lib.check(L.impulse_fft_cols_from_parts(_lib.F64, nparts, ptrs, rpp, cols, col0, ncols, out.data_ptr(), ncols,
                                                     int(fwd), 0.5, None))
            want = checker.c2c(np.ascontiguousarray(full[:, col0:col0 + ncols]), [0], fwd, 0.5)
            assert oracle.rel_l2(out.cpu().numpy(), want) <= tol(nparts * rpp), (nparts, rpp, cols, col0, ncols, fwd)


def test_config5_full_image_size_properties(ib, torch_mod, checker):
    """BASELINE config 5 at the full 4096 x 4096 float32 image size (batch reduced to 4): identity kernel
    returns the image (2-D r2c -> c2r round trip), a 31x31 kernel matches the direct sum at sampled
    pixels, and the transform is linear."""
    from impulse_b200.filter import FFTFilter2D
    g = torch_mod.Generator(device="cuda").manual_seed(77)
    img = torch_mod.rand((4, 4096, 4096), generator=g, device="cuda", dtype=torch_mod.float32)
    delta = torch_mod.zeros((31, 31), device="cuda", dtype=torch_mod.float32)
    delta[15, 15] = 1.0
    same = FFTFilter2D(delta, 4096, 4096).apply(img)
    err = float(torch_mod.linalg.vector_norm(same - img) / torch_mod.linalg.vector_norm(img))
    assert err <= 1e-5 * 12, err
    ker = torch_mod.rand((31, 31), generator=g, device="cuda", dtype=torch_mod.float32)
    ker /= ker.sum()
    f = FFTFilter2D(ker, 4096, 4096)
    out = f.apply(img)
    kn = ker.cpu().numpy().astype(np.float64)
    im0 = img[1].cpu().numpy().astype(np.float64)
    for (r, c) in ((0, 0), (15, 4000), (2048, 2048), (4095, 4095), (100, 7)):
        rr = (r - (np.arange(31) - 15)) % 4096      # circular: out[r,c] = sum k[i,j] * img[r-(i-15), c-(j-15)]
        cc = (c - (np.arange(31) - 15)) % 4096
        direct = float((kn * im0[np.ix_(rr, cc)]).sum())
        assert abs(float(out[1, r, c]) - direct) <= 2e-5, (r, c)
    # every pixel of every image against the same composition on the oracle (float32 pocketfft, all host threads)
    pad = np.zeros((4096, 4096), np.float32)
    ii = (np.arange(31) - 15) % 4096
    pad[np.ix_(ii, ii)] = ker.cpu().numpy()
    kspec = checker.r2c(pad, [0, 1], True, 1.0, nthreads=0)
    imgs = img.cpu().numpy()
    spec = checker.r2c(imgs, [1, 2], True, 1.0, nthreads=0) * kspec[None]
    want = checker.c2r(np.ascontiguousarray(spec), imgs.shape, [1, 2], False, 1.0 / (4096.0 * 4096.0), nthreads=0)
    got = out.cpu().numpy()
    for b in range(4):
        assert oracle.rel_l2(got[b], want[b]) <= 1e-5 * 12, b
    assert float(np.abs(got - want).max()) <= 2e-5
    del spec, want, got, imgs
    lin = f.apply(2.0 * img[:2] + img[2:4])
    ref = 2.0 * out[:2] + out[2:4]
    assert float(torch_mod.linalg.vector_norm(lin - ref) / torch_mod.linalg.vector_norm(ref)) <= 1e-5


def test_c2c_fused_multiply(ib, torch_mod, checker):
    """impulse_fft_c2c_mul: out = c2c(in) * mul[offset % period], through every kernel that can carry it
    (generic line kernel, strided register kernel, both launches of the split), in place and out of place."""
    import ctypes as C
    from impulse_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(81)
    used = set()
    cases = [((6, 64), [1], 64, np.complex128), ((3, 40, 24), [1], 960, np.complex128), ((3, 40, 24), [1, 2], 960, np.complex64),
             ((2, 4099), [1], 2 * 4099, np.complex128), ((5, 16384), [1], 16384, np.complex128),
             ((3, 256, 100), [1], 25600, np.complex64), ((2, 4096, 2049), [1], 4096 * 2049, np.complex64),
             ((4, 1024, 24), [1], 1024 * 24 * 2, np.complex128)]
    for shape, axes, period, cdt in cases:
        x = rnd(rng, shape, cdt)
        m = rnd(rng, (period,), cdt)
        xd, md = torch_mod.from_numpy(x).cuda(), torch_mod.from_numpy(m).cuda()
        code = _lib.F64 if cdt == np.complex128 else _lib.F32
        n = len(shape)
        st = (C.c_ssize_t * n)(*x.strides)
        for fwd in (True, False):
            want = (checker.c2c(x, axes, fwd, 0.5).reshape(-1, period) * m).reshape(shape)
            for inplace in (False, True):
                src = xd.clone()
                dst = src if inplace else torch_mod.empty_like(src)
                _lib.check(L.impulse_fft_c2c_mul(code, n, (C.c_size_t * n)(*shape), st, st, len(axes),
                                                 (C.c_size_t * len(axes))(*axes), int(fwd), src.data_ptr(), dst.data_ptr(), 0.5,
                                                 md.data_ptr(), period, None))
                used.add(ib.last_kernel())
                big = max(shape[a] for a in axes)
                assert oracle.rel_l2(dst.cpu().numpy(), want) <= 2 * tol(big, np.float64 if cdt == np.complex128 else np.float32), \
                    (shape, axes, fwd, inplace)
    print(sorted(used))
    assert any(k.startswith("colfast2") for k in used) and any(k.startswith("line_fft") for k in used)
    # a real transform cannot take a multiplier; host pointers are refused
    x = np.zeros((4, 8), np.complex128)
    st = (C.c_ssize_t * 2)(*x.strides)
    rc = L.impulse_fft_c2c_mul(_lib.F64, 2, (C.c_size_t * 2)(4, 8), st, st, 1, (C.c_size_t * 1)(1), 1, x.ctypes.data,
                               x.ctypes.data, 1.0, x.ctypes.data, 8, None)
    assert rc == -1  # IMPULSE_FFT_ERR_INVALID


def test_fftpack_and_hartley(ib, torch_mod, ref):
    """impulse_fft_r2r_fftpack / _separable_hartley / _genuine_hartley against the compiled reference
    (pocketfft_hdronly.h:3392-3445): 1-D and N-D, both precisions, all four fftpack flag combinations,
    Bluestein lengths, in place, host buffers."""
    rng = np.random.default_rng(91)
    cases = [((7,), [0]), ((8,), [0]), ((1,), [0]), ((3, 10), [1]), ((6, 9), [0, 1]), ((4, 5, 6), [2, 0]), ((2, 191), [1]),
             ((64, 1000), [1]), ((16, 4099), [1]), ((512, 24), [0]), ((12, 16, 10), [0, 1, 2]), ((8, 4096), [1]),
             ((256, 256), [0, 1])]
    for dt in (np.float64, np.float32):
        for shape, axes in cases:
            a = rng.standard_normal(shape).astype(dt)
            ad = torch_mod.from_numpy(a).cuda()
            big = max(shape[x] for x in axes)
            t = tol(max(big, 2), dt) * len(axes)
            for r2h in (True, False):
                for fwd in (True, False):
                    out = torch_mod.empty_like(ad)
                    ib.r2r_fftpack(ib.DataDesc.init(out), ib.DataDesc.init(ad), axes, r2h, fwd, 0.25)
                    want = ref.r2r_real("fftpack", a, axes, r2h, fwd, 0.25)
                    assert oracle.rel_l2(out.cpu().numpy(), want) <= t, (shape, axes, r2h, fwd, dt)
            for name, fn in (("separable_hartley", ib.r2r_separable_hartley), ("genuine_hartley", ib.r2r_genuine_hartley)):
                out = torch_mod.empty_like(ad)
                fn(ib.DataDesc.init(out), ib.DataDesc.init(ad), axes, 0.5)
                want = ref.r2r_real(name, a, axes, fct=0.5)
                assert oracle.rel_l2(out.cpu().numpy(), want) <= t, (shape, axes, name, dt)
            inpl = ad.clone()                                       # in place
            ib.r2r_separable_hartley(ib.DataDesc.init(inpl), ib.DataDesc.init(inpl), axes)
            assert oracle.rel_l2(inpl.cpu().numpy(), ref.r2r_real("separable_hartley", a, axes)) <= t
            host = np.empty_like(a)                                 # host buffers are staged
            ib.r2r_genuine_hartley(ib.DataDesc.init(host), ib.DataDesc.init(a), axes)
            assert oracle.rel_l2(host, ref.r2r_real("genuine_hartley", a, axes)) <= t
    # Hartley is an involution up to 1/N (size-independent property at a full-size image)
    img = torch_mod.rand((2048, 2048), device="cuda", dtype=torch_mod.float64)
    h1, h2 = torch_mod.empty_like(img), torch_mod.empty_like(img)
    ib.r2r_genuine_hartley(ib.DataDesc.init(h1), ib.DataDesc.init(img), [0, 1])
    ib.r2r_genuine_hartley(ib.DataDesc.init(h2), ib.DataDesc.init(h1), [0, 1], 1.0 / (2048 * 2048))
    assert float(torch_mod.linalg.vector_norm(h2 - img) / torch_mod.linalg.vector_norm(img)) <= 1e-13


def test_strided_axis_whole_instances(ib, torch_mod, checker, monkeypatch):
    """every instance of the whole-axis kernel in plain-transform mode (IMPULSE_FFT_COL_WHOLE=2; by default only
    complex128 axes of 1024 points take it), forward and backward, vector and scalar global access, in place"""
    monkeypatch.setenv("IMPULSE_FFT_COL_WHOLE", "2")
    rng = np.random.default_rng(99)
    used = set()
    for shape, cdt in (((3, 1024, 22), np.complex64), ((2, 1024, 7), np.complex64), ((3, 2048, 12), np.complex64),
                       ((5, 1024, 6), np.complex128), ((2, 2048, 9), np.complex128), ((1024, 35), np.complex128)):
        axis = len(shape) - 2
        x = rnd(rng, shape, cdt)
        xd = torch_mod.from_numpy(x).cuda()
        for fwd in (True, False):
            yd = torch_mod.empty_like(xd)
            ib.FFTDesc.init(axes=[axis], forward=fwd, scalingFactor=0.5).apply(ib.DataDesc.init(yd), ib.DataDesc.init(xd))
            used.add(ib.last_kernel())
            assert oracle.rel_l2(yd.cpu().numpy(), checker.c2c(x, [axis], fwd, 0.5)) <= 2 * tol(shape[axis], np.float64 if cdt == np.complex128 else np.float32), (shape, fwd)
        zd = xd.clone()
        ib.FFTDesc.init(axes=[axis], forward=True).apply(ib.DataDesc.init(zd), ib.DataDesc.init(zd))
        assert oracle.rel_l2(zd.cpu().numpy(), checker.c2c(x, [axis], True, 1.0)) <= 2 * tol(shape[axis], np.float64 if cdt == np.complex128 else np.float32)
    assert all(k.startswith("colconvw_kernel") for k in used), used
    assert any("+bwd" in k for k in used) and any("+scalar" in k for k in used), used


def test_convolve_axis_whole_long(ib, torch_mod, checker, monkeypatch):
    """the whole-axis kernel at 4096 points (32-byte runs: off by default, IMPULSE_FFT_CONV_WHOLE=2) and at 2048"""
    import ctypes as C
    from impulse_b200 import _lib
    monkeypatch.setenv("IMPULSE_FFT_CONV_WHOLE", "2")
    L = _lib.lib()
    rng = np.random.default_rng(98)
    used = set()
    for shape, cdt in (((3, 2048, 11), np.complex64), ((2, 2048, 6), np.complex128), ((5, 4096, 14), np.complex64), ((3, 4096, 5), np.complex128)):
        x = rnd(rng, shape, cdt)
        m = rnd(rng, (shape[1] * shape[2],), cdt)
        n = shape[1]
        spec = checker.c2c(x, [1], True, 1.0) * m.reshape(shape[1:])
        want = checker.c2c(spec.astype(cdt), [1], False, 1.0 / n)
        xd, md = torch_mod.from_numpy(x).cuda(), torch_mod.from_numpy(m).cuda()
        st = (C.c_ssize_t * 3)(*x.strides)
        _lib.check(L.impulse_fft_convolve_axis(_lib.F64 if cdt == np.complex128 else _lib.F32, 3, (C.c_size_t * 3)(*shape), st, st, 1,
                                               xd.data_ptr(), xd.data_ptr(), 1.0 / n, md.data_ptr(), m.size, None))
        used.add(ib.last_kernel())
        assert oracle.rel_l2(xd.cpu().numpy(), want) <= 3 * tol(n, np.float64 if cdt == np.complex128 else np.float32), (shape, cdt)
    assert all(k.startswith("colconvw_kernel") for k in used), used


def test_convolve_axis(ib, torch_mod, checker):
    """impulse_fft_convolve_axis: IFFT_axis(FFT_axis(x) * m) — the whole-axis kernel (colconvw_kernel, one launch)
    on strided power-of-two axes of 512..4096 points, the fused three-pass plan (colconv2 kernel) on 8192 / 16384
    points, ragged column counts, both precisions, in place and out of place; and the plain two-transform plan
    everywhere else.  Checked against the oracle's transforms."""
    import ctypes as C
    from impulse_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(97)
    used = set()
    cases = [((2, 1024, 24), np.complex128), ((2, 2048, 40), np.complex64), ((3, 4096, 29), np.complex64),
             ((2, 4096, 16), np.complex128), ((1, 8192, 16), np.complex64), ((2, 16384, 8), np.complex128),
             ((4, 64), np.complex128), ((3, 40, 24), np.complex64), ((2, 4099, 8), np.complex128), ((2, 4096), np.complex128),
             ((3, 512, 21), np.complex64), ((2, 512, 10), np.complex128), ((2, 1024, 9), np.complex64), ((5, 2048, 9), np.complex128),
             ((7, 4096, 13), np.complex64), ((2, 3, 4096, 7), np.complex128)]
    for shape, cdt in cases:
        axis = len(shape) - 2 if len(shape) > 2 else 1
        x = rnd(rng, shape, cdt)
        period = int(np.prod(shape[axis:]))
        m = rnd(rng, (period,), cdt)
        n = shape[axis]
        spec = checker.c2c(x, [axis], True, 1.0) * m.reshape(shape[axis:])
        want = checker.c2c(spec.astype(cdt), [axis], False, 1.0 / n)
        xd, md = torch_mod.from_numpy(x).cuda(), torch_mod.from_numpy(m).cuda()
        code = _lib.F64 if cdt == np.complex128 else _lib.F32
        nd = len(shape)
        st = (C.c_ssize_t * nd)(*x.strides)
        for inplace in (False, True):
            src = xd.clone()
            dst = src if inplace else torch_mod.empty_like(src)
            _lib.check(L.impulse_fft_convolve_axis(code, nd, (C.c_size_t * nd)(*shape), st, st, axis, src.data_ptr(),
                                                   dst.data_ptr(), 1.0 / n, md.data_ptr(), period, None))
            used.add(ib.last_kernel())
            assert oracle.rel_l2(dst.cpu().numpy(), want) <= 3 * tol(n, np.float64 if cdt == np.complex128 else np.float32), \
                (shape, cdt, inplace)
    print(sorted(used))
    assert any(k.startswith("colconvw_kernel<float,16,8,8") for k in used), used
    assert any(k.startswith("colconvw_kernel<double,8,8,8") for k in used), used
    # launches per call: the whole-axis kernel is one, the three-pass plan (pass A, colconv2, pass B) three
    for n, launches in ((2048, 1), (4096, 3)):
        before = ib.launch_count()
        x = torch_mod.zeros((2, n, 32), dtype=torch_mod.complex64, device="cuda")
        m = torch_mod.ones(n * 32, dtype=torch_mod.complex64, device="cuda")
        st = (C.c_ssize_t * 3)(n * 32 * 8, 32 * 8, 8)
        _lib.check(L.impulse_fft_convolve_axis(_lib.F32, 3, (C.c_size_t * 3)(2, n, 32), st, st, 1, x.data_ptr(), x.data_ptr(), 1.0,
                                               m.data_ptr(), n * 32, None))
        assert ib.launch_count() - before == launches, (n, ib.launch_count() - before)


def test_long_lines(ib, torch_mod, checker):
    """Lines beyond one CTA's shared memory: real transforms with N/2 > 14464 or odd N > 14464 (complex
    transform on a work array + elementwise conversion passes), complex Bluestein lengths whose line does
    not fit (elementwise chirp passes), through the high-level API (in-place packed, rfft) and DataDesc."""
    rng = np.random.default_rng(101)
    for n in (32768, 65536, 30000, 29999, 40001, 131072, 1 << 20):
        rows = 2 if n < (1 << 20) else 1
        x = rng.uniform(-0.5, 0.5, (rows, n))
        xd = torch_mod.from_numpy(x).cuda()
        spec = torch_mod.empty((rows, n // 2 + 1), dtype=torch_mod.complex128, device="cuda")
        apply_nd(ib, "r2c", xd, spec, [1], True, 1.0)
        want = checker.r2c(x, [1], True, 1.0)
        assert oracle.max_row_rel_l2(spec.cpu().numpy(), want) <= tol(n), n
        back = torch_mod.empty_like(xd)
        apply_nd(ib, "c2r", spec, back, [1], False, 1.0 / n)
        assert oracle.max_row_rel_l2(back.cpu().numpy(), x) <= 2e-15 * np.log2(n), n
        d = xd[0].clone()                               # FFTPACK packing, in place (the C-backend API)
        ib.fft_inplace(d, forward=True)
        wantp = checker.rfft_rows(x[:1].copy(), True, 1.0)[0]
        assert oracle.rel_l2(d.cpu().numpy(), wantp) <= tol(n), n
        ib.fft_inplace(d, forward=False)
        assert oracle.rel_l2(d.cpu().numpy(), x[0]) <= 2e-15 * np.log2(n), n
    xs = rng.uniform(-0.5, 0.5, (3, 32768)).astype(np.float32)          # float32, host buffers
    out = np.empty((3, 16385), np.complex64)
    apply_nd(ib, "r2c", xs, out, [1], True, 1.0)
    assert oracle.max_row_rel_l2(out, checker.r2c(xs, [1], True, 1.0)) <= tol(32768, np.float32)
    for n in (100003, 20011, 65537):                                     # complex primes
        z = rnd(rng, (2, n), np.complex128)
        zd = torch_mod.from_numpy(z).cuda()
        for fwd in (True, False):
            got = apply_nd(ib, "c2c", zd, torch_mod.empty_like(zd), [1], fwd, 0.5).cpu().numpy()
            assert oracle.max_row_rel_l2(got, checker.c2c(z, [1], fwd, 0.5)) <= tol(n), (n, fwd)
    # long EVEN real lines that are strided or whose rows are an odd number of elements apart (round-1 advice: these
    # returned ERR_UNSUPPORTED; general_r2c / general_c2r take any byte stride, pocketfft_hdronly.h:3125-3250)
    y = rng.uniform(-0.5, 0.5, (40000, 4))
    yd = torch_mod.from_numpy(y).cuda()
    spec = apply_nd(ib, "r2c", yd, torch_mod.empty((20001, 4), dtype=torch_mod.complex128, device="cuda"), [0], True, 1.0)
    assert oracle.rel_l2(spec.cpu().numpy(), checker.r2c(y, [0], True, 1.0)) <= tol(40000)
    back = apply_nd(ib, "c2r", spec, torch_mod.empty_like(yd), [0], False, 1.0 / 40000)
    assert oracle.rel_l2(back.cpu().numpy(), y) <= 2e-15 * 16
    wide = torch_mod.from_numpy(rng.uniform(-0.5, 0.5, (3, 40001))).cuda()
    yv = wide[:, :40000]                                                   # row stride 40001 elements
    spec = apply_nd(ib, "r2c", yv, torch_mod.empty((3, 20001), dtype=torch_mod.complex128, device="cuda"), [1], True, 1.0)
    assert oracle.max_row_rel_l2(spec.cpu().numpy(), checker.r2c(np.ascontiguousarray(yv.cpu().numpy()), [1], True, 1.0)) <= tol(40000)
    a = rng.uniform(-0.5, 0.5, (2, 32768))                                 # r2r_fftpack, real2hermitian != forward, long line
    got = np.empty_like(a)
    ib.r2r_fftpack(ib.DataDesc.init(got), ib.DataDesc.init(a), [1], True, False, 1.0)
    assert oracle.rel_l2(got, oracle.fftpack_numpy(a, [1], True, False, 1.0)) <= tol(32768)
    # 2-D real transform whose last axis is long
    img = rng.uniform(-0.5, 0.5, (6, 40000))
    got = apply_nd(ib, "r2c", torch_mod.from_numpy(img).cuda(), torch_mod.empty((6, 20001), dtype=torch_mod.complex128,
                                                                                  device="cuda"), [0, 1], True, 1.0)
    assert oracle.rel_l2(got.cpu().numpy(), checker.r2c(img, [0, 1], True, 1.0)) <= tol(40000)


def test_randomized_shapes_strides_against_reference(ib, torch_mod, ref):
    """300 seeded random cases against the compiled reference: kind (c2c/r2c/c2r/dct/dst/hartley/fftpack),
    precision, 1-4 dimensions, any subset and order of axes, lengths drawn to hit every kernel family
    (short rows, register sizes, composite, prime, long), non-contiguous views on both sides (steps, offsets,
    permuted dimensions) and misaligned bases — the kernel SELECTION logic is what this exercises."""
    rng = np.random.default_rng(20261017)
    pool = [1, 2, 3, 4, 5, 7, 8, 12, 16, 17, 30, 32, 33, 64, 97, 100, 128, 191, 243, 255, 256, 257, 500, 512, 625, 1000, 1024,
            1025, 2000, 2048, 4096, 4099, 8192, 16384, 20000]
    kinds_seen, kernels = set(), set()
    for case in range(300):
        kind = ["c2c", "c2c", "r2c", "r2c", "c2r", "c2r", "dct", "dst", "hartley_s", "hartley_g", "fftpack"][rng.integers(0, 11)]
        f64 = bool(rng.integers(0, 2))
        rdt, cdt = (np.float64, np.complex128) if f64 else (np.float32, np.complex64)
        nd = int(rng.integers(1, 5))
        budget = 1 << 19
        shape = []
        for _ in range(nd):
            cand = [p for p in pool if p <= max(1, budget)]
            n = int(cand[rng.integers(0, len(cand))])
            shape.append(n)
            budget //= n
        rng.shuffle(shape)
        shape = tuple(int(s) for s in shape)
        naxes = int(rng.integers(1, nd + 1))
        axes = [int(a) for a in rng.permutation(nd)[:naxes]]
        if kind in ("dct", "dst") and any(shape[a] > 4000 for a in axes):
            kind = "c2c"
        if kind in ("dct", "dst", "hartley_s", "hartley_g", "fftpack"):
            axes = axes[:2]
        fwd = bool(rng.integers(0, 2))
        fct = float(rng.choice([1.0, 0.5, 1.0 / 3.0]))

        def view_of(shp, dtype):
            """A device tensor of logical shape `shp` that is a strided view into a larger allocation."""
            steps = [int(rng.choice([1, 1, 1, 2])) for _ in shp]
            offs = [int(rng.integers(0, 2)) for _ in shp]
            big = tuple(o + s * (n - 1) + 1 + int(rng.integers(0, 2)) for o, s, n in zip(offs, steps, shp))
            base = torch_mod.zeros(big, dtype=dtype, device="cuda")
            sl = tuple(slice(o, o + s * (n - 1) + 1, s) for o, s, n in zip(offs, steps, shp))
            return base[sl]

        tdt = {np.float64: torch_mod.float64, np.float32: torch_mod.float32, np.complex128: torch_mod.complex128,
               np.complex64: torch_mod.complex64}
        last = axes[-1]
        cshape = tuple(s // 2 + 1 if i == last else s for i, s in enumerate(shape))
        if kind == "c2c":
            a = rnd(rng, shape, cdt)
            want = ref.c2c(a, axes, fwd, fct)
            src = view_of(shape, tdt[cdt])
            dst = src if rng.integers(0, 3) == 0 else view_of(shape, tdt[cdt])     # one in three in place
        elif kind == "r2c":
            a = rnd(rng, shape, rdt)
            want = ref.r2c(a, axes, fwd, fct)
            src, dst = view_of(shape, tdt[rdt]), view_of(cshape, tdt[cdt])
        elif kind == "c2r":
            a = rnd(rng, cshape, cdt)
            want = ref.c2r(a, shape, axes, fwd, fct)
            src, dst = view_of(cshape, tdt[cdt]), view_of(shape, tdt[rdt])
        else:
            a = rnd(rng, shape, rdt)
            src, dst = view_of(shape, tdt[rdt]), view_of(shape, tdt[rdt])
        src.copy_(torch_mod.from_numpy(a).cuda())
        din, dout = ib.DataDesc.init(src), ib.DataDesc.init(dst)
        typ = int(rng.integers(1, 5))
        if kind in ("c2c", "r2c", "c2r"):
            ib.FFTDesc.init(axes=axes, forward=fwd, scalingFactor=fct).apply(dout, din)
        elif kind in ("dct", "dst"):
            ortho = bool(rng.integers(0, 2))
            if typ == 1 and any(shape[x] < 2 for x in axes):
                continue
            want = ref.r2r(kind == "dct", typ, a, axes, fct, ortho)
            ib.DCTDesc.init(axes=axes, dctType=typ, ortho=ortho, scalingFactor=fct, sine=(kind == "dst")).apply(dout, din)
        elif kind == "fftpack":
            r2h = bool(rng.integers(0, 2))
            want = ref.r2r_real("fftpack", a, axes, r2h, fwd, fct)
            ib.r2r_fftpack(dout, din, axes, r2h, fwd, fct)
        else:
            name = "separable_hartley" if kind == "hartley_s" else "genuine_hartley"
            want = ref.r2r_real(name, a, axes, fct=fct)
            (ib.r2r_separable_hartley if kind == "hartley_s" else ib.r2r_genuine_hartley)(dout, din, axes, fct)
        kinds_seen.add(kind)
        kernels.add(ib.last_kernel().split("<")[0])
        got = dst.cpu().numpy()
        big = max(shape[x] for x in axes)
        bound = (1e-12 if f64 else 1e-5) * max(1.0, np.log2(big)) * len(axes) * (4 if kind in ("dct", "dst") else 1)
        err = oracle.rel_l2(got, want)
        assert err <= bound, (case, kind, f64, shape, axes, fwd, typ, err, ib.last_kernel())
    print(sorted(kinds_seen), sorted(kernels))
    assert len(kinds_seen) == 8 and len(kernels) >= 6
