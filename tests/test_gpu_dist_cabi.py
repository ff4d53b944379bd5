"""GPU: the single-process multi-GPU entry points of the C ABI (impulse_fft_dist_*), on however many B200s are visible
(1 on the driver's GPU-test lease, 2-8 under `gpurun --gpus N`), against the oracle.  SURVEY 8(b) export list, 8(e)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import impulse_b200 as ib
    from impulse_b200 import _lib, dist
    n = torch.cuda.device_count()
    counts = sorted({1, 2 if n >= 2 else 1, 4 if n >= 4 else 1, 8 if n >= 8 else 1})
    return ib, _lib, dist, torch, counts


def test_batch_shard_host_arrays(env, checker):
    ib, _lib, dist, torch, counts = env
    rng = np.random.default_rng(5)
    for g in counts:
        devs = list(range(g))
        # c2c rows (config 2 shape, reduced), r2c and c2r (config 3 shapes, reduced; 4099 = Bluestein), odd row counts
        x = rng.uniform(-0.5, 0.5, (1001, 1024)) + 1j * rng.uniform(-0.5, 0.5, (1001, 1024))
        y = np.full_like(x, np.nan)
        p = dist.DistPlan("batch", _lib.C2C, _lib.F64, x.shape, x.strides, y.strides, [1], True, devs)
        assert [p.shard(i) for i in range(g)] == [(1001 * i // g, 1001 * (i + 1) // g) for i in range(g)]
        p(x, y, 0.5)
        assert oracle.max_row_rel_l2(y, checker.c2c(x, [1], True, 0.5)) <= 1e-12 * 10, g
        p.close()
        for n in (1000, 4099):
            r = rng.uniform(-0.5, 0.5, (333, n))
            spec = np.full((333, n // 2 + 1), np.nan, np.complex128)
            p = dist.DistPlan("batch", _lib.R2C, _lib.F64, r.shape, r.strides, spec.strides, [1], True, devs)
            p(r, spec)
            assert oracle.max_row_rel_l2(spec, checker.r2c(r, [1], True, 1.0)) <= 1e-12 * 12, (g, n)
            p.close()
            back = np.full_like(r, np.nan)
            p = dist.DistPlan("batch", _lib.C2R, _lib.F64, r.shape, spec.strides, back.strides, [1], False, devs)
            p(spec, back, 1.0 / n)
            assert oracle.max_row_rel_l2(back, r) <= 1e-12 * 12, (g, n)
            p.close()
        # a 3-D batch with a transform over the two trailing axes (images): float32 r2c
        img = rng.uniform(0, 1, (7, 64, 96)).astype(np.float32)
        sp = np.zeros((7, 64, 49), np.complex64)
        p = dist.DistPlan("batch", _lib.R2C, _lib.F32, img.shape, img.strides, sp.strides, [1, 2], True, devs)
        p(img, sp)
        assert oracle.rel_l2(sp, checker.r2c(img, [1, 2], True, 1.0)) <= 1e-5 * 7, g
        p.close()


def test_slab_fft2_host_and_device(env, checker):
    ib, _lib, dist, torch, counts = env
    rng = np.random.default_rng(6)
    for g in counts:
        devs = list(range(g))
        for (R, Cn) in ((512, 1024), (2048, 2048)):
            x = rng.uniform(-0.5, 0.5, (R, Cn)) + 1j * rng.uniform(-0.5, 0.5, (R, Cn))
            for fwd in (True, False):
                want = checker.c2c(x, [0, 1], fwd, 0.25, nthreads=0)
                y = np.full_like(x, np.nan)
                p = dist.DistPlan("slab", _lib.C2C, _lib.F64, x.shape, x.strides, y.strides, [0, 1], fwd, devs)
                p(x, y, 0.25)                                  # host arrays: natural layout back
                assert oracle.rel_l2(y, want) <= 1e-12 * 11, (g, R, fwd)
                # device shards: row slabs in, column slabs out
                ins = [torch.from_numpy(x[i * R // g:(i + 1) * R // g]).to(f"cuda:{d}") for i, d in enumerate(devs)]
                outs = [torch.full((R, Cn // g), float("nan"), dtype=torch.complex128, device=f"cuda:{d}") for d in devs]
                p.run_parts(ins, outs, 0.25)
                got = np.concatenate([o.cpu().numpy() for o in outs], axis=1)
                assert oracle.rel_l2(got, want) <= 1e-12 * 11, (g, R, fwd, "parts")
                p.close()


def test_dist_errors(env):
    ib, _lib, dist, torch, counts = env
    x = np.zeros((8, 16), np.complex128)
    with pytest.raises(ib.FFTError):     # dimension 0 is the sharded batch dimension
        dist.DistPlan("batch", _lib.C2C, _lib.F64, x.shape, x.strides, x.strides, [0], True, [0])
    with pytest.raises(ib.FFTError):     # a device listed twice
        dist.DistPlan("batch", _lib.C2C, _lib.F64, x.shape, x.strides, x.strides, [1], True, [0, 0])
    with pytest.raises(ib.FFTError):     # the slab transform is 2-D complex over both axes
        dist.DistPlan("slab", _lib.C2C, _lib.F64, x.shape, x.strides, x.strides, [1], True, [0])
    if torch.cuda.device_count() >= 2:
        y = np.zeros((9, 16), np.complex128)
        with pytest.raises(ib.FFTError):  # rows not divisible by the device count
            dist.DistPlan("slab", _lib.C2C, _lib.F64, y.shape, y.strides, y.strides, [0, 1], True, [0, 1])


def test_plain_c_host_drives_the_dist_abi(env):
    """tests/cpp/test_dist_api.c: a C host (no Python, no torch) shards a batch and runs a slab fft2 over every visible
    device, checking against a direct DFT."""
    so_dir = os.path.join(ROOT, "impulse_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "test_dist_api")
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_dist_api.c"),
           "-o", exe, "-L", so_dir, "-limpulse_fft_b200", f"-Wl,-rpath,{so_dir}", "-lm"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    import torch
    out = subprocess.run([exe, str(min(8, torch.cuda.device_count()))], capture_output=True, text=True, timeout=300)
    print(out.stdout[-1500:])
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
