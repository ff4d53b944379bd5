"""CPU-only: the reference arm of bench.py (`--impl reference`, the reference's own pocketfft on the host cores) prints
ONE JSON line with the keys the driver reads; and bench.py refuses to run its own arm without a CUDA device (no CPU
path).  The GPU arm's line is produced on the B200 only (profiles/r02_bench_default.json is the last one)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"] == "batched fp64 FFT throughput (algorithmic GB/s)" and d["config"]["workload"] == "c2c_65536x1024_c128"
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_gpu_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_default.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == d["e2e"]["d2h_bytes_per_step"] == 65536 * 1024 * 16
    assert d["gpu_launches"] == d["steps"]
    # round 2: every row of the step's output is compared with the oracle, and every other BASELINE config rides along
    assert d["accuracy"]["pass"] is True and d["accuracy"]["rows_checked"] == 65536
    cfg = d["configs"]
    assert set(cfg) == {"1_r2c_1024x4096", "3a_r2c_16384x1000", "3a_c2r_16384x1000", "3b_r2c_16384x3888", "3b_c2r_16384x3888",
                        "3c_r2c_16384x4099", "3c_c2r_16384x4099", "4_fft2_8192x8192", "5_filter2d_64x4096x4096"}
    for k, v in cfg.items():
        assert v["accuracy"]["pass"] is True, k
        assert v["ms_per_step"] > 0 and 0 < v["frac_8TBps"] < 1.1 and v["e2e"]["value"] > 0 and v["cpu_baseline"]["value"] > 0, k
    assert d["configs_accuracy_all_pass"] is True


def test_committed_multi_gpu_line_carries_the_slab_transform():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n8.json")))
    assert d["n_gpus"] == 8 and d["scaling"] == "weak"
    slab = d["configs"]["4_fft2_8192x8192_slab"]
    for v in ("fft2_slab_p2p", "fft2_slab_nccl"):
        assert slab[v]["accuracy"]["pass"] is True and slab[v]["accuracy"]["elements_checked"] == 8192 * 8192
        assert slab[v]["scaling"] == "strong" and slab[v]["ms_per_step"] > 0
    assert d["configs"]["2_c2c_65536x1024_strong"]["rows_per_gpu"] == 8192


def test_own_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU path" in (out.stderr + out.stdout)


def test_roofline_traffic_entries_name_the_kernel_they_were_measured_on():
    """profiles/roofline_traffic.json: the figure bench.py reports as roofline.traffic is tied to a kernel name, and the
    one for the default workload is the kernel the committed bench line ran."""
    t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_default.json")))
    ent = t[d["config"]["workload"]]
    assert ent["kernel"] == d["roofline"]["kernel"] and ent["bytes"] == d["roofline"]["traffic"]
    assert os.path.exists(os.path.join(ROOT, ent["capture"]))
    sys.path.insert(0, ROOT)
    import bench
    assert bench.committed_traffic(d["config"]["workload"], ent["kernel"]) == ent["bytes"]
    assert bench.committed_traffic(d["config"]["workload"], "some_other_kernel") is None
