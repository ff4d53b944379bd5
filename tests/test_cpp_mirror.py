"""Builds tests/cpp/test_fft_api.cpp against include/impulse_fft.hpp + libimpulse_fft_b200.so.
CPU: it must compile and link (the host mirror stays in sync with the ABI).  GPU: it must pass —
that is the reference's own test programme (tests/test_fft.nim, tests/test_fft2.nim, README
examples) restated in the compiled host language."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_fft_api")


def build():
    so_dir = os.path.join(ROOT, "impulse_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_fft_api.cpp"), "-o", EXE,
           "-L", so_dir, "-limpulse_fft_b200", f"-Wl,-rpath,{so_dir}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_cpp_mirror_builds():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_tests():
    build()
    out = subprocess.run([EXE, "3"], capture_output=True, text=True, timeout=900)
    print(out.stdout[-2000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "all checks passed" in out.stdout


def test_plain_c_dist_driver_builds_and_fails_loudly_without_a_gpu():
    """tests/cpp/test_dist_api.c (a C host on the impulse_fft_dist_* entry points) compiles against the header and links;
    without a device every create fails with a status code — no CPU path."""
    so_dir = os.path.join(ROOT, "impulse_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "test_dist_api")
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_dist_api.c"),
           "-o", exe, "-L", so_dir, "-limpulse_fft_b200", f"-Wl,-rpath,{so_dir}", "-lm"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    import torch
    if not torch.cuda.is_available():
        run = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=120)
        assert run.returncode == 1 and "dist_create" in run.stdout and "axis 0 accepted" not in run.stdout
