"""CPU-only: the C-ABI library loads, exports every symbol declared in include/*.h, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("impulse_fft_b200.h", "pocketfft.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b([a-z_][a-z0-9_]*)\s*\(", src):
            n = m.group(1)
            if n.startswith(("impulse_fft_", "make_", "destroy_", "cfft_", "rfft_")):
                names.add(n)
    return names


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "impulse_b200", "libimpulse_fft_b200.so")):
        g.build()
    from impulse_b200 import _lib
    return _lib


def test_exports_match_headers(lib):
    L = lib.lib()
    decl = declared_symbols()
    assert decl == set(lib.EXPORTS), decl ^ set(lib.EXPORTS)
    for name in decl:
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_gpu(lib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    import impulse_b200 as ib
    with pytest.raises(ib.FFTError) as ei:
        ib.fft(np.arange(8.0))
    assert ei.value.code == -6
    L = lib.lib()
    assert L.make_cfft_plan(8) is None  # NULL plan, as pocketfft.c:2068-2070 does on failure
    assert L.make_rfft_plan(0) is None


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under impulse_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "impulse_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)
                assert "tests.emu" not in src and "emu.cpp" not in src.replace("(tests/emu)", ""), f
