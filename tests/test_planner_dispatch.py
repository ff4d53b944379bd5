"""CPU-only: which specialised kernel the planner selects for the BASELINE configs (and what the A/B switches change).
A silent fall back to the generic engine would keep every parity test green while costing 3-6x in throughput; this
pins the selection.  Ids are FastId of impulse_b200/csrc/fft_types.h."""
import os
import re

import numpy as np
import pytest

from tests.emu import harness as emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fast_ids():
    src = open(os.path.join(ROOT, "impulse_b200", "csrc", "fft_types.h")).read()
    body = src[src.index("enum FastId"):]
    body = body[:body.index("};")]
    return {m.group(1): int(m.group(2)) for m in re.finditer(r"(\w+)\s*=\s*(\d+)", body)}


ID = fast_ids()


def ids_1d(kind, n, rows, dt=np.float64):
    cdt = np.complex128 if dt == np.float64 else np.complex64
    if kind == "c2c":
        a = b = np.empty((rows, n), cdt)
    elif kind == "r2c":
        a, b = np.empty((rows, n), dt), np.empty((rows, n // 2 + 1), cdt)
    else:
        a, b = np.empty((rows, n // 2 + 1), cdt), np.empty((rows, n), dt)
    shape = (rows, n)
    return emu.nd_fast_ids(kind, a, b, shape, [1], kind != "c2r")


def test_baseline_configs_run_on_register_kernels():
    assert ids_1d("c2c", 1024, 64) == [ID["FAST2_1024_F64"]]                       # config 2
    assert ids_1d("r2c", 4096, 64) == [ID["FAST3_2048_F64"]]                       # config 1
    assert ids_1d("r2c", 1000, 64) == [ID["FAST3R_500_F64"]]                       # config 3a
    assert ids_1d("c2r", 1000, 64) == [ID["FAST3_500_F64"]]
    assert ids_1d("r2c", 3888, 64) == [ID["FAST3R_1944_F64"]]                      # config 3b
    assert ids_1d("c2r", 3888, 64) == [ID["FAST3_1944_F64"]]
    assert ids_1d("r2c", 4099, 64) == [ID["FASTBLUE_8192_F64"]]                    # config 3c
    assert ids_1d("c2r", 4099, 64) == [ID["FASTBLUE_8192_F64"]]
    assert ids_1d("c2r", 4096, 64) == [ID["FAST3C_2048_F64"]]
    assert ids_1d("r2c", 4096, 64, np.float32) == [ID["FAST3_2048_F32"]]           # config 5 rows
    assert ids_1d("c2r", 4096, 64, np.float32) == [ID["FAST3C_2048_F32"]]
    assert ids_1d("r2c", 1024, 64) == [ID["FAST3P_512_F64"]] and ids_1d("c2r", 1024, 64) == [ID["FAST3P_512_F64"]]
    # config 4: fft2 8192 x 8192 = the two launches of the column split (axis 0 first, as general_nd), then the row kernel
    a = np.empty((8192, 8192), np.complex128)
    assert emu.nd_fast_ids("c2c", a, a, a.shape, [0, 1], True) == [ID["COL2_64_F64"], ID["COL2_128_F64"], ID["FAST3_8192_F64"]]


def test_switches_restore_the_previous_shapes(monkeypatch):
    monkeypatch.setenv("IMPULSE_FFT_R2C_PAIR", "0")
    monkeypatch.setenv("IMPULSE_FFT_C2R_PAIR", "0")
    monkeypatch.setenv("IMPULSE_FFT_F3_512P", "0")
    assert ids_1d("r2c", 1000, 64) == [ID["FAST3_500_F64"]]
    assert ids_1d("r2c", 3888, 64) == [ID["FAST3_1944_F64"]]
    assert ids_1d("c2r", 4096, 64) == [ID["FAST3_2048_F64"]]
    assert ids_1d("c2r", 2048, 64) == [ID["FAST3R_1024_F64"]]
    assert ids_1d("r2c", 1024, 64) == [ID["FAST3R_512_F64"]]
    monkeypatch.setenv("IMPULSE_FFT_NO_FAST", "1")
    assert ids_1d("c2c", 1024, 64) == [0] and ids_1d("r2c", 1000, 64) == [0]


@pytest.mark.parametrize("n", [5000, 10000, 12288])
def test_lengths_without_a_register_kernel_use_the_generic_engine(n):
    assert ids_1d("c2c", n, 16) == [0]


def test_mixed_radix_shapes_measured_in_round_2_are_on_by_default(monkeypatch):
    assert ids_1d("c2c", 1536, 16) == [ID["FAST3_1536_F64"]]
    assert ids_1d("c2c", 2000, 16) == [ID["FAST3_2000_F64"]]
    assert ids_1d("c2c", 4000, 16) == [ID["FAST3_4000_F64"]]
    assert ids_1d("c2c", 2187, 16) == [ID["FAST3_2187_F64"]] and ids_1d("c2c", 3000, 16) == [ID["FAST3_3000_F64"]]
    assert ids_1d("c2c", 6561, 16) == [ID["FAST3_6561_F64"]]
    assert ids_1d("c2c", 100, 64) == [ID["FAST2_100_F64"]] and ids_1d("c2c", 243, 64, np.float32) == [ID["FAST2_243_F32"]]
    assert ids_1d("c2c", 625, 64) == [ID["FAST2_625_F64"]]
    assert ids_1d("r2c", 4099, 16, np.float32) == [ID["FASTBLUE_8192_F32"]]
    monkeypatch.setenv("IMPULSE_FFT_MORE_SHAPES", "0")
    monkeypatch.setenv("IMPULSE_FFT_BLUE_F32", "0")
    assert ids_1d("c2c", 1536, 16) == [0] and ids_1d("r2c", 4099, 16, np.float32) == [0]


def test_split_column_transform_is_marked_for_the_fused_kernel():
    """fft2 8192 x 8192 (BASELINE config 4): the two launches of the column split are one fusable pair (colfuse2_kernel on
    the device, intermediate in an L2-resident ring); the emulation still sees — and runs — two ordinary steps."""
    a = np.empty((8192, 8192), np.complex128)
    steps = emu.nd_steps("c2c", a, a, a.shape, [0, 1], True)
    assert steps == 3
    info = emu.nd_fuse_flags("c2c", a, a, a.shape, [0, 1], True)
    assert info == [1, 0, 0], info
    b = np.empty((4, 8192, 64), np.complex64)      # strided axis of a batch: tiles = column groups x batch
    assert emu.nd_fuse_flags("c2c", b, b, b.shape, [1], True) == [1, 0]
    b = np.empty((4, 4096, 64), np.complex64)      # 64 x 64: measured slower fused, stays two launches
    assert emu.nd_fuse_flags("c2c", b, b, b.shape, [1], True) == [0, 0]
    c = np.empty((3, 2048, 32), np.complex128)     # 2048 = 32 x 64: no fused instance for that pair
    assert sum(emu.nd_fuse_flags("c2c", c, c, c.shape, [1], True)) == 0


def test_second_session_shapes_are_selected(monkeypatch):
    """tiny rows, the extra two-pass shapes, real rows on the three-pass shapes, the whole-axis kernels"""
    assert ids_1d("c2c", 4, 64) == [ID["FAST2_4_F64"]] and ids_1d("c2c", 8, 64, np.float32) == [ID["FAST2_8_F32"]]
    assert ids_1d("r2c", 16, 64) == [ID["FAST2R_8_F64"]] and ids_1d("c2r", 8, 64, np.float32) == [ID["FAST2R_4_F32"]]
    for n in (50, 72, 81, 96, 192, 200, 400, 576, 729, 900):
        assert ids_1d("c2c", n, 64) == [ID[f"FAST2_{n}_F64"]], n
        assert ids_1d("c2c", n, 64, np.float32) == [ID[f"FAST2_{n}_F32"]], n
    assert ids_1d("r2c", 6000, 64) == [ID["FAST3_3000_F64"]] and ids_1d("c2r", 4374, 64) == [ID["FAST3_2187_F64"]]
    assert ids_1d("r2c", 8000, 64, np.float32) == [ID["FAST3_4000_F32"]] and ids_1d("c2r", 13122, 64, np.float32) == [ID["FAST3_6561_F32"]]
    monkeypatch.setenv("IMPULSE_FFT_MORE_SHAPES", "0")
    assert ids_1d("c2c", 400, 64) == [0] and ids_1d("r2c", 6000, 64) == [0]        # generic engine
    monkeypatch.delenv("IMPULSE_FFT_MORE_SHAPES")
    # strided complex128 axes of 1024 / 2048 points: ONE launch on the whole-axis kernel (the device backend only)
    emu.set_fast_cols(2)
    try:
        for n, name in ((1024, "COLW_1024_F64"), (2048, "COLW_2048_F64")):
            a = np.empty((2, n, 24), np.complex128)
            assert emu.nd_fast_ids("c2c", a, a, a.shape, [1], True) == [ID[name]], n
        a = np.empty((2, 1024, 24), np.complex64)        # complex64: the two-launch split (measured faster)
        assert len(emu.nd_fast_ids("c2c", a, a, a.shape, [1], True)) == 2
    finally:
        emu.set_fast_cols(None)
