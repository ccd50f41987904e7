"""The C-ABI library loads and exports every symbol include/pvr_b200.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

import pytest

from pvr_habitat_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pvr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pvr_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    return _lib.lib()


def test_header_and_binding_agree():
    assert header_symbols() == _lib.declared_symbols()


def test_library_exports_every_declared_symbol(built):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(raw, name), name


def test_abi_version_and_error_string(built):
    assert built.pvr_abi_version() == 3
    assert isinstance(built.pvr_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu(built):
    rc = built.pvr_preprocess_u8(None, 1, 64, 64, 1, 256, 256, 16, 16, 224, None, None, None, 0, 0, None)
    assert rc == -1 and b"invalid argument" in built.pvr_last_error()
    h = ctypes.c_void_p()
    assert built.pvr_encoder_create(None, 0, None, 0, 0, ctypes.byref(h)) == -1
    assert built.pvr_encoder_workspace_bytes(None, 4) < 0
    assert built.pvr_gemm_bf16(None, 0, None, 0, None, 0, None, None, None, 0, 1, 1, 64, 64, 0, None) == -1


def test_op_struct_layout_matches_header():
    """pvr_op: 34 int32 fields then 4 pointers (ctypes mirror of the C struct)."""
    assert ctypes.sizeof(_lib.pvr_op) == 34 * 4 + 4 * 8
    assert _lib.pvr_op.weight.offset == 136


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    offenders = []
    for base, _, files in os.walk(os.path.join(root, "pvr_habitat_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(base, f), errors="ignore") as fh:
                    if pat.search(fh.read()):
                        offenders.append(os.path.join(base, f))
    assert offenders == []
    # the GPU arm of bench.py reaches the oracle only through its CPU legs
    src = open(os.path.join(root, "bench.py")).read()
    for m in pat.finditer(src):
        func = src[:m.start()].rsplit("\ndef ", 1)[1].split("(", 1)[0]
        assert func in ("oracle_parts", "oracle_embed", "cpu_port_frames_per_s", "cpu_port_bc_steps_per_s",
                        "cpu_port_finetune_steps_per_s", "run_reference"), func


def test_bench_clock_sampler_counts_only_samples_of_the_timed_region():
    """bench.py `clocks`: nvidia-smi samples read outside the timed region are idle (no power cap, maximum SM clock)
    and must not enter the median; with no sample inside the window everything read is used."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    def row(sm, cap):
        return [str(sm), "1965", "700.0", "Not Active", "Not Active", "Not Active", "Active" if cap else "Not Active"]

    rows = [(0.00, row(1965, False)), (0.06, row(1965, False)), (0.12, row(1650, True)), (0.17, row(1620, True)),
            (0.22, row(1630, True)), (0.30, row(1640, True)), (0.36, row(1965, False)), (0.41, ["garbage"])]
    got = bench.ClockSampler.summarise(rows, 0.05, 0.28)
    assert got == {"sm_mhz": 1635.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 4,
                   "samples_in_timed_region": 4}
    assert bench.ClockSampler.summarise(rows, None, None)["samples"] == 7
    assert bench.ClockSampler.summarise(rows, 10.0, 11.0)["samples_in_timed_region"] == 0
