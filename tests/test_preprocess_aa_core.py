"""The arithmetic core of the antialiased bicubic resize kernel (pvr_habitat_b200/csrc/preprocess_aa_core.cuh: weights,
tap ranges, accumulation order — shared __host__ __device__ code) compiled for the host and compared bit for bit with
the oracle (oracle/restate.py), which is itself bit-identical with ATen. CPU only: the CUDA kernel around this core
(staging, indexing, output formats) is exercised by the -m gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("aa") / "libaa_core.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                    "-I", os.path.join(ROOT, "pvr_habitat_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "aa_core_harness.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.aa_max_taps.restype = ctypes.c_int
    return lib


@pytest.mark.parametrize("sizes", [(64, 224), (640, 298), (480, 224), (200, 224), (333, 224), (1000, 224), (75, 224),
                                   (225, 224), (1680, 224)])
def test_weights_and_tap_ranges_bit_exact(core, sizes):
    n_in, n_out = sizes
    taps = core.aa_max_taps()
    xmin = np.zeros(n_out, np.int32)
    size = np.zeros(n_out, np.int32)
    w = np.zeros((n_out, taps), np.float32)
    core.aa_weights(n_in, n_out, xmin.ctypes.data_as(ctypes.c_void_p), size.ctypes.data_as(ctypes.c_void_p),
                    w.ctypes.data_as(ctypes.c_void_p))
    ref = restate._aa_weights(n_in, n_out)
    assert max(len(r[1]) for r in ref) <= taps
    for i, (lo, wr) in enumerate(ref):
        assert (xmin[i], size[i]) == (lo, len(wr)), i
        assert np.array_equal(w[i, :len(wr)].view(np.uint32), wr.view(np.uint32)), i


@pytest.mark.parametrize("hw", [(64, 64), (96, 128), (100, 75), (480, 640), (300, 200), (225, 231)])
def test_two_pass_resize_bit_exact(core, hw):
    h, w = hw
    x = np.random.default_rng(h * 31 + w).integers(0, 256, (4, h, w), dtype=np.uint8)
    rh, rw, _, _ = restate.resize_geometry(h, w, 224, 224)
    out = np.zeros((4, rh, rw), np.float32)
    core.aa_resize(x.ctypes.data_as(ctypes.c_void_p), 4, h, w, rh, rw, out.ctypes.data_as(ctypes.c_void_p))
    ref = restate.resize_bicubic_aa_f32(x[None], rh, rw)[0]
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), f"{int((out != ref).sum())} values differ"
