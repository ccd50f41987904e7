"""The core of the antialiased bicubic resize kernel (pvr_habitat_b200/csrc/preprocess_aa_core.cuh: weights, tap ranges,
accumulation order AND the kernel's work decomposition — bands, horizontal / vertical passes, intermediate layout,
frame split — all shared __host__ __device__ code) compiled for the host and compared bit for bit with the oracle
(oracle/restate.py), which is itself bit-identical with ATen. CPU only: what the CUDA kernel adds around this core
(shared-memory staging by bulk copy, the bf16 / NHWC stores, the launch) is exercised by the -m gpu tests."""
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


@pytest.mark.parametrize("hw,nf,n,rows,sample_major", [((64, 64), 1, 3, 16, 0), ((64, 64), 2, 2, 16, 1),
                                                        ((96, 128), 1, 2, 8, 0), ((100, 75), 2, 1, 16, 1),
                                                        ((480, 640), 1, 1, 4, 0), ((300, 200), 3, 1, 5, 1)])
def test_kernel_decomposition_bit_exact_vs_oracle_transforms(core, hw, nf, n, rows, sample_major):
    """Bands, tap-range bookkeeping, the float32 intermediate layout, frame split, crop offsets, clamp / round / LUT —
    the kernel's own host/device functions driven block by block — against the oracle's CLIP transforms."""
    from oracle import restate_vit
    from pvr_habitat_b200.embeddings import CLIP_MEAN, CLIP_STD, resize_geometry
    h, w = hw
    obs = np.random.default_rng(h * 3 + w + nf).integers(0, 256, (n, h, w, 3 * nf), dtype=np.uint8)
    rh, rw, top, left = resize_geometry(h, w, 224, 224)
    out = np.full((nf * n, 3, 224, 224), np.nan, np.float32)
    mean = (ctypes.c_float * 3)(*CLIP_MEAN)
    std = (ctypes.c_float * 3)(*CLIP_STD)
    core.aa_preprocess(obs.ctypes.data_as(ctypes.c_void_p), n, h, w, nf, rh, rw, top, left, 224, mean, std,
                       out.ctypes.data_as(ctypes.c_void_p), sample_major, rows)
    frames, _ = restate.split_frames(obs)                      # frame-major: image f*N + i
    want = restate_vit.clip_transforms(frames)
    if sample_major:                                           # image i*nf + f
        want = want.reshape(nf, n, 3, 224, 224).transpose(1, 0, 2, 3, 4).reshape(nf * n, 3, 224, 224)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), f"{int((out != want).sum())} values differ"


@pytest.mark.parametrize("H,rows", [(64, 16), (480, 4), (480, 16), (1000, 2), (225, 16), (100, 16), (1680, 1)])
def test_shared_memory_row_bound_covers_every_band(core, H, rows):
    """pvr_preprocess_u8_aa sizes shared memory for ceil(scale * rows) + taps + 1 input rows per band (csrc/
    preprocess_aa.cu); the kernel traps if a band needs more."""
    import math
    rh = 224
    scale = np.float32(H) / np.float32(rh)
    support = np.float32(2) * scale if scale >= 1 else np.float32(2)
    taps = int(math.ceil(support)) * 2 + 1
    bound = int(math.ceil(float(scale) * rows)) + taps + 1
    assert core.aa_max_rows_in(H, rh, 0, 224, rows) <= bound
