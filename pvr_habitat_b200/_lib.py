"""ctypes binding of libpvr_b200.so (include/pvr_b200.h). There is no fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpvr_b200.so")

PVR_FMT_NCHW_F32 = 0
PVR_FMT_NHWC4_BF16 = 1
PVR_FMT_STEM_BF16 = 2
PVR_OP_CONV, PVR_OP_MAXPOOL, PVR_OP_AVGPOOL, PVR_OP_HEAD = 1, 2, 3, 4


class PvrError(RuntimeError):
    pass


class pvr_op(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "kind", "in_slot", "out_slot", "res_slot", "c_in", "h_in", "w_in", "in_pitch", "c_out", "h_out", "w_out",
        "out_pitch", "res_pitch", "out_coff", "res_coff", "r", "s", "stride_h", "stride_w", "lower_h", "lower_w",
        "relu_n", "block_n", "k_pad", "n_pad", "emb_offset")] + [
        ("weight", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("aux", ctypes.c_void_p)]


class pvr_slot(ctypes.Structure):
    _fields_ = [("elems_per_image", ctypes.c_int64)]


_lib = None

_SIGNATURES = {
    "pvr_last_error": (ctypes.c_char_p, []),
    "pvr_abi_version": (ctypes.c_int, []),
    "pvr_preprocess_u8": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 9 + [
        ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_int,
        ctypes.c_int, ctypes.c_void_p]),
    "pvr_encoder_create": (ctypes.c_int, [ctypes.POINTER(pvr_op), ctypes.c_int, ctypes.POINTER(pvr_slot),
                                          ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "pvr_encoder_workspace_bytes": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int]),
    "pvr_encoder_bind": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.POINTER(ctypes.c_void_p)]),
    "pvr_encoder_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "pvr_encoder_forward_timed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                                 ctypes.POINTER(ctypes.c_float)]),
    "pvr_encoder_slot_ptr": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    "pvr_encoder_launch_count": (ctypes.c_int, [ctypes.c_void_p]),
    "pvr_encoder_destroy": (None, [ctypes.c_void_p]),
    "pvr_gemm_bf16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
}


def declared_symbols():
    """Every symbol include/pvr_b200.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the C-ABI library. Raises if it has not been built: no silent fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PvrError(
                f"{LIB_PATH} is missing: build it with `python -m pvr_habitat_b200.build` "
                "(or __graft_entry__.build()). pvr_habitat_b200 has no CPU / PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().pvr_last_error().decode("utf-8", "replace")
        raise PvrError(f"{what}: error {rc}: {msg}")


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
