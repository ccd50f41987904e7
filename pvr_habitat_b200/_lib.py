"""ctypes binding of libpvr_b200.so (include/pvr_b200.h). There is no fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpvr_b200.so")

PVR_FMT_NCHW_F32 = 0
PVR_FMT_NHWC4_BF16 = 1
PVR_FMT_STEM_BF16 = 2
PVR_FMT_NHWC4_F32 = 3
PVR_FMT_STEM_PAD_BF16 = 4
PVR_RESIZE_BICUBIC = 0x100
PVR_RESIZE_FLOAT = 0x200
PVR_SWAP_ROWS_0_2 = 0x400
PVR_COMM_F32, PVR_COMM_F64, PVR_COMM_BF16, PVR_COMM_I64 = 0, 1, 2, 3
PVR_OP_FP32 = 2
PVR_CONV_OUT_F32 = 1
PVR_GEMM_PDL, PVR_GEMM_MN = 1, 2
PVR_LSTM_CONT_PREV, PVR_LSTM_CONT_NEXT = 1, 2
PVR_OP_CONV, PVR_OP_MAXPOOL, PVR_OP_AVGPOOL, PVR_OP_HEAD, PVR_OP_FLATTEN, PVR_OP_AVGPOOL2 = 1, 2, 3, 4, 5, 6


class PvrError(RuntimeError):
    pass


class pvr_op(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "kind", "in_slot", "out_slot", "res_slot", "c_in", "h_in", "w_in", "in_pitch", "c_out", "h_out", "w_out",
        "out_pitch", "res_pitch", "out_coff", "res_coff", "r", "s", "stride_h", "stride_w", "lower_h", "lower_w",
        "relu_n", "block_n", "k_pad", "n_pad", "emb_offset", "act", "in2_slot", "in2_c", "in2_h", "in2_w", "in2_pitch",
        "in2_stride", "flags")] + [
        ("weight", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("aux", ctypes.c_void_p)]


class pvr_gemm_desc(ctypes.Structure):
    _fields_ = [("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("out", ctypes.c_void_p), ("scale", ctypes.c_void_p),
                ("bias", ctypes.c_void_p), ("res", ctypes.c_void_p), ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64),
                ("ldo", ctypes.c_int64), ("ldr", ctypes.c_int64), ("m", ctypes.c_int32), ("n", ctypes.c_int32),
                ("n_pad", ctypes.c_int32), ("k", ctypes.c_int32), ("relu", ctypes.c_int32),
                ("res_mode", ctypes.c_int32), ("out_f32", ctypes.c_int32), ("split_k", ctypes.c_int32),
                ("act", ctypes.c_int32), ("flags", ctypes.c_int32)]


class pvr_lstm_fwd(ctypes.Structure):
    _fields_ = [("T", ctypes.c_int32), ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("flags", ctypes.c_int32)] + [
        (n, ctypes.c_void_p) for n in ("w_hh", "xp", "nd", "h0", "c_all", "hm", "h_out", "gates", "g_tmp", "h_last",
                                       "counters")] + [("counters_bytes", ctypes.c_int64)]


class pvr_lstm_bwd(ctypes.Structure):
    _fields_ = [("T", ctypes.c_int32), ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("flags", ctypes.c_int32)] + [
        (n, ctypes.c_void_p) for n in ("w_hh_t", "nd", "gates", "c_all", "dh_out", "dh_rec", "dc_rec", "dG", "dbias",
                                       "counters")] + [("counters_bytes", ctypes.c_int64)]


def lstm_counter_bytes(T, B):
    """PVR_LSTM_COUNTER_BYTES of include/pvr_b200.h"""
    return (T + 1) * ((B + 31) // 32) * 16 * 4


class pvr_slot(ctypes.Structure):
    _fields_ = [("elems_per_image", ctypes.c_int64)]


_lib = None
_vp, _i, _i64, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

_SIGNATURES = {
    "pvr_last_error": (ctypes.c_char_p, []),
    "pvr_abi_version": (ctypes.c_int, []),
    "pvr_preprocess_u8": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 9 + [
        ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_int,
        ctypes.c_int, ctypes.c_void_p]),
    "pvr_preprocess_u8_aa": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 9 + [
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
    "pvr_gemm": (ctypes.c_int, [ctypes.POINTER(pvr_gemm_desc), ctypes.c_void_p]),
    "pvr_layernorm": (ctypes.c_int, [_vp, _i64, _i64, _i, _vp, _vp, _f, _vp, _vp]),
    "pvr_layernorm_f32": (ctypes.c_int, [_vp, _i64, _i64, _i, _vp, _vp, _f, _vp, _i64, _vp]),
    "pvr_vit_embed": (ctypes.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _f, _vp, _vp]),
    "pvr_attention": (ctypes.c_int, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "pvr_attention_mma": (ctypes.c_int, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "pvr_attnpool_tokens": (ctypes.c_int, [_vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "pvr_vit_patchify": (ctypes.c_int, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "pvr_bn1d_stats": (ctypes.c_int, [_vp, _i64, _i, _i, _vp, _vp]),
    "pvr_bn1d_normalize": (ctypes.c_int, [_vp, _i64, _i, _i, _vp, ctypes.c_double, _f, _f, _vp, _vp, _vp, _vp, _vp,
                                          _vp, _vp, _i64, _vp]),
    "pvr_bn1d_eval": (ctypes.c_int, [_vp, _i64, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pvr_bn1d_backward": (ctypes.c_int, [_vp, _i64, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pvr_cast_rows_bf16": (ctypes.c_int, [_vp, _i64, _i, _i, _vp, _i64, _vp]),
    "pvr_lstm_cell_forward": (ctypes.c_int, [_vp] * 5 + [_i, _i] + [_vp] * 6),
    "pvr_lstm_cell_backward": (ctypes.c_int, [_vp] * 8 + [_i, _i, _vp, _vp]),
    "pvr_lstm_forward": (ctypes.c_int, [ctypes.POINTER(pvr_lstm_fwd), _vp]),
    "pvr_lstm_backward": (ctypes.c_int, [ctypes.POINTER(pvr_lstm_bwd), _vp]),
    "pvr_lstm_persist_supported": (ctypes.c_int, [_i, _i, _i]),
    "pvr_lstm_persist_profile": (ctypes.c_int, [_vp]),
    "pvr_heads_forward": (ctypes.c_int, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "pvr_heads_backward": (ctypes.c_int, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "pvr_ce_loss": (ctypes.c_int, [_vp, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "pvr_colsum_bf16": (ctypes.c_int, [_vp, _i64, _i, _i, _vp, _vp]),
    "pvr_transpose_bf16": (ctypes.c_int, [_vp, _i64, _i, _i, _vp, _i64, _vp]),
    "pvr_cast_weight": (ctypes.c_int, [_vp, _i, _i, _vp, _i64, _vp, _i64, _vp]),
    "pvr_convfeat_gather": (ctypes.c_int, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pvr_convfeat_scatter": (ctypes.c_int, [_vp, _i64, _i, _i, _i, _i, _i, _vp, _vp]),
    "pvr_elu_backward": (ctypes.c_int, [_vp, _vp, _i, _i64, _i, _vp, _vp]),
    "pvr_elu_backward_fused": (ctypes.c_int, [_vp, _vp, _i, _i64, _i, _vp, _vp, _i64, _vp, _vp]),
    "pvr_im2col_t": (ctypes.c_int, [_vp, _i, _i, _i, _i, _i, _i, _i, _i64, _vp, _vp]),
    "pvr_col2im": (ctypes.c_int, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pvr_small_conv1_wgrad": (ctypes.c_int, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pvr_bn1d_backward_dx": (ctypes.c_int, [_vp, _i64, _vp, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_double,
                                            _vp, _i64, _vp]),
    "pvr_bf16_rows_to_f32": (ctypes.c_int, [_vp, _i64, _i64, _i, _vp, _i64, _vp]),
    "pvr_optim_sumsq": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.POINTER(_i64), _i, _vp, _vp]),
    "pvr_optim_step": (ctypes.c_int, [_i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                      ctypes.POINTER(_vp), ctypes.POINTER(_i64), _i, _vp, _f, _f, _f, _f, _f, _f, _i,
                                      _vp, _vp]),
    "pvr_optim_step_dev": (ctypes.c_int, [_i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                          ctypes.POINTER(_vp), ctypes.POINTER(_i64), _i, _vp, _f, _f, _vp, _f, _f, _f, _i,
                                          _vp, _vp]),
    "pvr_gemm_f32": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp]),
    "pvr_vit_embed_f32": (ctypes.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _f, _vp, _vp]),
    "pvr_attention_f32": (ctypes.c_int, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "pvr_comm_load": (ctypes.c_int, [ctypes.c_char_p]),
    "pvr_comm_version": (ctypes.c_int, []),
    "pvr_comm_unique_id": (ctypes.c_int, [_vp]),
    "pvr_comm_init": (ctypes.c_int, [_i, _i, _vp, ctypes.POINTER(ctypes.c_void_p)]),
    "pvr_comm_destroy": (ctypes.c_int, [_vp]),
    "pvr_comm_allreduce": (ctypes.c_int, [_vp, _vp, _i64, _i, _vp]),
    "pvr_comm_reduce_scatter": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i, _vp]),
    "pvr_comm_allgather": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i, _vp]),
    "pvr_comm_broadcast": (ctypes.c_int, [_vp, _vp, _i64, _i, _i, _vp]),
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
