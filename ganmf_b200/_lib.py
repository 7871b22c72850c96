"""ctypes binding of libganmf_b200.so (the C ABI declared in include/ganmf_b200.h).

There is deliberately no fallback: if the shared library is missing or cannot be loaded every
entry point raises -- the product has no CPU path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libganmf_b200.so")

KIND_GANMF, KIND_DISGANMF, KIND_MF = 0, 1, 2
ACT = {"linear": 0, None: 0, "tanh": 1, "relu": 2, "sigmoid": 3}
CSR_TRAIN, CSR_SEEN, CSR_TEST = 0, 1, 2
GEMM_AUTO, GEMM_SIMT, GEMM_TC, GEMM_TC3, GEMM_RESIDENT_A = 0, 1, 2, 3, 4
MC_NAMES = ["PRECISION", "RECALL", "PRECISION_RECALL_MIN_DEN", "MAP", "NDCG", "MRR", "ARHR", "ROC_AUC",
            "HIT_RATE", "NOVELTY", "AVERAGE_POPULARITY", "COVERED", "RMSE"]
MC_NCOL = len(MC_NAMES)
TOPK_MAX = 128


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("kind", "n_rows", "width", "num_factors", "emb_dim", "d_layers", "d_nodes", "d_act",
                 "max_batch", "item_mode", "row_id_offset", "device", "gemm_path",
                 "global_width", "item_offset", "tp_rank", "tp_world")]


_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_ctx = C.c_void_p

# name -> (restype, argtypes); must list EVERY symbol of include/ganmf_b200.h (tests check this)
SIGNATURES = {
    "ganmf_last_error": (C.c_char_p, []),
    "ganmf_create": (C.c_int, [C.POINTER(Config), C.POINTER(_ctx)]),
    "ganmf_destroy": (None, [_ctx]),
    "ganmf_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "ganmf_synchronize": (C.c_int, [_ctx]),
    "ganmf_set_csr": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f32p]),
    "ganmf_get_csr": (C.c_int, [_ctx, C.c_int, _i32p, _i32p, _i64p, _i32p, _i32p, _f32p]),
    "ganmf_set_csr_transposed": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f32p]),
    "ganmf_set_csr_device": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "ganmf_param_count": (C.c_int, [_ctx]),
    "ganmf_param_info": (C.c_int, [_ctx, C.c_int, C.c_char_p, C.c_int, _i32p, _i32p, _i32p]),
    "ganmf_set_param": (C.c_int, [_ctx, C.c_char_p, _f32p, C.c_int64]),
    "ganmf_get_param": (C.c_int, [_ctx, C.c_char_p, _f32p, C.c_int64]),
    "ganmf_init_params": (C.c_int, [_ctx, C.c_uint64]),
    "ganmf_reset_optimizers": (C.c_int, [_ctx]),
    "ganmf_snapshot": (C.c_int, [_ctx]),
    "ganmf_restore": (C.c_int, [_ctx]),
    "ganmf_upload_ids": (C.c_int, [_ctx, _i32p, C.c_int]),
    "ganmf_d_step": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "ganmf_g_step": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "ganmf_d_forward": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "ganmf_d_backward": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_float]),
    "ganmf_d_backward_phase": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_float, C.c_int]),
    "ganmf_d_apply": (C.c_int, [_ctx, C.c_float, C.c_float, C.c_int]),
    "ganmf_g_forward_backward": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_float]),
    "ganmf_g_forward_backward_part": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]),
    "ganmf_g_apply": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "ganmf_d_apply_ranges": (C.c_int, [_ctx, C.c_float, C.c_float, _i64p, _i64p, C.c_int, C.c_int]),
    "ganmf_d_forward_phase": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int]),
    "ganmf_set_gemm_sms": (C.c_int, [_ctx, C.c_int]),
    "ganmf_finalize_loss": (C.c_int, [_ctx, C.c_float, C.c_int]),
    "ganmf_tp_d_phase": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "ganmf_tp_g_phase": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "ganmf_device_buffer_ld": (C.c_int, [_ctx, C.c_char_p, _i32p]),
    "ganmf_train_epoch": (C.c_int, [_ctx, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_float, C.c_float, _f32p, _f32p]),
    "ganmf_read_losses": (C.c_int, [_ctx, _f32p, C.c_int]),
    "ganmf_device_buffer": (C.c_int, [_ctx, C.c_char_p, C.POINTER(C.c_void_p), _i64p]),
    "ganmf_score": (C.c_int, [_ctx, _i32p, C.c_int, _f32p]),
    "ganmf_encode": (C.c_int, [_ctx, _i32p, C.c_int, _f32p]),
    "ganmf_mask_topk": (C.c_int, [_ctx, _f32p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, _i32p, _f32p, C.c_int]),
    "ganmf_recommend": (C.c_int, [_ctx, _i32p, C.c_int, C.c_int, C.c_int, _i32p, _f32p, _f32p]),
    "ganmf_set_eval_tables": (C.c_int, [_ctx, _f32p, _f32p, _f32p, C.c_int, _f64p, _u8p, _f64p, C.c_int]),
    "ganmf_evaluate": (C.c_int, [_ctx, _i32p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int, _f64p, _i64p]),
    "ganmf_evaluate_values": (C.c_int, [_ctx, _i32p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int]),
    "ganmf_evaluate_sums": (C.c_int, [_ctx, _f64p, _f64p, _i64p]),
    "ganmf_eval_stats": (C.c_int, [_ctx, _i64p, _i64p]),
    "ganmf_eval_begin": (C.c_int, [_ctx, C.c_int, _i32p, C.c_int]),
    "ganmf_eval_scores_block": (C.c_int, [_ctx, _f32p, _i32p, C.c_int, C.c_int, C.c_int]),
    "ganmf_eval_end": (C.c_int, [_ctx, _f64p, _i64p]),
    "ganmf_metrics_from_topk": (C.c_int, [_ctx, _i32p, C.c_int, _i32p, C.c_int, _i32p, C.c_int, _f64p, _f64p,
                                          _i64p]),
    "ganmf_k_gemm": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "ganmf_k_csr_gather_dense": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "ganmf_k_csr_encode_rows": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "ganmf_step_routes": (C.c_int, [_ctx, _i32p, _i32p, _i32p]),
    "ganmf_k_adam": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                               C.c_float]),
    "ganmf_k_topk": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ganmf_profile": (C.c_int, [_ctx, C.c_int]),
    "ganmf_profile_read": (C.c_int, [_ctx, _f64p, _f64p, _i64p]),
    "ganmf_profile_records": (C.c_int, [_ctx, _f64p, _i32p, C.c_int, _i32p]),
    "ganmf_launch_count": (C.c_int64, [_ctx]),
}

_lib = None


def load():
    """Load the shared library (building it is __graft_entry__.build()'s / ganmf_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("ganmf_b200: %s is missing -- run `python -m ganmf_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the ABI and the binding disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GanmfError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise GanmfError(load().ganmf_last_error().decode("utf-8", "replace"))


def i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)
