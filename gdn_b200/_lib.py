"""ctypes binding of libgdn_b200.so (include/gdn_b200.h).  No torch types cross the ABI:
only raw device pointers, ints and POD structs.  There is no CPU fallback: a missing
library or a non-zero return code raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgdn_b200.so")

c_fp = C.c_void_p  # device pointers travel as integers


class Dims(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("W", C.c_int), ("D", C.c_int), ("K", C.c_int)]


class LayerParams(C.Structure):
    _fields_ = [(n, c_fp) for n in ("lin_weight", "att_i", "att_j", "att_em_i", "att_em_j", "bias")]


class LayerGrads(C.Structure):
    _fields_ = [(n, c_fp) for n in ("lin_weight", "att_i", "att_j", "att_em_i", "att_em_j", "bias", "embedding")]


class BN(C.Structure):
    _fields_ = [(n, c_fp) for n in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]


class HeadParams(C.Structure):
    _fields_ = [("bn1", BN), ("bn2", BN), ("out_w", c_fp), ("out_b", c_fp)]


class HeadGrads(C.Structure):
    _fields_ = [(n, c_fp) for n in ("bn1_weight", "bn1_bias", "bn2_weight", "bn2_bias", "out_w", "out_b")]


class Dropout(C.Structure):
    _fields_ = [("mask", c_fp), ("seed", C.c_uint64), ("offset", C.c_uint64), ("p", C.c_float),
                ("offset_dev", c_fp)]


SYNC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p)


class Sync(C.Structure):
    _fields_ = [("world", C.c_int), ("allreduce_sum_f64", SYNC_FN), ("user", C.c_void_p)]


# name -> (restype, argtypes); must list every symbol include/gdn_b200.h declares
_P = C.POINTER
SIGNATURES = {
    "gdn_version": (C.c_int, []),
    "gdn_last_error": (C.c_char_p, []),
    "gdn_launch_count": (C.c_longlong, []),
    "gdn_stage_f64_to_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "gdn_profile_enable": (C.c_int, [C.c_int]),
    "gdn_profile_collect": (C.c_int, [C.c_char_p, C.c_size_t]),
    "gdn_graph_build_ws_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gdn_graph_build": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_size_t, C.c_int, c_fp]),
    "gdn_graph_build_warm": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_size_t, C.c_int, c_fp,
                                       C.c_float, c_fp]),
    "gdn_graph_build_rows": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_size_t,
                                       C.c_int, c_fp, C.c_float, c_fp]),
    "gdn_graphlayer_ctx_bytes": (C.c_size_t, [_P(Dims)]),
    "gdn_graphlayer_ws_bytes": (C.c_size_t, [_P(Dims)]),
    "gdn_graphlayer_fwd": (C.c_int, [_P(Dims), c_fp, c_fp, c_fp, _P(LayerParams), c_fp, c_fp, c_fp, c_fp,
                                     C.c_size_t, c_fp]),
    "gdn_graphlayer_bwd": (C.c_int, [_P(Dims), c_fp, c_fp, c_fp, _P(LayerParams), c_fp, _P(LayerGrads), c_fp,
                                     C.c_size_t, c_fp]),
    "gdn_csr_fwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64] + [c_fp] * 14 + [C.c_float, c_fp]),
    "gdn_csr_bwd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64] + [c_fp] * 17 + [C.c_float, c_fp]),
    "gdn_fused_ctx_bytes": (C.c_size_t, [_P(Dims)]),
    "gdn_fused_ws_bytes": (C.c_size_t, [_P(Dims)]),
    "gdn_fused_fwd": (C.c_int, [_P(Dims), c_fp, c_fp, c_fp, _P(LayerParams), _P(HeadParams), _P(Dropout), C.c_int,
                                c_fp, c_fp, c_fp, C.c_size_t, c_fp]),
    "gdn_fused_bwd": (C.c_int, [_P(Dims), c_fp, c_fp, c_fp, _P(LayerParams), _P(HeadParams), _P(Dropout), c_fp,
                                _P(LayerGrads), _P(HeadGrads), c_fp, C.c_size_t, c_fp]),
    "gdn_fused_fwd_sync": (C.c_int, [_P(Dims), c_fp, c_fp, c_fp, _P(LayerParams), _P(HeadParams), _P(Dropout), C.c_int,
                                     c_fp, c_fp, c_fp, C.c_size_t, _P(Sync), c_fp]),
    "gdn_fused_bwd_sync": (C.c_int, [_P(Dims), c_fp, c_fp, c_fp, _P(LayerParams), _P(HeadParams), _P(Dropout), c_fp,
                                     _P(LayerGrads), _P(HeadGrads), c_fp, C.c_size_t, _P(Sync), c_fp]),
    "gdn_ctx_alpha": (C.c_int, [_P(Dims), c_fp, c_fp, c_fp, c_fp]),
    "gdn_score_ws_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gdn_score": (C.c_int, [c_fp, c_fp, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, C.c_size_t, c_fp]),
    "gdn_adam_flat": (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_longlong, C.c_float, c_fp]),
    "gdn_adam_flat_dev": (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                    c_fp, C.c_float, c_fp]),
    "gdn_nvls_adam": (C.c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, C.c_longlong, C.c_longlong, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_longlong, C.c_float, c_fp]),
    "gdn_f1_sweep": (C.c_int, [c_fp, c_fp, C.c_int, c_fp, c_fp, C.c_int, c_fp, c_fp, c_fp]),
    "gdn_binary_counts": (C.c_int, [c_fp, c_fp, C.c_int, C.c_double, c_fp, c_fp]),
    "gdn_auc_ranksum": (C.c_int, [c_fp, c_fp, C.c_int, c_fp, c_fp, c_fp]),
    "gdn_window_batch": (C.c_int, [c_fp, c_fp, C.c_int, C.c_int, C.c_int, c_fp, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m gdn_b200.build` "
                "(gdn_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().gdn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
