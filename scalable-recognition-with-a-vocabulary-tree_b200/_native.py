"""ctypes binding of ``libvtree_b200.so`` (the C ABI declared in ``include/vtree_b200.h``).

There is no CPU fallback: if the library is missing or no sm_100 device is present, every
compute entry point raises.  The library is looked up next to this file (built in-tree by
``csrc/build.py``) and compiled on first use when nvcc is available.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvtree_b200.so")

VT_U8, VT_F32 = 0, 1
SCORE_TF_L2, SCORE_W_L2, SCORE_W_L1 = 0, 1, 2

_vp, _i32, _i64, _int = C.c_void_p, C.c_int32, C.c_int64, C.c_int

# name -> (restype, argtypes); mirrors include/vtree_b200.h declaration by declaration
SIGNATURES = {
    "vt_abi_version": (_int, []),
    "vt_last_error": (C.c_char_p, []),
    "vt_device_info": (_int, [C.POINTER(_int)] * 3),
    "vt_host_pack": (_int, [_vp, _vp, _i64, _vp, _int]),
    "vt_stage_f32_to_u8": (_int, [_vp, _i64, _int, _vp, _int, _vp, _vp]),
    "vt_stage_pad": (_int, [_vp, _i64, _int, _vp, _int, _int, _vp]),
    "vt_quantize": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _int, _int, _i32, _vp, _vp, _vp, _vp]),
    "vt_quantize_sorted_work_bytes": (_i64, [_i64]),
    "vt_quantize_sorted": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _i64, _vp]),
    "vt_quantize_sorted_profile": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "vt_quantize_mma_work_bytes": (_i64, [_i64, _int, _int]),
    "vt_quantize_mma": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _i64, _vp]),
    "vt_quantize_mma_planes_bytes": (_i64, [_int, _int]),
    "vt_quantize_mma_prepare": (_int, [_vp, _int, _int, _vp, _vp, _vp]),
    "vt_quantize_mma_prepared": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "vt_quantize_mma_profile": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "vt_quantize_mma_pair_min": (_i64, [_i64]),
    "vt_quantize_host": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _int, _int, _vp, _i64]),
    "vt_host_release": (_int, []),
    "vt_kmeans_init": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "vt_kmeans_assign": (_int, [_vp, _int, _int, _vp, _vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vt_kmeans_planes_bytes": (_i64, [_i64, _int]),
    "vt_kmeans_planes": (_int, [_vp, _i64, _int, _int, _vp, _vp, _vp]),
    "vt_kmeans_assign_mma": (_int, [_vp, _int, _int, _vp, _vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vt_kmeans_update": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "vt_column_sums": (_int, [_vp, _int, _i64, _int, _vp, _vp]),
    "vt_partition_keys": (_int, [_vp, _vp, _i64, _vp, _int, C.c_uint32, _vp, _vp]),
    "vt_sort_scratch_bytes": (_i64, [_i64]),
    "vt_sort_pairs": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp, C.POINTER(_int), _vp]),
    "vt_encode": (_int, [_vp, _vp, _i64, _i32, _int, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vt_encode_max_words": (_int, []),
    "vt_csr_compact": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "vt_posting_pack": (_int, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp]),
    "vt_posting_offsets": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "vt_score_shard_size": (_int, []),
    "vt_score_max_topk": (_int, []),
    "vt_rnorm": (_int, [_vp, _i64, _vp, _vp]),
    "vt_score_topk": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "vt_topk_merge_scored": (_int, [_vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp]),
    "vt_dense_scatter": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _int, _vp, _vp]),
    "vt_dense_dots": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _i64, _vp]),
    "vt_score_group_count": (_int, [_i64, _i64]),
    "vt_score_topk_carried": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _int, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "vt_score_topk_seeded": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _int, _vp, _i64, _vp, _vp, _vp, _vp]),
    "vt_dense_scores": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "vt_topk_rows": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "vt_idf": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "vt_weighted_rnorm": (_int, [_vp, _vp, _vp, _i64, _vp, _int, _vp, _vp]),
    "vt_query_weights": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _int, _vp, _vp, _vp]),
    "vt_score_topk_weighted_carried": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "vt_score_topk_weighted": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _vp, _vp]),
    "vt_measure_fp32_peak": (_int, [_vp, _int, _int, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp]),
    "vt_topk_merge": (_int, [_vp, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def ensure_built() -> str:
    from .csrc import build as _build
    return _build.build()


def load():
    """Load (building if needed) the native library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        ensure_built()
    except Exception as exc:                                   # no nvcc and no prebuilt library
        if not os.path.exists(LIB_PATH):
            raise NativeError("libvtree_b200.so is missing and could not be built: %s" % exc) from exc
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)                                # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    if lib.vt_abi_version() != 1:
        raise NativeError("libvtree_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().vt_last_error().decode("utf-8", "replace")
        raise NativeError("%s failed (%d): %s" % (what or "native call", rc, msg))


def ptr(t):
    """Device (or pinned host) pointer of a torch tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_device():
    """The product path never falls back to the CPU: fail loudly without a B200-class GPU."""
    import torch
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device: the vocabulary-tree path has no CPU fallback")
    lib = load()
    sm, ma, mi = _int(), _int(), _int()
    check(lib.vt_device_info(C.byref(sm), C.byref(ma), C.byref(mi)), "vt_device_info")
    return sm.value, ma.value, mi.value
