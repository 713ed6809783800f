"""ctypes binding of librsoptsc_b200.so (the C ABI declared in include/rsoptsc_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
compute entry point is called, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librsoptsc_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
# int (*rsoptsc_progress_fn)(int32_t stage, int32_t iter, double value, void* user)
PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_int32, C.c_int32, C.c_double, C.c_void_p)

# name -> (restype, argtypes); mirrors include/rsoptsc_b200.h one to one
SIGNATURES = {
    "rsoptsc_abi_version": (C.c_int, []),
    "rsoptsc_init": (C.c_int, [C.c_int]),
    "rsoptsc_finalize": (C.c_int, []),
    "rsoptsc_last_error": (C.c_char_p, []),
    "rsoptsc_kernel_launches": (C.c_int64, []),
    "rsoptsc_phase_ms": (C.c_double, [C.c_char_p]),
    "rsoptsc_phase_calls": (C.c_int64, [C.c_char_p]),
    "rsoptsc_timers_reset": (C.c_int, []),
    "rsoptsc_timers_enable": (C.c_int, [C.c_int]),
    "rsoptsc_phase_name": (C.c_char_p, [C.c_int32]),
    "rsoptsc_counter": (C.c_double, [C.c_char_p]),
    "rsoptsc_timer_begin": (C.c_int, [C.c_char_p]),
    "rsoptsc_timer_end": (C.c_int, [C.c_char_p]),
    "rsoptsc_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "rsoptsc_comm_init": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "rsoptsc_comm_finalize": (C.c_int, []),
    "rsoptsc_comm_info": (C.c_int, [c_int32_p, c_int32_p, c_int32_p, c_int64_p, c_double_p]),
    "rsoptsc_dev_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "rsoptsc_dev_free": (C.c_int, [C.c_void_p]),
    "rsoptsc_dev_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "rsoptsc_dev_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "rsoptsc_dev_sync": (C.c_int, []),
    "rsoptsc_normalize": (C.c_int, [c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "rsoptsc_normalize_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "rsoptsc_pca_spectrum": (C.c_int, [c_double_p, C.c_int64, C.c_int64, c_double_p, c_int32_p, c_double_p, C.c_int32]),
    "rsoptsc_knn": (C.c_int, [c_double_p, C.c_int64, C.c_int32, C.c_int32, c_int32_p]),
    "rsoptsc_computeM": (C.c_int, [c_double_p, C.c_int64, C.c_int64, c_int32_p, C.c_int32, C.c_double, c_double_p,
                                   c_double_p, c_int32_p, c_int32_p, c_double_p]),
    "rsoptsc_computeM_dense_mask": (C.c_int, [c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_double, c_double_p,
                                              c_double_p, c_int32_p, c_int32_p, c_double_p]),
    "rsoptsc_threshold_symmetrise": (C.c_int, [c_double_p, C.c_int64, c_double_p]),
    "rsoptsc_threshold_symmetrise_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "rsoptsc_similarity": (C.c_int, [c_double_p, C.c_int64, C.c_int64, c_double_p, C.c_int32, C.c_double, C.c_int32,
                                     c_double_p, c_int64_p, c_int32_p, c_int32_p, c_double_p, c_double_p, c_int32_p,
                                     c_int32_p, c_double_p]),
    "rsoptsc_similarity_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_double,
                                         C.c_int32, C.c_void_p, c_double_p, c_int32_p, c_int32_p]),
    "rsoptsc_similarity_knn": (C.c_int, [c_double_p, C.c_int64, C.c_int64, c_int32_p, C.c_int32, C.c_double, c_double_p,
                                         c_double_p, c_int32_p]),
    "rsoptsc_set_progress": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rsoptsc_nmf_lee": (C.c_int, [c_double_p, C.c_int64, C.c_int32, C.c_int32, c_double_p, c_double_p, c_int32_p,
                                  c_int32_p]),
    "rsoptsc_nmf_lee_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, c_double_p, c_double_p, c_int32_p,
                                      c_int32_p]),
    "rsoptsc_nmf_sweep": (C.c_int, [c_double_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_int32_p, c_int32_p]),
    "rsoptsc_nmf_sweep_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_int32_p, c_int32_p]),
    "rsoptsc_get_components": (C.c_int, [c_double_p, C.c_int64, C.c_double, c_double_p, c_int32_p]),
    "rsoptsc_consensus": (C.c_int, [c_int32_p, C.c_int32, C.c_int64, C.c_double, c_double_p]),
    "rsoptsc_get_ensemble": (C.c_int, [c_double_p, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int32, c_double_p]),
    "rsoptsc_pam": (C.c_int, [c_double_p, C.c_int64, C.c_int32, C.c_int32, c_int32_p, c_int32_p]),
    "rsoptsc_count_clusters": (C.c_int, [c_double_p, C.c_int64, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                                         c_int32_p, c_int32_p, c_double_p, c_int32_p]),
    "rsoptsc_dist_flat": (C.c_int, [c_double_p, C.c_int64, C.c_int32, c_double_p]),
    "rsoptsc_adjacency": (C.c_int, [c_double_p, C.c_int64, c_double_p]),
    "rsoptsc_graph_components": (C.c_int, [c_double_p, C.c_int64, c_int32_p, c_int32_p]),
    "rsoptsc_graph_distances": (C.c_int, [c_double_p, C.c_int64, c_double_p]),
    "rsoptsc_log_normalize": (C.c_int, [c_double_p, C.c_int64, C.c_int64, C.c_double, C.c_int32, c_double_p]),
    "rsoptsc_scale_center": (C.c_int, [c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "rsoptsc_count_threshold": (C.c_int, [c_double_p, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                          c_int32_p]),
    "rsoptsc_select_data": (C.c_int, [c_double_p, C.c_int64, C.c_int64, C.c_double, C.c_int32, c_int32_p, c_int32_p,
                                      c_int32_p]),
    "rsoptsc_signaling_pair": (C.c_int, [c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int32,
                                         c_double_p]),
    "rsoptsc_set_gemm_backend": (C.c_int, [C.c_int32]),
    "rsoptsc_get_gemm_backend": (C.c_int32, []),
    "rsoptsc_debug_gemm": (C.c_int, [C.c_int32, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int64, C.c_int64,
                                     C.c_int32, C.c_int32, c_double_p]),
    "rsoptsc_debug_sym_eig": (C.c_int, [c_double_p, C.c_int64, C.c_int32, c_double_p, c_double_p, C.c_int32, c_double_p]),
    "rsoptsc_host_tridiag_eig": (C.c_int, [C.c_int32, c_double_p, c_double_p, c_double_p]),
    "rsoptsc_host_tridiag_bisect": (C.c_int, [C.c_int32, c_double_p, c_double_p, c_double_p]),
    "rsoptsc_host_choose_K": (C.c_int32, [c_double_p, C.c_int64]),
    "rsoptsc_host_mg_tables": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_int64_p, c_int32_p]),
    "rsoptsc_host_shard_owner": (C.c_int, [c_int32_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64, c_int32_p, c_int32_p,
                                           c_int32_p]),
}


class RsoptscError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (no CUDA call is made).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RsoptscError(
            "librsoptsc_b200.so is missing (%s): build it with `python -m rsoptsc_b200.build`; "
            "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().rsoptsc_last_error()
        raise RsoptscError("rsoptsc_b200 error %d: %s" % (status, msg.decode() if msg else "?"))


_initialized_device = None


def init(device=None):
    """Bind this process to one GPU (one process per GPU)."""
    global _initialized_device
    lib = load()
    dev = -1 if device is None else int(device)
    if _initialized_device is not None and (dev < 0 or dev == _initialized_device):
        return lib
    check(lib.rsoptsc_init(dev))
    _initialized_device = dev if dev >= 0 else 0
    return lib
