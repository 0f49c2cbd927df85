"""ctypes binding of libomb200.so (include/omb200.h).  There is no CPU fallback: if the CUDA
library is missing or no CUDA device is present, importing callers get a loud error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libomb200.so")

_vp, _i64, _int, _dbl, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64

# name -> (restype, argtypes); must list every symbol declared in include/omb200.h
SIGNATURES = {
    "omb_version": (_int, []),
    "omb_last_error": (C.c_char_p, []),
    "omb_launch_count": (_i64, []),
    "omb_launch_count_reset": (None, []),
    "omb_fp64_peak": (_int, [_dbl, _vp, _i64, _vp, _vp, _vp]),
    "omb_synth_ws_bytes": (_i64, [_i64, _i64, _i64]),
    "omb_synth_fill": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _u64, _vp, _vp, _dbl, _vp, _vp]),
    "omb_row_means": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "omb_center_rows": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "omb_center_rows_padded": (_int, [_vp, _i64, _i64, _i64, _int, _vp, _vp, _vp]),
    "omb_copy_modes": (_int, [_vp, _i64, _vp, _i64, _i64, _vp]),
    "omb_block_stats_ws_bytes": (_i64, [_i64, _i64]),
    "omb_block_stats": (_int, [_vp, _i64, _i64, _int, _i64, _vp, _vp, _vp]),
    "omb_finalize_scale": (_int, [_vp, _i64, _i64, _int, _vp, _int, _vp, _i64, _vp]),
    "omb_scale_rows": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "omb_unscale": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "omb_gram_ws_bytes": (_i64, [_i64, _i64, _i64]),
    "omb_gram_rowmeans": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "omb_gram": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "omb_gram_combine": (_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "omb_eigh_max_m": (_int, []),
    "omb_eigh_jacobi": (_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "omb_backproject": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "omb_qrcp_ws_bytes": (_i64, [_i64, _i64]),
    "omb_qrcp_set_lazy": (_dbl, [_dbl]),
    "omb_qrcp_stats": (_int, [_vp, _i64, _vp, _vp]),
    "omb_qrcp": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _int, _i64, _vp, _vp, _vp, _vp]),
    "omb_qrcp_record_doubles": (_i64, []),
    "omb_qrcp_mr_start": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _int, _int, _vp]),
    "omb_qrcp_mr_local": (_int, [_vp, _vp, _i64, _i64, _vp, _int, _i64, _i64, _i64, _i64, _int, _int, _vp, _vp]),
    "omb_qrcp_mr_step": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _i64, _i64, _i64, _i64, _int, _int, _vp,
                                 _vp, _vp, _vp, _vp]),
    "omb_qrcp_p2p_error_index": (_i64, [_int]),
    "omb_qrcp_p2p_buffer_doubles": (_i64, [_int]),
    "omb_qrcp_p2p": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _int, _i64, _i64, _i64, _int, _int, _vp, _vp,
                             _i64, _vp, _vp, _vp, _vp]),
    "omb_p2p_allgather_buffer_doubles": (_i64, [_int, _i64]),
    "omb_p2p_allgather_error_index": (_i64, [_int, _i64]),
    "omb_p2p_allgather": (_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _int, _int, _vp]),
    "omb_p2p_allreduce": (_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _int, _int, _int, _vp]),
    "omb_pod_weights": (_int, [_vp, _vp, _i64, C.c_double, _vp, _vp, _vp]),
    "omb_gem_ws_bytes": (_i64, []),
    "omb_gem_max_sensors": (_int, []),
    "omb_gem_variance": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "omb_gem_step": (_int, [_vp, _i64, _i64, C.c_double, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "omb_gem_exclude": (_int, [_vp, _i64, _i64, _vp, C.c_double, _vp, _vp]),
    "omb_gem_exclude_point": (_int, [_vp, _i64, _i64, C.c_double, C.c_double, C.c_double, C.c_double, _vp, _vp]),
    "omb_wols_smem_bytes": (_i64, [_i64, _i64]),
    "omb_wols_predict": (_int, [_vp, _i64, _i64, _vp, _vp, _i64, C.c_double, _vp, _vp, _vp, _vp]),
    "omb_csr_ws_bytes": (_i64, [_i64, _i64, _i64]),
    "omb_csr_times_basis": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "omb_gather_rows": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "omb_modes_to_rows": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "omb_rows_to_modes": (_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "omb_ols_predict": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "omb_reconstruct": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
}

_lib = None


class OmbError(RuntimeError):
    pass


def load():
    """Load libomb200.so and bind every entry point (no compute is run)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OmbError(
            f"{LIB_PATH} not found: build it with `python -m openmeasure_b200.build` "
            "(there is no CPU fallback for the CUDA hot path)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = load().omb_last_error().decode(errors="replace")
        raise OmbError(f"{what or 'libomb200'} failed (rc={rc}): {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args), name)
