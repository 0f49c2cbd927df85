"""ctypes bindings for oracle/liboracle.so.  TEST INFRASTRUCTURE ONLY (see oracle/csrc/oracle.c)."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int64)
        L.omo_pairwise_sum.restype = C.c_double
        L.omo_pairwise_sum.argtypes = [dp, C.c_int64]
        L.omo_block_stats.restype = None
        L.omo_block_stats.argtypes = [dp, C.c_int64, dp, dp, dp, dp, dp]
        L.omo_row_means.restype = None
        L.omo_row_means.argtypes = [dp, C.c_int64, C.c_int64, dp]
        L.omo_qrcp_dlaqp2.restype = C.c_int
        L.omo_qrcp_dlaqp2.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, ip, dp, dp, ip]
        L.omo_synth_u01.restype = C.c_double
        L.omo_synth_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.omo_synth_fill.restype = None
        L.omo_synth_fill.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                     C.c_int64, C.c_uint64, dp, dp, C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def pairwise_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    return float(lib().omo_pairwise_sum(_dp(a), a.size))


def block_stats(a):
    """(sum, mean, sum of squared deviations, min, max) of a contiguous block, numpy's tree."""
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    out = [C.c_double() for _ in range(5)]
    lib().omo_block_stats(_dp(a), a.size, *[C.byref(o) for o in out])
    return tuple(o.value for o in out)


def row_means(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(x.shape[0])
    lib().omo_row_means(_dp(x), x.shape[0], x.shape[1], _dp(out))
    return out


def qrcp_dlaqp2(Ur, nsteps=None):
    """LAPACK dlaqp2 pivots of Ur.T (Ur is (n, r) C-order).  Returns dict(piv, rdiag, gap, nrecomp)."""
    A = np.array(Ur, dtype=np.float64, order="C", copy=True)
    n, r = A.shape
    k = min(n, r) if nsteps is None else int(nsteps)
    jpvt = np.empty(n, dtype=np.int64)
    rdiag = np.zeros(k)
    gap = np.zeros(k)
    nre = np.zeros(k, dtype=np.int64)
    rc = lib().omo_qrcp_dlaqp2(_dp(A), r, n, k, _ip(jpvt), _dp(rdiag), _dp(gap), _ip(nre))
    if rc != 0:
        raise MemoryError("omo_qrcp_dlaqp2 failed")
    return dict(piv=jpvt[:k].copy(), rdiag=rdiag, gap=gap, nrecomp=nre)
