"""ctypes mirror of include/speedy_eval.h: the reference's evaluation tools.

DynamicTimeWarping (/root/reference/dynamic_time_warping.h:28-118), the Teager
energy statistics and the path slopes of /root/reference/sonic_test.cc:85-209,
used to re-run the reference's statistical tests against the CUDA path's output.
Host-only; no CUDA needed.
"""
import ctypes as C
import os

import numpy as np

EVAL_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libspeedy_eval.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(EVAL_LIB_PATH):
            raise RuntimeError("libspeedy_eval.so is missing: run `python -m speedy_b200.build`")
        L = C.CDLL(EVAL_LIB_PATH)
        fp, ip, sp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_short)
        L.speedyEvalDtw.restype = C.c_float
        L.speedyEvalDtw.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, ip, ip, ip]
        L.speedyEvalTeagerVarianceShort.argtypes = [sp, C.c_int, fp, fp]
        L.speedyEvalTeagerVarianceFloat.argtypes = [fp, C.c_int, fp, fp]
        L.speedyEvalTeagerShort.argtypes = [sp, C.c_int, fp]
        L.speedyEvalTeagerOutlierCountShort.argtypes = [sp, C.c_int, C.c_float]
        L.speedyEvalLinearSlopeInt.restype = C.c_float
        L.speedyEvalLinearSlopeInt.argtypes = [ip, ip, C.c_int]
        L.speedyEvalLinearSlopeEverywhereInt.argtypes = [ip, ip, C.c_int, C.c_int, fp]
        L.speedyEvalMean.restype = C.c_float
        L.speedyEvalMean.argtypes = [fp, C.c_int]
        L.speedyEvalStandardDeviation.restype = C.c_float
        L.speedyEvalStandardDeviation.argtypes = [fp, C.c_int]
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def dtw(seq1, seq2):
    """Optimal cost and warping path (path1, path2) between two [len][dim] sequences."""
    a, pa = _f(np.atleast_2d(seq1))
    b, pb = _f(np.atleast_2d(seq2))
    assert a.shape[1] == b.shape[1], "points must have the same dimensionality"
    p1 = np.zeros(a.shape[0] + b.shape[0], np.int32)
    p2 = np.zeros_like(p1)
    n = C.c_int(0)
    cost = lib().speedyEvalDtw(pa, a.shape[0], pb, b.shape[0], a.shape[1], p1.ctypes.data_as(C.POINTER(C.c_int)),
                               p2.ctypes.data_as(C.POINTER(C.c_int)), C.byref(n))
    return float(cost), p1[:n.value].copy(), p2[:n.value].copy()


def teager_variance(x):
    """(mean, variance) of the Teager energy of an int16 or float32 signal."""
    m, v = C.c_float(0), C.c_float(0)
    x = np.ascontiguousarray(x)
    if x.dtype == np.int16:
        lib().speedyEvalTeagerVarianceShort(x.ctypes.data_as(C.POINTER(C.c_short)), len(x), C.byref(m), C.byref(v))
    else:
        x, px = _f(x)
        lib().speedyEvalTeagerVarianceFloat(px, len(x), C.byref(m), C.byref(v))
    return m.value, v.value


def teager(x):
    x = np.ascontiguousarray(x, dtype=np.int16)
    out = np.zeros(max(len(x) - 2, 0), np.float32)
    n = lib().speedyEvalTeagerShort(x.ctypes.data_as(C.POINTER(C.c_short)), len(x), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out[:n]


def teager_outlier_count(x, thresh_fraction):
    x = np.ascontiguousarray(x, dtype=np.int16)
    return lib().speedyEvalTeagerOutlierCountShort(x.ctypes.data_as(C.POINTER(C.c_short)), len(x), thresh_fraction)


def linear_slope(x, y):
    x, px = _i(x)
    y, py = _i(y)
    assert len(x) == len(y)
    return float(lib().speedyEvalLinearSlopeInt(px, py, len(x)))


def linear_slope_everywhere(x, y, half_width):
    x, px = _i(x)
    y, py = _i(y)
    assert len(x) == len(y)
    out = np.zeros(max(len(x) - 2 * half_width, 0), np.float32)
    n = lib().speedyEvalLinearSlopeEverywhereInt(px, py, len(x), half_width, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out[:n]


def mean(v):
    v, pv = _f(v)
    return float(lib().speedyEvalMean(pv, len(v)))


def standard_deviation(v):
    v, pv = _f(v)
    return float(lib().speedyEvalStandardDeviation(pv, len(v)))
