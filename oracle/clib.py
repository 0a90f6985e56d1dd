"""ctypes binding of oracle/c/asr_oracle.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libasr_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def topk(Q, D, k, normalise=True, idx_base=0):
    Q = np.ascontiguousarray(Q, np.float32)
    D = np.ascontiguousarray(D, np.float32)
    out_s = np.empty((Q.shape[0], k), np.float32)
    out_i = np.empty((Q.shape[0], k), np.int64)
    lib().asr_oracle_topk(_f(Q), ctypes.c_long(Q.shape[0]), _f(D), ctypes.c_long(D.shape[0]),
                          ctypes.c_int(Q.shape[1]), ctypes.c_int(k), ctypes.c_int(int(normalise)),
                          ctypes.c_longlong(idx_base), _f(out_s),
                          out_i.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    return out_s, out_i


def rank(Q, D, normalise=True):
    Q = np.ascontiguousarray(Q, np.float32)
    D = np.ascontiguousarray(D, np.float32)
    n1, n2 = Q.shape[0], D.shape[0]
    kg = n2 // n1 if n2 > n1 else 1
    hg = n1 // n2 if n1 > n2 else 1
    ranks = np.empty(n1, np.int64)
    ts = np.empty(n1, np.float32)
    lib().asr_oracle_rank(_f(Q), ctypes.c_long(n1), _f(D), ctypes.c_long(n2), ctypes.c_int(Q.shape[1]),
                          ctypes.c_long(kg), ctypes.c_long(hg), ctypes.c_int(int(normalise)),
                          ranks.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), _f(ts))
    return ranks, ts


def set_threads(n):
    lib().asr_oracle_set_threads(ctypes.c_int(int(n)))
