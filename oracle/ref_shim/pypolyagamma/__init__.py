"""ORACLE shim: stand-in for pypolyagamma's three entry points used at
pyglm/regression.py:474-477 and :501-508, backed by oracle/pg_devroye.c (per-thread RNG mode,
the structure of pgdrawvpar)."""
import ctypes
import os

import numpy as np

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "libpg_oracle.so")
        _lib = ctypes.CDLL(os.path.abspath(path))
        _lib.pg1_drawv.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.pg_omp_max_threads.restype = ctypes.c_int
    return _lib


def get_omp_num_threads():
    return int(_load().pg_omp_max_threads())


class PyPolyaGamma(object):
    def __init__(self, seed=0):
        self.seed = int(seed)
        self.calls = 0


def pgdrawvpar(ppgs, ns, zs, out):
    assert ns.shape == zs.shape == out.shape and out.flags.c_contiguous
    assert np.all(ns == 1.0), "oracle restates PG(1, z) only (all the Bernoulli path uses)"
    zs = np.ascontiguousarray(zs, dtype=np.float64)
    head = ppgs[0]
    head.calls += 1
    _load().pg1_drawv(zs.ctypes.data, zs.size, head.seed, head.calls, 1, len(ppgs), out.ctypes.data)
