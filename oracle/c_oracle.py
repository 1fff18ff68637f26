"""ctypes driver of oracle/pa_oracle.c (CPU oracle / CPU baseline; test and
bench infrastructure only -- see the header of pa_oracle.c)."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpa_oracle.so")


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.pao_transpose.restype = C.c_int
        _lib.pao_sizes.restype = C.c_int
        _lib.pao_max_threads.restype = C.c_int
    return _lib


def _i64(v):
    return (C.c_int64 * max(1, len(v)))(*v)


def _int(v):
    return (C.c_int * max(1, len(v)))(*v)


def sizes(pdims, size_global, decomp_in, decomp_out, extra, rank):
    out = (C.c_int64 * 4)()
    lib().pao_sizes(len(pdims), _i64(pdims), len(size_global), _i64(size_global), _int(decomp_in),
                    _int(decomp_out), len(extra), _i64(extra), rank, out)
    return tuple(out)


class CTranspose:
    """All ranks of one transpose!: allocates staging once, runs many times."""

    def __init__(self, pdims, size_global, decomp_in, perm_in, decomp_out, perm_out, extra, dtype):
        self.args = (pdims, size_global, decomp_in, perm_in, decomp_out, perm_out, extra)
        self.dtype = np.dtype(dtype)
        self.nranks = math.prod(pdims)
        self.sz = [sizes(pdims, size_global, decomp_in, decomp_out, extra, r)
                   for r in range(self.nranks)]
        self.send = [np.zeros(max(1, s[2]), dtype=dtype) for s in self.sz]
        self.recv = [np.zeros(max(1, s[3]), dtype=dtype) for s in self.sz]

    def run(self, srcs, dsts, nthreads=None):
        pdims, size_global, decomp_in, perm_in, decomp_out, perm_out, extra = self.args
        P = C.c_void_p * self.nranks
        ph = (C.c_double * 3)()
        rc = lib().pao_transpose(
            len(pdims), _i64(pdims), len(size_global), _i64(size_global), _int(decomp_in),
            _int(perm_in) if perm_in else None, _int(decomp_out),
            _int(perm_out) if perm_out else None, len(extra), _i64(extra), self.dtype.itemsize,
            P(*[a.ctypes.data for a in srcs]), P(*[a.ctypes.data for a in dsts]),
            P(*[a.ctypes.data for a in self.send]), P(*[a.ctypes.data for a in self.recv]),
            nthreads or lib().pao_max_threads(), ph)
        if rc != 0:
            raise ValueError("ArgumentError: pencil decompositions must differ in at most one dimension.")
        return tuple(ph)
