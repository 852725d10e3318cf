"""Oracle (TEST INFRASTRUCTURE ONLY): ctypes access to ``oracle/c/liboracle_dcr.so`` (plain-C paper-flavour BFC).

Used by tests as a fast second checker and by ``bench.py`` as the timed CPU baseline; built by
``__graft_entry__.build()`` (``make -C oracle/c``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "liboracle_dcr.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-s", "-C", _DIR], check=True)
        lib = C.CDLL(_SO)
        lib.oracle_bfc_paper.restype = C.c_int
        lib.oracle_bfc_paper.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def bfc_paper_c(rowptr: np.ndarray, col: np.ndarray, esrc: np.ndarray, edst: np.ndarray, threads: int = 1) -> dict:
    """Fields of ``bfc_naive.bfc_edge`` for the listed edges (bfc_naive.py:7-40), ``threads`` pthreads."""
    lib = load()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    esrc = np.ascontiguousarray(esrc, dtype=np.int32)
    edst = np.ascontiguousarray(edst, dtype=np.int32)
    m = esrc.size
    tri, s1, s2, gm = (np.zeros(m, dtype=np.int32) for _ in range(4))
    val = np.zeros(m, dtype=np.float64)
    rc = lib.oracle_bfc_paper(rowptr.ctypes.data, col.ctypes.data, rowptr.size - 1, esrc.ctypes.data,
                              edst.ctypes.data, m, tri.ctypes.data, s1.ctypes.data, s2.ctypes.data, gm.ctypes.data,
                              val.ctypes.data, int(threads))
    if rc != 0:
        raise MemoryError("oracle_bfc_paper failed")
    return {"tri": tri, "sq_i": s1, "sq_j": s2, "gamma": gm, "bfc": val}
