"""ctypes loader of the oracle's C restatement of the CRR recursion (oracle/crr_tree.c; TEST
INFRASTRUCTURE, see oracle/__init__.py).  `make -C oracle` builds oracle/_c/liboracle_crr.so."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_c", "liboracle_crr.so")
_lib = None


def load(build=True):
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "crr_tree.c")
        if build and (not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src)):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        sig = [C.c_void_p] * 7 + [C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        for name in ("oracle_crr_loop_ref", "oracle_crr_loop_fast"):
            getattr(_lib, name).argtypes = sig
            getattr(_lib, name).restype = C.c_int
    return _lib


def _run(name, S0, K, T, r, q, sigma, is_call, N, american):
    lib = load()
    cols = [np.ascontiguousarray(np.asarray(c, dtype=np.float64)) for c in (S0, K, T, r, q, sigma)]
    call = np.ascontiguousarray(np.asarray(is_call, dtype=np.uint8))
    n = cols[0].size
    price = np.empty(n)
    status = np.zeros(n, dtype=np.uint8)
    rc = getattr(lib, name)(*[c.ctypes.data_as(C.c_void_p) for c in cols], call.ctypes.data_as(C.c_void_p), n,
                            int(N), 1 if american else 0, price.ctypes.data_as(C.c_void_p),
                            status.ctypes.data_as(C.c_void_p))
    if rc == 1:
        raise ValueError("N must be positive.")
    if rc:
        raise MemoryError("oracle_crr: allocation failed")
    return price, status


def crr_loop_ref(S0, K, T, r, q, sigma, is_call, N, american):
    """binomial.py:7-38 in the reference's operation order (two pow() per node)."""
    return _run("oracle_crr_loop_ref", S0, K, T, r, q, sigma, is_call, N, american)


def crr_loop_fast(S0, K, T, r, q, sigma, is_call, N, american):
    """Same lattice from a table of u^m, no pow() in the level loop (not bit-identical: O(1e-13) relative)."""
    return _run("oracle_crr_loop_fast", S0, K, T, r, q, sigma, is_call, N, american)
