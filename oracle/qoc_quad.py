"""ctypes wrapper + build recipe for oracle/qoc_quad.c: exp(X) and its Frechet derivative in binary128, the ground truth
the oracle's own expm restatements are pinned against.  TEST INFRASTRUCTURE ONLY (see qoc_quad.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "qoc_quad.c")
LIB = os.path.join(HERE, "_build", "libqoc_quad.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    r = subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stderr)
    return LIB


def expm_quad(X, E=None):
    """-> exp(X) [, L(X, E)] computed in binary128, rounded to complex128."""
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.qoc_quad_expm.restype = C.c_int
    X = np.ascontiguousarray(X, dtype=np.complex128)
    d = X.shape[0]
    U = np.zeros((d, d), dtype=np.complex128)
    dp = C.POINTER(C.c_double)
    if E is None:
        rc = _lib.qoc_quad_expm(d, X.ctypes.data_as(dp), None, U.ctypes.data_as(dp), None)
        assert rc == 0
        return U
    E = np.ascontiguousarray(E, dtype=np.complex128)
    L = np.zeros((d, d), dtype=np.complex128)
    rc = _lib.qoc_quad_expm(d, X.ctypes.data_as(dp), E.ctypes.data_as(dp), U.ctypes.data_as(dp), L.ctypes.data_as(dp))
    assert rc == 0
    return U, L


def chain_quad(U, x0):
    """x_N = U_{N-1} ... U_0 x0 with the products accumulated in binary128 (U_k given in double) -> complex128."""
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.qoc_quad_expm.restype = C.c_int
    U = np.ascontiguousarray(U, dtype=np.complex128)
    x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.complex128).reshape(U.shape[1], -1))
    nt, d, m = U.shape[0], U.shape[1], x0.shape[1]
    out = np.zeros((d, m), dtype=np.complex128)
    dp = C.POINTER(C.c_double)
    _lib.qoc_quad_chain.restype = C.c_int
    rc = _lib.qoc_quad_chain(d, m, nt, U.ctypes.data_as(dp), x0.ctypes.data_as(dp), out.ctypes.data_as(dp))
    assert rc == 0
    return out
