"""ctypes wrapper + build recipe for oracle/qoc_ref.c (the plain-C restatement of the reference; CPU baseline).

TEST INFRASTRUCTURE ONLY (see qoc_ref.c header).  Built with gcc into oracle/_build/ (git-ignored, travels to the
GPU box with the repo snapshot).  `oracle/_ref/` (the real reference compiled) does not exist for this project:
the reference is Julia and no julia toolchain is in the image."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "qoc_ref.c")
LIB = os.path.join(HERE, "_build", "libqoc_ref.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # -march=x86-64-v3 (AVX2+FMA) rather than native: the .so is built in one container and runs on the GPU box
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.qoc_ref_eval.restype = C.c_int
        _lib.qoc_ref_max_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def ref_eval(cfg, order=3, nthreads=0, want_grad=True, u=None, penalty=None, want_cache=False):
    """One fidelity+gradient evaluation of one pulse by the C restatement.
    -> dict(J, dJdu (nc x Nt), flops, [Uk, x, lam, dU])."""
    lib = load()
    A0 = np.asfortranarray(cfg["A0"], dtype=np.complex128)
    A = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.complex128).T for a in cfg["A"]]))  # col-major each
    u = np.asarray(cfg["u"] if u is None else u, dtype=np.float64)
    nc, nt = u.shape
    uu = np.ascontiguousarray(u.T)
    x0 = np.asfortranarray(cfg["x0"], dtype=np.complex128)
    T = np.asfortranarray(cfg["T"], dtype=np.complex128)
    d, m = x0.shape
    J = C.c_double(0.0)
    fl = C.c_double(0.0)
    g = np.zeros((nt, nc))
    rows = cols = None
    mu = 0.0
    if penalty is not None:
        rows = np.ascontiguousarray(penalty[0], dtype=np.int32)
        cols = np.ascontiguousarray(penalty[1], dtype=np.int32)
        mu = float(penalty[2])
    Uk = xs = lams = dU = None
    if want_cache:
        Uk = np.zeros((nt, d, d), dtype=np.complex128)
        xs = np.zeros((nt + 1, m, d), dtype=np.complex128)
        lams = np.zeros((nt + 1, m, d), dtype=np.complex128)
        dU = np.zeros((nt, nc, d, d), dtype=np.complex128)
    rc = lib.qoc_ref_eval(C.c_int(d), C.c_int(m), C.c_int(nc), C.c_int(nt), _p(A0), _p(A), _p(uu), _p(x0), _p(T),
                          C.c_int(cfg["cost"]), C.c_int(cfg["n"]), C.c_int(order), C.c_int(nthreads),
                          C.c_int(1 if want_grad else 0), C.c_int(0 if rows is None else len(rows)),
                          C.c_int(0 if cols is None else len(cols)), _p(rows, C.c_int32), _p(cols, C.c_int32),
                          C.c_double(mu), C.byref(J), _p(g), _p(Uk), _p(xs), _p(lams), _p(dU), C.byref(fl))
    if rc != 0:
        raise RuntimeError("qoc_ref_eval failed (singular Pade denominator)")
    out = dict(J=J.value, dJdu=g.T.copy(), flops=fl.value)
    if want_cache:  # stored column-major per matrix -> transpose the trailing two axes
        out.update(Uk=np.transpose(Uk, (0, 2, 1)), x=np.transpose(xs, (0, 2, 1)), lam=np.transpose(lams, (0, 2, 1)),
                   dU=np.transpose(dU, (0, 1, 3, 2)))
    return out


def max_threads():
    return load().qoc_ref_max_threads()
