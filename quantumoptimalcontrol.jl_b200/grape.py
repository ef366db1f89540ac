"""Host-side mirror of the reference's GRAPE interface over the C ABI (include/qoc_b200.h).

Same names, argument meaning and error behaviour as src/gradient_computations.jl and src/penalty_fcns.jl of
olof3/QuantumOptimalControl.jl, so the parity tests read like the reference's own scripts:

    cache = setup_grape_cache(A0, x0, (nc, Nt))                    # src/gradient_computations.jl:79-96
    x = propagate(A0dt, [A1dt, A2dt], u, x0, cache)                # :2-32   (x[-1] is the final state)
    dJdu = grape_sensitivity(A0dt, [A1dt, A2dt], dJfinal_dx, u, x0, cache, dUkdp_order=3, dL_dx=None)   # :35-77
    J, dJ_dx = setup_infidelity(x_target, n)                       # src/penalty_fcns.jl:15-24

The Julia reference cannot run in this image (no julia), so the host side that the north star places in Julia
is written here in Python; INTEGRATION.md carries the equivalent Julia `ccall` shim.  All numerics happen in
the CUDA library; this module only marshals numpy arrays (column-major complex128 == Julia ComplexF64).
There is no CPU fallback: without a B200 every compute call raises QOCError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (COST_ABS_TRACE, COST_INFIDELITY, COST_NONE, COST_ZCAL, ORDER_FRECHET)  # noqa: F401

__all__ = ["QOCError", "GrapeCache", "setup_grape_cache", "propagate", "grape_sensitivity", "evaluate",
           "setup_infidelity", "setup_infidelity_abs_trace", "setup_state_penalty", "setup_bilinear_matrices",
           "ORDER_FRECHET", "COST_INFIDELITY", "COST_ABS_TRACE", "COST_NONE", "COST_ZCAL"]


class QOCError(RuntimeError):
    def __init__(self, status, detail=""):
        lib = _lib.load()
        base = lib.qoc_status_string(int(status)).decode()
        super().__init__(f"{base}" + (f": {detail}" if detail and detail != base else ""))
        self.status = int(status)


def _c128(a):
    """-> Fortran-ordered complex128 copy (Julia Matrix{ComplexF64} memory layout)."""
    return np.asfortranarray(np.asarray(a, dtype=np.complex128))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


# ---- cost closures (src/penalty_fcns.jl) --------------------------------------------------------------------------
class _BuiltinCost:
    """A callable with the reference's closure semantics that additionally tells grape_sensitivity it can be
    evaluated on the device (fused into K2) instead of through a host round trip."""
    kind = COST_NONE

    def __init__(self, x_target, n):
        T = _c128(x_target)
        self.T = T.reshape(T.shape[0], -1, order="F")
        self.n = int(n) if n is not None else self.T.shape[1]


class _InfidelityJ(_BuiltinCost):
    kind = COST_INFIDELITY

    def __call__(self, x):  # src/penalty_fcns.jl:16-18
        return float(1 - abs(np.vdot(self.T, np.asarray(x).reshape(self.T.shape))) ** 2 / self.n ** 2)


class _InfidelityGrad(_BuiltinCost):
    kind = COST_INFIDELITY

    def __call__(self, x):  # src/penalty_fcns.jl:19-22
        om = np.vdot(self.T, np.asarray(x).reshape(self.T.shape))
        return (-2 * om / self.n ** 2) * self.T


class _AbsTraceJ(_BuiltinCost):
    kind = COST_ABS_TRACE

    def __call__(self, x):  # test/test_gradient_computation.jl:24
        return float(1 - abs(np.vdot(self.T, np.asarray(x).reshape(self.T.shape))))


class _AbsTraceGrad(_BuiltinCost):
    kind = COST_ABS_TRACE

    def __call__(self, x):  # Zygote gradient of the above (test/test_gradient_computation.jl:25)
        om = np.vdot(self.T, np.asarray(x).reshape(self.T.shape))
        return -(om / abs(om)) * self.T


def setup_infidelity(x_target, n=None):
    """src/penalty_fcns.jl:15-24 -> (J, dJ_dx).  The returned callables also work on plain numpy arrays."""
    return _InfidelityJ(x_target, n), _InfidelityGrad(x_target, n)


def setup_infidelity_abs_trace(x_target):
    """J = 1 - |tr(T'x)| with its Zygote-convention gradient (test/test_gradient_computation.jl:24-25)."""
    return _AbsTraceJ(x_target, None), _AbsTraceGrad(x_target, None)


class _StatePenalty:
    def __init__(self, rows, cols, mu):
        self.rows = np.asarray(rows, dtype=np.int32)
        self.cols = np.asarray(cols, dtype=np.int32)
        self.mu = float(mu)


class _StatePenaltyL(_StatePenalty):
    def __call__(self, x):  # src/penalty_fcns.jl:2-4
        return float(self.mu * np.sum(np.abs(np.asarray(x)[np.ix_(self.rows, self.cols)]) ** 2))


class _StatePenaltyGrad(_StatePenalty):
    def __call__(self, x):  # src/penalty_fcns.jl:5-9
        x = np.asarray(x)
        g = np.zeros_like(x)
        g[np.ix_(self.rows, self.cols)] = 2 * self.mu * x[np.ix_(self.rows, self.cols)]
        return g


def setup_state_penalty(inds_penalty, inds_css, mu):
    """src/penalty_fcns.jl:1-11 (0-based indices) -> (L, dL_dx)."""
    return _StatePenaltyL(inds_penalty, inds_css, mu), _StatePenaltyGrad(inds_penalty, inds_css, mu)


def setup_bilinear_matrices(H0, Tc, dt=1.0):
    """src/utils.jl:86-91 (pure input preparation, stays on the host as in the reference)."""
    H0 = np.asarray(H0, dtype=np.complex128)
    Tc = np.asarray(Tc, dtype=np.complex128)
    return -1j * H0 * dt, -1j * (Tc + Tc.conj().T) * dt, -1j * (1j * (Tc - Tc.conj().T)) * dt


# ---- cache / handle ---------------------------------------------------------------------------------------------------
class GrapeCache:
    """Stands in for the NamedTuple of setup_grape_cache (x, lambda, dJdu, Uk_vec, exp_cache, u): the arrays live in
    HBM inside the C handle and are fetched on attribute access."""

    def __init__(self, A0, x0, u_size, batch=1, device=0, dUkdp_order=3, store_costates=True):
        A0 = np.asarray(A0)
        x0 = np.asarray(x0)
        x0m = x0.reshape(x0.shape[0], -1)
        d = A0.shape[0]
        # src/gradient_computations.jl:84-87
        # (the real 2d-row "c2r" representation of :84 belongs to the ODE path, which is out of scope here)
        if x0m.shape[0] != d:
            raise QOCError(_lib.ERR_DIMENSION)
        self.d, self.m = d, x0m.shape[1]
        self.nc, self.nt = int(u_size[0]), int(u_size[1])
        self.batch, self.device = int(batch), int(device)
        self.order = int(dUkdp_order)
        self.store_costates = bool(store_costates)
        self._h = None
        self._key = None
        self._cost_key = None
        self._penalty = None
        self.J = None
        self.x_final = None   # set by propagate(); closure costs in grape_sensitivity read it

    # -- handle management --
    def _ensure(self, A0, A, x0, penalty=None):
        A0c = _c128(A0)
        Ac = np.stack([_c128(a) for a in A], axis=0) if len(A) else np.zeros((0, self.d, self.d), np.complex128)
        x0c = _c128(np.asarray(x0).reshape(np.asarray(x0).shape[0], -1))
        if x0c.shape[0] != self.d:
            raise QOCError(_lib.ERR_DIMENSION)
        pen_key = None if penalty is None else (penalty.rows.tobytes(), penalty.cols.tobytes(), penalty.mu)
        key = (A0c.tobytes(), Ac.tobytes(), x0c.tobytes(), pen_key)
        if self._h is not None and key == self._key:
            return
        self.close()
        lib = _lib.load()
        pr = _lib.Problem()
        pr.d, pr.m, pr.nc, pr.nt, pr.batch = self.d, x0c.shape[1], len(A), self.nt, self.batch
        pr.order, pr.cost, pr.n, pr.device = self.order, COST_NONE, 0, self.device
        pr.store_costates = 1 if self.store_costates else 0
        keep = []
        if penalty is not None and len(penalty.rows) and len(penalty.cols):
            rows = np.ascontiguousarray(penalty.rows, dtype=np.int32)
            cols = np.ascontiguousarray(penalty.cols, dtype=np.int32)
            keep += [rows, cols]
            pr.n_pen_rows, pr.n_pen_cols = len(rows), len(cols)
            pr.pen_rows = rows.ctypes.data_as(C.POINTER(C.c_int32))
            pr.pen_cols = cols.ctypes.data_as(C.POINTER(C.c_int32))
            pr.mu = penalty.mu
        h = C.c_void_p()
        # per-matrix Fortran order: (nc, d, d) stack of column-major matrices
        Aflat = np.ascontiguousarray(np.stack([a.T for a in Ac], axis=0)) if len(A) else Ac
        rc = lib.qoc_create(C.byref(pr), _dptr(A0c), _dptr(Aflat), _dptr(x0c), None, C.byref(h))
        if rc != _lib.OK:
            raise QOCError(rc, lib.qoc_last_error(None).decode())
        self._h, self._key, self._cost_key, self._penalty = h, key, None, penalty
        self._ctor_args = (A0, list(A), x0, penalty)
        self.m = x0c.shape[1]

    def _set_cost(self, cost_obj):
        lib = _lib.load()
        if isinstance(cost_obj, _BuiltinCost):
            key = (cost_obj.kind, cost_obj.T.tobytes(), cost_obj.n)
            if key != self._cost_key:
                self._check(lib.qoc_set_cost(self._h, cost_obj.kind, _dptr(cost_obj.T), cost_obj.n))
                self._cost_key = key
        elif self._cost_key is not None:
            self._check(lib.qoc_set_cost(self._h, COST_NONE, None, 0))
            self._cost_key = None

    def _ensure_x0_cost(self, cost_obj, x0):
        """Time sharding: (re)create the handle with the pulse's true x0 and install the built-in cost."""
        A0c, Ac, _, pen = self._ctor_args
        self._ensure(A0c, Ac, x0, pen)
        self._set_cost(cost_obj)

    def _check(self, rc):
        if rc != _lib.OK:
            raise QOCError(rc, _lib.load().qoc_last_error(self._h).decode())

    def close(self):
        if self._h is not None:
            _lib.load().qoc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the NamedTuple fields of the reference cache --
    def _states_shape(self):
        return (self.d, self.m, self.nt + 1, self.batch)

    @property
    def x(self):
        """cache.x: array (Nt+1, d, m) (or (batch, Nt+1, d, m))."""
        out = np.zeros(self._states_shape(), dtype=np.complex128, order="F")
        self._check(_lib.load().qoc_get_states(self._h, _dptr(out)))
        return self._squeeze(np.transpose(out, (3, 2, 0, 1)))

    @property
    def lam(self):
        out = np.zeros(self._states_shape(), dtype=np.complex128, order="F")
        self._check(_lib.load().qoc_get_costates(self._h, _dptr(out)))
        return self._squeeze(np.transpose(out, (3, 2, 0, 1)))

    @property
    def Uk_vec(self):
        out = np.zeros((self.d, self.d, self.nt, self.batch), dtype=np.complex128, order="F")
        self._check(_lib.load().qoc_get_propagators(self._h, _dptr(out)))
        return self._squeeze(np.transpose(out, (3, 2, 0, 1)))

    @property
    def dUkdu(self):
        out = np.zeros((self.d, self.d, self.nc, self.nt, self.batch), dtype=np.complex128, order="F")
        self._check(_lib.load().qoc_get_jacobians(self._h, _dptr(out)))
        return self._squeeze(np.transpose(out, (4, 3, 2, 0, 1)))

    def _squeeze(self, a):
        return a[0] if self.batch == 1 else a

    # -- diagnostics --
    def launch_count(self):
        return _lib.load().qoc_last_launch_count(self._h)

    def alg_flops(self):
        return _lib.load().qoc_last_alg_flops(self._h)

    def exec_flops(self):
        return _lib.load().qoc_last_exec_flops(self._h)

    def set_profiling(self, on=True):
        self._check(_lib.load().qoc_set_profiling(self._h, 1 if on else 0))

    def stage_ms(self):
        lib = _lib.load()
        return [lib.qoc_stage_ms(self._h, i) for i in range(3)]

    @property
    def handle(self):
        return self._h


def setup_grape_cache(A0, x0, u_size, batch=1, device=0, dUkdp_order=3, store_costates=True):
    """src/gradient_computations.jl:79-96.  Raises the reference's dimension error (:84-87)."""
    return GrapeCache(A0, x0, u_size, batch=batch, device=device, dUkdp_order=dUkdp_order,
                      store_costates=store_costates)


def _u_arr(u, cache):
    u = np.asarray(u, dtype=np.float64)
    if cache.batch == 1:
        if u.shape != (cache.nc, cache.nt):
            raise QOCError(_lib.ERR_DIMENSION, f"u has shape {u.shape}, cache expects {(cache.nc, cache.nt)}")
        return np.ascontiguousarray(u.T)  # element (j,k) at j + nc*k
    if u.shape != (cache.batch, cache.nc, cache.nt):
        raise QOCError(_lib.ERR_DIMENSION, f"u has shape {u.shape}, expected (batch, nc, Nt)")
    return np.ascontiguousarray(np.transpose(u, (0, 2, 1)))  # (b, k, j): j fastest


def propagate(A0, A, u, x0, cache=None, Jfinal=None, penalty=None, eager_jacobians=False):
    """src/gradient_computations.jl:2-32.  Returns the cache (cache.x are the states; the reference returns x).
    Like the reference's propagate this computes the exponentials only; grape_sensitivity adds the Jacobians on the cached u
    (eager_jacobians=True produces them here, in the same kernel pass, when a gradient always follows).
    Jfinal: optional built-in cost from setup_infidelity*: J = Jfinal(x[end]) (+ sum(L, x)) is then formed on the
    device and left in cache.J (examples/ipopt_callbacks_exp.jl:18).  penalty = (L, dL_dx) from setup_state_penalty."""
    u = np.asarray(u, dtype=np.float64)
    if cache is None:
        nb = 1 if u.ndim == 2 else u.shape[0]
        cache = setup_grape_cache(A0, np.asarray(x0, dtype=np.complex128), u.shape[-2:], batch=nb)
    pen = penalty[0] if isinstance(penalty, tuple) else penalty
    cache._ensure(A0, A, np.asarray(x0, dtype=np.complex128), pen)
    if Jfinal is not None:
        cache._set_cost(Jfinal)
    lib = _lib.load()
    cache._check(lib.qoc_set_eager_jacobians(cache._h, 1 if eager_jacobians else 0))
    uu = _u_arr(u, cache)
    J = np.zeros(cache.batch)
    xf = np.zeros((cache.d, cache.m, cache.batch), dtype=np.complex128, order="F")
    cache._check(lib.qoc_propagate(cache._h, _dptr(uu), _dptr(J), _dptr(xf)))
    cache.J = (J[0] if cache.batch == 1 else J) if cache._cost_key is not None or pen is not None else None
    cache.x_final = xf[:, :, 0] if cache.batch == 1 else np.transpose(xf, (2, 0, 1))
    return cache


def grape_sensitivity(A0, A, dJfinal_dx, u, x0, cache, dUkdp_order=3, dL_dx=None):
    """src/gradient_computations.jl:35-77.  dJfinal_dx: a built-in cost gradient (evaluated on the device) or any
    callable x_final -> d x m array (evaluated here on the host, like the reference's closure)."""
    lib = _lib.load()
    if cache._h is None:
        raise QOCError(_lib.ERR_STALE_CACHE, "propagate has not been called on this cache")
    if dL_dx is not None and cache._penalty is None:
        raise QOCError(_lib.ERR_INVALID, "dL_dx given but the cache was propagated without that penalty")
    uu = _u_arr(u, cache)
    cache._check(lib.qoc_set_order(cache._h, int(dUkdp_order)))
    cache.order = int(dUkdp_order)
    g = np.zeros(uu.shape)
    if isinstance(dJfinal_dx, _BuiltinCost):
        cache._set_cost(dJfinal_dx)
        cache._check(lib.qoc_gradient(cache._h, _dptr(uu), None, _dptr(g)))
    else:
        xf = cache.x_final
        if xf is None:  # evaluate() alone does not bring x[end] to the host
            raise QOCError(_lib.ERR_STALE_CACHE, "a host-closure dJfinal_dx needs propagate() on this cache first")
        if cache.batch == 1:
            lam = _c128(dJfinal_dx(xf)).reshape(cache.d, cache.m, 1, order="F")
        elif getattr(dJfinal_dx, "batched", False):   # a closure that takes the whole (batch, d, m) stack at once
            lam = np.transpose(np.asarray(dJfinal_dx(xf), dtype=np.complex128), (1, 2, 0))
        else:
            lam = np.asfortranarray(np.stack([_c128(dJfinal_dx(xf[b])) for b in range(cache.batch)], axis=2))
        lam = np.asfortranarray(lam)
        cache._check(lib.qoc_gradient(cache._h, _dptr(uu), _dptr(lam), _dptr(g)))
    return g.T.copy() if cache.batch == 1 else np.transpose(g, (0, 2, 1)).copy()


def evaluate(cache, A0, A, u, x0, cost, dUkdp_order=None, penalty=None):
    """Fused f + f_grad (examples/ipopt_callbacks_exp.jl:11-31) -> (J, dJdu) through qoc_eval."""
    lib = _lib.load()
    pen = penalty[0] if isinstance(penalty, tuple) else penalty
    cache._ensure(A0, A, np.asarray(x0, dtype=np.complex128), pen)
    if dUkdp_order is not None:
        cache._check(lib.qoc_set_order(cache._h, int(dUkdp_order)))
        cache.order = int(dUkdp_order)
    cache._set_cost(cost)
    uu = _u_arr(u, cache)
    J = np.zeros(cache.batch)
    g = np.zeros(uu.shape)
    cache._check(lib.qoc_eval(cache._h, _dptr(uu), _dptr(J), _dptr(g)))
    cache.x_final = None   # belongs to an earlier propagate(), not to this u
    if cache.batch == 1:
        return float(J[0]), g.T.copy()
    return J, np.transpose(g, (0, 2, 1)).copy()
