"""The immediate caller of the GRAPE path: spline-parameterised pulses and the optimiser callbacks
(examples/ipopt_callbacks_exp.jl:1-54 of olof3/QuantumOptimalControl.jl), mirrored over the C ABI.

    f, g, f_grad, g_jac, nu, ng, nx, nc, cache = setup_ipopt_callbacks(A0dt, A1dt, A2dt, x0, u_prototype,
                                                                       (Jfinal, dJfinal_dx), (L, dL_dx), B)

f / f_grad keep the reference's two-call protocol (f_grad re-runs f when it is asked for a new point, :22-25).
`f_and_grad` is the fused fast path: u = (B c)' and dJ/dc = B' (dJ/du)' are formed on the device
(qoc_set_basis / qoc_eval_coeffs), so only the ns x nu coefficients cross the bus per evaluation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, grape

__all__ = ["setup_ipopt_callbacks", "constraints", "constraints_jacobian", "minimize_lbfgs"]


def constraints(c, nsplines, nu):
    """g_oop of examples/ipopt_callbacks_exp.jl:33-37: [norm(c); norm(diff(c, dims=1))], c reshaped nsplines x nu."""
    cm = np.asarray(c, dtype=np.float64).reshape(nu, nsplines).T          # column-major reshape(c, nsplines, nu)
    return np.array([np.linalg.norm(cm), np.linalg.norm(np.diff(cm, axis=0))])


def constraints_jacobian(c, nsplines, nu):
    """Jacobian (ng x nc) of `constraints` -- what Zygote.jacobian(g_oop, c) returns at :48."""
    cm = np.asarray(c, dtype=np.float64).reshape(nu, nsplines).T
    n1 = np.linalg.norm(cm)
    dc = np.diff(cm, axis=0)
    n2 = np.linalg.norm(dc)
    j1 = cm / n1 if n1 > 0 else np.zeros_like(cm)
    j2 = np.zeros_like(cm)
    if n2 > 0:
        j2[1:] += dc / n2
        j2[:-1] -= dc / n2
    return np.stack([j1.T.reshape(-1), j2.T.reshape(-1)], axis=0)         # back to the flat column-major order of c


def setup_ipopt_callbacks(A0dt, A1dt, A2dt, x0, u_prototype, cost, penalty, B, dUkdp_order=3, device=0):
    """examples/ipopt_callbacks_exp.jl:1-54.  cost = (Jfinal, dJfinal_dx) from setup_infidelity*, penalty = (L, dL_dx)
    from setup_state_penalty or None (the shipped example disables it, examples/zz_coupling_ipopt_exp.jl:46).
    Returns (f, g, f_grad, g_jac, nu, ng, nx, nc, cache) like the reference, with f.f_and_grad as the fused path."""
    Jfinal, dJfinal_dx = cost
    B = np.asfortranarray(np.asarray(B, dtype=np.float64))
    nu = np.asarray(u_prototype).shape[0]
    segment_count = B.shape[0]
    ng = 2
    x0 = np.asarray(x0, dtype=np.complex128)
    nx = x0.size
    nsplines = B.shape[1]
    nc = nu * nsplines
    A = [A1dt, A2dt][:nu] if nu <= 2 else None
    if A is None:
        raise ValueError("setup_ipopt_callbacks mirrors the reference's two-control interface")
    c_prev = np.full(nc, np.nan)
    cache = grape.setup_grape_cache(A0dt, x0, (nu, segment_count), device=device, dUkdp_order=dUkdp_order)
    pen = None if penalty is None else penalty
    state = {"basis_set": False}

    def to_u(c):
        cm = np.asarray(c, dtype=np.float64).reshape(nu, nsplines).T      # reshape(c, nsplines, nu)
        return (B @ cm).T                                                 # u = transpose(B*c)      (:13-14)

    def f(c):
        c_prev[:] = c
        u = to_u(c)
        grape.propagate(A0dt, A, u, x0, cache, Jfinal=Jfinal, penalty=pen)
        f.u = u
        return float(cache.J)                                             # Jfinal(x[end]) + sum(L, x)   (:18)

    def f_grad(c, f_grad_out=None):
        if not np.array_equal(c_prev, np.asarray(c)):
            f(c)                                                          # (:22-25)
        dJdu = grape.grape_sensitivity(A0dt, A, dJfinal_dx, f.u, x0, cache, dUkdp_order=dUkdp_order,
                                       dL_dx=None if pen is None else pen[1])
        dJdc = B.T @ dJdu.T                                               # (:28)
        out = dJdc.T.reshape(-1)                                          # dJdc[:] (column-major flattening)
        if f_grad_out is not None:
            f_grad_out[:] = out
        return out

    def f_and_grad(c):
        """Fused device path: J and dJ/dc with both skinny products on the device."""
        lib = _lib.load()
        cache._ensure(A0dt, A, x0, None if pen is None else pen[0])
        cache._check(lib.qoc_set_order(cache._h, int(dUkdp_order)))
        cache._set_cost(dJfinal_dx)
        if not state["basis_set"] or state.get("h") != cache._h.value:
            cache._check(lib.qoc_set_basis(cache._h, B.ctypes.data_as(C.POINTER(C.c_double)), nsplines))
            state["basis_set"], state["h"] = True, cache._h.value
        cc = np.ascontiguousarray(np.asarray(c, dtype=np.float64))        # flat c is already ns x nu column-major
        J = np.zeros(1)
        dJdc = np.zeros(nc)
        cache._check(lib.qoc_eval_coeffs(cache._h, cc.ctypes.data_as(C.POINTER(C.c_double)),
                                         J.ctypes.data_as(C.POINTER(C.c_double)), dJdc.ctypes.data_as(C.POINTER(C.c_double))))
        c_prev[:] = np.nan
        return float(J[0]), dJdc

    def g(c, g_out=None):
        v = constraints(c, nsplines, nu)
        if g_out is not None:
            g_out[:] = v
        return v

    def g_jac(c, mode, rows, cols, g_jac_out):
        if mode == "Structure":                                           # (:43-46), 1-based like the reference
            cols[:] = np.kron(np.ones(ng), np.arange(1, nc + 1))
            rows[:] = np.kron(np.arange(1, ng + 1), np.ones(nc))
        else:
            g_jac_out[:] = constraints_jacobian(c, nsplines, nu).reshape(-1)   # transpose(g_jac_tmp)[:]

    f.f_and_grad = f_and_grad
    f.to_u = to_u
    return f, g, f_grad, g_jac, nu, ng, nx, nc, cache


def minimize_lbfgs(f_and_grad, c0, bounds=None, maxiter=50):
    """A complete optimiser loop around the path without Ipopt (absent from this image): scipy's L-BFGS-B with the
    reference's box bounds on the coefficients (examples/zz_coupling_ipopt_exp.jl:52-56).  The norm constraints of the
    reference (g <= [2, 1]) are not enforced here; `constraints` reports them."""
    from scipy.optimize import minimize
    res = minimize(lambda c: f_and_grad(c), np.asarray(c0, dtype=np.float64), jac=True, method="L-BFGS-B", bounds=bounds,
                   options={"maxiter": maxiter})
    return res
