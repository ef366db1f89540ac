"""Two-qubit fidelity with virtual-Z phase calibration: the scalar tail of the hot path (src/fidelities.jl and
src/penalty_fcns.jl:27-42 of the reference), host side, vectorised over a batch of pulses.

Everything here works on m = diag(T' x) -- four complex numbers per pulse -- so it stays on the host exactly as in the
reference; only x_final crosses PCIe.  All functions accept m of shape (4,) or (..., 4).

    abs_sum_phase_calibrated(m, calibration)     src/fidelities.jl:11-40
    optimal_calibration(m, theta_tol)            :81-101   (golden-section search :105-137, run in lock-step over the batch)
    abs_sum_phase_calibrated_grad / rrule        :42-56
    infidelity(U_target, Uf, calibration)        :1-7
    setup_infidelity_zcalibrated(x_target)       src/penalty_fcns.jl:27-42
"""
from __future__ import annotations

import numpy as np

__all__ = ["abs_sum_phase_calibrated", "optimal_calibration", "basic_calibration", "grid_calibration",
           "abs_sum_phase_calibrated_grad", "abs_sum_phase_calibrated_rrule", "abs_trace_phase_calibrated", "infidelity",
           "setup_infidelity_zcalibrated"]

_GOLD = 0.5 * (3.0 - np.sqrt(5.0))


def _m4(m):
    m = np.asarray(m, dtype=np.complex128)
    if m.shape[-1] != 4:
        raise ValueError("m must hold four complex numbers per pulse")
    return m


def _golden_section_batch(f, lo, hi, tol):
    """src/fidelities.jl:105-137, every batch element advancing in lock-step; elements whose bracket is already
    below tol are frozen (the reference's while condition, per element)."""
    lo = np.array(lo, dtype=np.float64)
    hi = np.array(hi, dtype=np.float64)
    if np.any(lo > hi):
        raise ValueError("x_lower must be less than x_upper")
    xm = lo + _GOLD * (hi - lo)
    fm = f(xm)
    while True:
        act = (hi - lo) >= tol
        if not np.any(act):
            break
        right = (hi - xm) > (xm - lo)
        xn = np.where(right, xm + _GOLD * (hi - xm), xm - _GOLD * (xm - lo))
        fn = f(xn)
        better = fn < fm
        # four cases of the reference's branches
        new_lo = np.where(right & better, xm, np.where(~right & ~better, xn, lo))
        new_hi = np.where(right & ~better, xn, np.where(~right & better, xm, hi))
        new_xm = np.where(better, xn, xm)
        new_fm = np.where(better, fn, fm)
        lo = np.where(act, new_lo, lo)
        hi = np.where(act, new_hi, hi)
        xm = np.where(act, new_xm, xm)
        fm = np.where(act, new_fm, fm)
    return fm, xm


def optimal_calibration(m, theta_tol=1e-9):
    """src/fidelities.jl:81-101 -> (F, theta) with theta[..., 0:2] = (theta1, theta2)."""
    m = _m4(m)
    a1 = np.abs(m[..., 0]) ** 2 + np.abs(m[..., 1]) ** 2
    b1 = 2 * np.abs(m[..., 0]) * np.abs(m[..., 1])
    a2 = np.abs(m[..., 2]) ** 2 + np.abs(m[..., 3]) ** 2
    b2 = 2 * np.abs(m[..., 2]) * np.abs(m[..., 3])
    two_pi = 2 * np.pi
    p1 = np.mod(np.angle(m[..., 0]) - np.angle(m[..., 1]), two_pi)
    p2 = np.mod(np.angle(m[..., 2]) - np.angle(m[..., 3]), two_pi)
    near = np.abs(p2 - p1) <= np.pi
    pm = np.where(near, (p1 + p2) / 2, (two_pi + p1 + p2) / 2)
    D = np.where(near, np.abs(p2 - p1) / 2, np.pi - np.abs(p2 - p1) / 2)
    al = np.where(near, np.where(p1 < p2, 1.0, -1.0), np.where(p1 < p2, -1.0, 1.0))

    def negJ(dl):
        return -(np.sqrt(np.maximum(a1 + b1 * np.cos(dl + D), 0.0)) + np.sqrt(np.maximum(a2 + b2 * np.cos(dl - D), 0.0)))

    fmin, d_opt = _golden_section_batch(negJ, -D, D, theta_tol)
    t1 = pm + al * d_opt
    e1 = np.exp(1j * t1)
    t2 = np.angle(m[..., 0] + m[..., 1] * e1) - np.angle(m[..., 2] + m[..., 3] * e1)
    return -fmin, np.stack([t1, t2], axis=-1)


def basic_calibration(m):
    """src/fidelities.jl:65-69."""
    m = _m4(m)
    t0 = np.angle(m[..., 0])
    th1 = -(np.angle(m[..., 1]) - t0)
    th2 = -(np.angle(m[..., 2]) - t0)
    val = np.abs(m[..., 0] + m[..., 1] * np.exp(1j * th1) + m[..., 2] * np.exp(1j * th2) + m[..., 3] * np.exp(1j * (th1 + th2)))
    return val, np.stack([th1, th2], axis=-1)


def grid_calibration(m):
    """src/fidelities.jl:72-79 (100-point grid)."""
    m = _m4(m)
    th = np.linspace(0, 2 * np.pi, 100)
    e = np.exp(1j * th)
    vals = np.abs(m[..., 0, None] + m[..., 1, None] * e) + np.abs(m[..., 2, None] + m[..., 3, None] * e)
    k = np.argmax(vals, axis=-1)
    return np.take_along_axis(vals, k[..., None], axis=-1)[..., 0], th[k]


def abs_sum_phase_calibrated(m, calibration="optimal"):
    """src/fidelities.jl:11-40."""
    m = _m4(m)
    if calibration in ("lms_phase", "lms_phase2", "lms_phase3"):
        c12 = np.conj(m[..., 0]) * m[..., 1]
        c34 = np.conj(m[..., 2]) * m[..., 3]
        if calibration == "lms_phase":
            t = -np.angle(c12 + c34)
        elif calibration == "lms_phase2":
            x1, x2 = np.sqrt(np.abs(m[..., 0] * m[..., 1])), np.sqrt(np.abs(m[..., 2] * m[..., 3]))
            eps = np.finfo(float).eps
            small = (x1 < eps) | (x2 < eps)
            t = -np.angle(c12 / np.where(small, 1.0, x1) + c34 / np.where(small, 1.0, x2))
            val = np.abs(m[..., 0] + m[..., 1] * np.exp(1j * t)) + np.abs(m[..., 2] + m[..., 3] * np.exp(1j * t))
            return np.where(small, np.abs(m).sum(axis=-1), val)
        else:
            x1, x2 = np.abs(m[..., 0]) + np.abs(m[..., 1]), np.abs(m[..., 2]) + np.abs(m[..., 3])
            t = -np.angle(c12 / x1 + c34 / x2)
        return np.abs(m[..., 0] + m[..., 1] * np.exp(1j * t)) + np.abs(m[..., 2] + m[..., 3] * np.exp(1j * t))
    if calibration == "optimal":
        return optimal_calibration(m)[0]
    if calibration == "basic":
        return basic_calibration(m)[0]
    if calibration == "none":
        return np.abs(m.sum(axis=-1))
    if calibration == "grid":
        return grid_calibration(m)[0]
    raise ValueError(f"unknown calibration {calibration!r}")  # the reference silently returns nothing here


def _dF_dm(m, theta1):
    e1 = np.exp(1j * theta1)
    v1 = m[..., 0] + e1 * m[..., 1]
    v2 = m[..., 2] + e1 * m[..., 3]
    u1, u2 = v1 / np.abs(v1), v2 / np.abs(v2)
    return np.stack([u1, u1 * np.conj(e1), u2, u2 * np.conj(e1)], axis=-1), np.abs(v1) + np.abs(v2)


def abs_sum_phase_calibrated_grad(m, theta1_opt):
    """src/fidelities.jl:42-46: gradient of F^2 (= 2 F dF/dm) at the calibrated phase."""
    m = _m4(m)
    dF, F = _dF_dm(m, np.asarray(theta1_opt))
    return 2 * F[..., None] * dF


def abs_sum_phase_calibrated_rrule(m):
    """src/fidelities.jl:48-56 -> (F, dF_dm); theta held at its optimum (envelope theorem)."""
    m = _m4(m)
    F, th = optimal_calibration(m)
    dF, _ = _dF_dm(m, th[..., 0])
    return F, dF


def abs_trace_phase_calibrated(M, calibration="optimal"):
    """src/fidelities.jl:9."""
    M = np.asarray(M)
    return abs_sum_phase_calibrated(np.diagonal(M, axis1=-2, axis2=-1), calibration)


def infidelity(U_target, Uf, calibration="lms_phase"):
    """src/fidelities.jl:1-7 (4 x 4 only, like the reference)."""
    U_target = np.asarray(U_target)
    if U_target.shape != (4, 4):
        raise ValueError("Not supported yet")
    return 1 - abs_trace_phase_calibrated(U_target.conj().T @ np.asarray(Uf), calibration) / 4


def _zcal_host(T):
    def J(x):
        m = np.sum(T.conj() * np.asarray(x), axis=-2)  # diag(T' x)
        return 1 - abs_sum_phase_calibrated(m) ** 2 / 16

    def dJ_dx(x):
        m = np.sum(T.conj() * np.asarray(x), axis=-2)
        F, dF = abs_sum_phase_calibrated_rrule(m)
        return (-2 * np.asarray(F)[..., None, None] / 16) * T * dF[..., None, :]

    J.batched = dJ_dx.batched = True   # both accept (..., d, 4) stacks (grape_sensitivity hands a batch over in one call)
    return J, dJ_dx


def setup_infidelity_zcalibrated(x_target, device=True):
    """src/penalty_fcns.jl:27-42 -> (J, dJ_dx) on d x 4 states (or batches (..., d, 4)).

    device=True (default): the returned callables are built-in costs (QOC_COST_ZCAL): handed to propagate / grape_sensitivity /
    evaluate they make the library reduce m = diag(T'x), run the golden-section search and the rrule, and form J and
    lambda_N per pulse on the device -- x[end] never crosses the bus.  Called directly on numpy arrays they evaluate the same
    formulas on the host (vectorised over a batch).  device=False: plain host closures, the reference's own route
    (x[end] down, terminal costate up)."""
    T = np.asarray(x_target, dtype=np.complex128)
    if T.ndim != 2 or T.shape[1] != 4:
        raise ValueError("Only works for two-qubit gates, x_target must have four columns")
    Jh, dJh = _zcal_host(T)
    if not device:
        return Jh, dJh
    from .grape import _BuiltinCost
    from ._lib import COST_ZCAL

    class _ZcalJ(_BuiltinCost):
        kind = COST_ZCAL
        host_J, host_grad = staticmethod(Jh), staticmethod(dJh)

        def __call__(self, x):
            return Jh(x)

    class _ZcalGrad(_BuiltinCost):
        kind = COST_ZCAL
        host_J, host_grad = staticmethod(Jh), staticmethod(dJh)

        def __call__(self, x):
            return dJh(x)

    return _ZcalJ(T, 4), _ZcalGrad(T, 4)
