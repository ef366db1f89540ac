"""CPU oracle for the piecewise-constant GRAPE hot path of olof3/QuantumOptimalControl.jl.

TEST INFRASTRUCTURE ONLY.  This file is a numpy/scipy restatement of the reference's algorithm; it is
imported only by tests/, tools/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (quantumoptimalcontrol.jl_b200/) never imports it and has no CPU fallback.

PARITY PINNING.  The reference (Julia) cannot run in this image (no julia binary, no depot), and its expm
lives in the un-vendored, un-pinned dependency ExponentialUtilities.jl (Project.toml:9; single call site
src/gradient_computations.jl:24) -> the U_k = exp(X_k) boundary itself is "parity unpinned" (no reference
test holds expm values).  What IS pinned, and checked in tests/test_oracle.py:
  * examples/cavity_qubit.jl:80-81      overlap "about 0.999979"            (propagate + model + CSV pulse)
  * examples/two_qubit_tunable_bus.jl:66-67  population "something like 0.937218" (propagate + bus model)
  * test/test_fidelities.jl:19-84       every abs_sum_phase_calibrated known answer
  * test/test_expm_jacobian.jl:18-35    truncated-Taylor Jacobian error thresholds (structure)
  * test/test_penalty_fcns.jl:13-40     cost gradients == Zygote-convention gradient (finite differences here)
  * test/test_gradient_computation.jl:97-98  analytic gradient vs finite differences of propagate()
  * expm_higham2005 (own restatement of the published Higham-2005 algorithm the dependency implements)
    agrees with scipy.linalg.expm (Al-Mohy-Higham 2009) to 1e-13.

Every function cites the reference file:line it follows.  Arrays are numpy complex128/float64; matrices are
ordinary 2-D arrays (row/col meaning identical to Julia's; memory order is irrelevant here, the C ABI is
column-major and the host mirror converts).
"""
from __future__ import annotations

import math
import os
from itertools import product as _iproduct

import numpy as np
import scipy.linalg as sla

import sys as _sys

_sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "quantumoptimalcontrol.jl_b200"))
# input generators (models, pulses, targets) are shared with the product so both sides see identical bits;
# they contain no propagation / expm arithmetic
from configs import *  # noqa: F401,F403,E402
from configs import COST_ABS_TRACE, COST_INFIDELITY, COST_NONE  # noqa: F401,E402

# --------------------------------------------------------------------------------------------------------
# ExponentialUtilities.exponential!(A, ExpMethodHigham2005())  -- third party, restated from the published
# algorithm (Higham, SIMAX 26(4) 2005; identical in structure to Julia stdlib LinearAlgebra.exp!).
# --------------------------------------------------------------------------------------------------------

PADE_B = {
    3: (120., 60., 12., 1.),
    5: (30240., 15120., 3360., 420., 30., 1.),
    7: (17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.),
    9: (17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.),
    13: (64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
         10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.),
}
# degree switch points as used by Julia's exp!/ExponentialUtilities (rounded Higham-2005 theta_m)
THETA_EXPM = {3: 0.015, 5: 0.25, 7: 0.95, 9: 2.1, 13: 5.4}
# Al-Mohy & Higham 2009 table 6.1 (l_m) used for the exact Frechet mode (same table as scipy _expm_frechet.py)
ELL_FRECHET = {3: 1.08e-2, 5: 2.00e-1, 7: 7.83e-1, 9: 1.78, 13: 4.74}


def select_degree(norm1: float, table=THETA_EXPM):
    """-> (q, s): Pade degree and number of squarings for a 1-norm."""
    for q in (3, 5, 7, 9):
        if norm1 <= table[q]:
            return q, 0
    s = 0
    if norm1 > table[13]:
        s = max(0, int(math.ceil(math.log2(norm1 / table[13]))))
    return 13, s


def expm_higham2005(A: np.ndarray, table=THETA_EXPM) -> np.ndarray:
    """Scaling-and-squaring Pade expm as called at src/gradient_computations.jl:24 (no balancing: gebal is a
    permutation-only no-op for the normal (skew-Hermitian) generators of this path; documented in DESIGN.md)."""
    A = np.array(A, dtype=np.complex128)
    n = A.shape[0]
    I = np.eye(n)
    nA = np.linalg.norm(A, 1)
    q, s = select_degree(nA, table)
    if q < 13:
        C = PADE_B[q]
        A2 = A @ A
        P = A2.copy()
        U = C[3] * P + C[1] * I
        V = C[2] * P + C[0] * I
        for k in range(2, len(C) // 2):
            P = P @ A2
            U = U + C[2 * k + 1] * P
            V = V + C[2 * k] * P
        U = A @ U
        return np.linalg.solve(V - U, V + U)
    if s > 0:
        A = A / 2.0 ** s
    C = PADE_B[13]
    A2 = A @ A
    A4 = A2 @ A2
    A6 = A2 @ A4
    U = A @ (A6 @ (C[13] * A6 + C[11] * A4 + C[9] * A2) + C[7] * A6 + C[5] * A4 + C[3] * A2 + C[1] * I)
    V = A6 @ (C[12] * A6 + C[10] * A4 + C[8] * A2) + C[6] * A6 + C[4] * A4 + C[2] * A2 + C[0] * I
    X = np.linalg.solve(V - U, V + U)
    for _ in range(s):
        X = X @ X
    return X


def expm_frechet_blocktri(X: np.ndarray, E: np.ndarray):
    """Exact Frechet derivative via Mathias' identity exp([[X,E],[0,X]]) = [[e^X, L(X,E)],[0,e^X]] (the
    'block-triangular augmented-matrix form' of the north star).  -> (expm(X), L(X,E))."""
    n = X.shape[0]
    M = np.zeros((2 * n, 2 * n), dtype=np.complex128)
    M[:n, :n] = X
    M[n:, n:] = X
    M[:n, n:] = E
    F = sla.expm(M)
    return F[:n, :n], F[:n, n:]


def expm_frechet_sps(A, E, table=ELL_FRECHET):
    """Al-Mohy & Higham 2009 Alg. 6.4 restated (structured block-triangular evaluation: every Pade product
    on [[A,E],[0,A]] costs the shared A-product plus two E-products).  This is the arithmetic the CUDA
    kernel's exact mode performs; -> (R, L).  Cross-checked against scipy.linalg.expm_frechet in tests."""
    A = np.array(A, dtype=np.complex128)
    E = np.array(E, dtype=np.complex128)
    n = A.shape[0]
    I = np.eye(n)
    nA = np.linalg.norm(A, 1)
    q, s = select_degree(nA, table)
    if s > 0:
        A = A / 2.0 ** s
        E = E / 2.0 ** s
    b = PADE_B[q]
    A2 = A @ A
    M2 = A @ E + E @ A
    if q < 13:
        # powers A^{2k}, M_{2k} = L_{x^{2k}}(A,E)
        Ap = [None, A2]
        Mp = [None, M2]
        for k in range(2, (q - 1) // 2 + 1):
            Ap.append(Ap[-1] @ A2)
            Mp.append(Mp[-1] @ A2 + Ap[k - 1] @ M2)
        W = b[1] * I
        V = b[0] * I
        Lw = np.zeros_like(A)
        Lv = np.zeros_like(A)
        for k in range(1, (q - 1) // 2 + 1):
            W = W + b[2 * k + 1] * Ap[k]
            V = V + b[2 * k] * Ap[k]
            Lw = Lw + b[2 * k + 1] * Mp[k]
            Lv = Lv + b[2 * k] * Mp[k]
        U = A @ W
        Lu = A @ Lw + E @ W
    else:
        A4 = A2 @ A2
        A6 = A2 @ A4
        M4 = A2 @ M2 + M2 @ A2
        M6 = A4 @ M2 + M4 @ A2
        W1 = b[13] * A6 + b[11] * A4 + b[9] * A2
        W2 = b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * I
        Z1 = b[12] * A6 + b[10] * A4 + b[8] * A2
        Z2 = b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * I
        W = A6 @ W1 + W2
        U = A @ W
        V = A6 @ Z1 + Z2
        Lw1 = b[13] * M6 + b[11] * M4 + b[9] * M2
        Lw2 = b[7] * M6 + b[5] * M4 + b[3] * M2
        Lz1 = b[12] * M6 + b[10] * M4 + b[8] * M2
        Lz2 = b[6] * M6 + b[4] * M4 + b[2] * M2
        Lw = A6 @ Lw1 + M6 @ W1 + Lw2
        Lu = A @ Lw + E @ W
        Lv = A6 @ Lz1 + M6 @ Z1 + Lz2
    lu_piv = sla.lu_factor(V - U)
    R = sla.lu_solve(lu_piv, U + V)
    L = sla.lu_solve(lu_piv, Lu + Lv + (Lu - Lv) @ R)
    for _ in range(s):
        L = R @ L + L @ R
        R = R @ R
    return R, L


# --------------------------------------------------------------------------------------------------------
# src/gradient_computations.jl  (exp path)
# --------------------------------------------------------------------------------------------------------


def generator(A0, A, p):
    """src/gradient_computations.jl:19-22 / :188-191."""
    X = np.array(A0, dtype=np.complex128)
    for j, Aj in enumerate(A):
        X = X + p[j] * Aj
    return X


def setup_grape_cache(A0, x0, u_size):
    """src/gradient_computations.jl:79-96 (dimension check :84-87)."""
    x0 = np.asarray(x0)
    x0m = x0.reshape(x0.shape[0], -1)
    d = A0.shape[0]
    if (np.isrealobj(x0) and x0m.shape[0] != 2 * d) or (np.iscomplexobj(x0) and x0m.shape[0] != d):
        raise ValueError("Error when creating cache, A0 and x0 have incompatiable dimensions")
    nc, Nt = u_size
    m = x0m.shape[1]
    return dict(x=np.zeros((Nt + 1, d, m), np.complex128), lam=np.zeros((Nt + 1, d, m), np.complex128),
                dJdu=np.zeros((nc, Nt)), Uk=np.zeros((Nt, d, d), np.complex128), u=np.full((nc, Nt), np.nan))


def propagate(A0, A, u, x0, cache=None, expm=expm_higham2005):
    """src/gradient_computations.jl:2-32.  x0 may be real (promoted, :4,:8).  -> cache dict (x = cache['x'])."""
    u = np.asarray(u, dtype=np.float64)
    x0 = np.asarray(x0, dtype=np.complex128)
    x0 = x0.reshape(x0.shape[0], -1)
    Nt = u.shape[1]
    if cache is None:
        cache = setup_grape_cache(A0, x0, u.shape)
    cache["u"][...] = u
    x, Uk = cache["x"], cache["Uk"]
    x[0] = x0
    for k in range(Nt):  # Threads.@threads in the reference (:17-25)
        Uk[k] = expm(generator(A0, A, u[:, k]))
    for k in range(Nt):  # :27-29
        x[k + 1] = Uk[k] @ x[k]
    return cache


def expm_jacobian(A0, A, p, order=2, dt=1.0):
    """src/gradient_computations.jl:177-213 truncated Taylor series of d exp(dt*X)/dp_j, same association
    order as the reference (AjX, XAj first, then right/left multiplications)."""
    out = [dt * np.array(Aj, dtype=np.complex128) for Aj in A]
    if order <= 1:
        return out
    X = generator(A0, A, p)
    for j, Aj in enumerate(A):
        AjX = Aj @ X
        XAj = X @ Aj
        if order >= 2:
            out[j] = out[j] + (dt ** 2 / 2) * (AjX + XAj)
        if order >= 3:
            out[j] = out[j] + (dt ** 3 / 6) * (AjX @ X)
            out[j] = out[j] + (dt ** 3 / 6) * (XAj @ X)
            out[j] = out[j] + (dt ** 3 / 6) * (X @ XAj)
        if order >= 4:
            X2 = X @ X
            out[j] = out[j] + (dt ** 4 / 24) * (AjX @ X2)
            out[j] = out[j] + (dt ** 4 / 24) * (XAj @ X2)
            out[j] = out[j] + (dt ** 4 / 24) * (X2 @ AjX)
            out[j] = out[j] + (dt ** 4 / 24) * (X2 @ XAj)
    return out


def compute_u_sensitivity(xk, lam_kp1, dU):
    """src/gradient_computations.jl:217-223  sum_l Re(dot(lam[:,l], dU, x[:,l]))."""
    return float(np.real(np.sum(lam_kp1.conj() * (dU @ xk))))


FRECHET = 0  # order value selecting the exact Frechet derivative (the reference has orders 1..4 only)


def grape_sensitivity(A0, A, dJfinal_dx, u, x0, cache, dUkdp_order=3, dL_dx=None):
    """src/gradient_computations.jl:35-77.  dUkdp_order in {1,2,3,4} follows the reference;
    dUkdp_order == FRECHET (0) uses the exact Frechet derivative (north-star mode)."""
    u = np.asarray(u, dtype=np.float64)
    if not np.array_equal(u, cache["u"]):
        raise RuntimeError("Cache data from other control signal u")  # :37-39
    x, lam, dJdu, Uk = cache["x"], cache["lam"], cache["dJdu"], cache["Uk"]
    Nt = u.shape[1]
    lam[Nt] = dJfinal_dx(x[Nt])
    if dL_dx is not None:
        lam[Nt] = lam[Nt] + dL_dx(x[Nt])
    for k in range(Nt - 1, -1, -1):  # :52-58
        lam[k] = Uk[k].conj().T @ lam[k + 1]
        if dL_dx is not None:
            lam[k] = lam[k] + dL_dx(x[k])
    for k in range(Nt - 1, -1, -1):  # :65-74
        if dUkdp_order == FRECHET:
            X = generator(A0, A, u[:, k])
            dU = [expm_frechet_sps(X, Aj)[1] for Aj in A]
        else:
            dU = expm_jacobian(A0, A, u[:, k], dUkdp_order)
        for j in range(len(A)):
            dJdu[j, k] = compute_u_sensitivity(x[k], lam[k + 1], dU[j])
    return dJdu


# --------------------------------------------------------------------------------------------------------
# src/penalty_fcns.jl
# --------------------------------------------------------------------------------------------------------


def setup_state_penalty(inds_penalty, inds_css, mu):
    """src/penalty_fcns.jl:1-11 (0-based index lists here)."""
    ip = np.asarray(inds_penalty)
    ic = np.asarray(inds_css)

    def L(x):
        return float(mu * np.sum(np.abs(x[np.ix_(ip, ic)]) ** 2))

    def dL_dx(x):
        g = np.zeros_like(x)
        g[np.ix_(ip, ic)] = 2 * mu * x[np.ix_(ip, ic)]
        return g

    return L, dL_dx


def setup_infidelity(x_target, n=None):
    """src/penalty_fcns.jl:15-24."""
    T = np.asarray(x_target, dtype=np.complex128)
    T = T.reshape(T.shape[0], -1)
    if n is None:
        n = T.shape[1]

    def J(x):
        return float(1 - abs(np.vdot(T, x)) ** 2 / n ** 2)  # tr(T'x) = sum conj(T).*x

    def dJ_dx(x):
        om = np.vdot(T, x)
        return (-2 * om / n ** 2) * T

    return J, dJ_dx


def setup_infidelity_abs_trace(x_target):
    """test/test_gradient_computation.jl:24-25: J = 1 - |tr(T'x)|, gradient in the Zygote convention
    g = dJ/dRe(x) + i dJ/dIm(x) = -(Omega/|Omega|) T."""
    T = np.asarray(x_target, dtype=np.complex128)
    T = T.reshape(T.shape[0], -1)

    def J(x):
        return float(1 - abs(np.vdot(T, x)))

    def dJ_dx(x):
        om = np.vdot(T, x)
        return -(om / abs(om)) * T

    return J, dJ_dx


def setup_infidelity_zcalibrated(x_target):
    """src/penalty_fcns.jl:27-42."""
    T = np.asarray(x_target, dtype=np.complex128)
    if T.ndim != 2 or T.shape[1] != 4:
        raise ValueError("Only works for two-qubit gates, x_target must have four columns")

    def J(x):
        m = np.sum(T.conj() * x, axis=0)  # diag(T'x)
        return float(1 - abs_sum_phase_calibrated(m) ** 2 / 16)

    def dJ_dx(x):
        m = np.sum(T.conj() * x, axis=0)
        F, dF_dm = abs_sum_phase_calibrated_rrule(m)
        return (-2 * F / 16) * T * dF_dm[None, :]

    return J, dJ_dx


# --------------------------------------------------------------------------------------------------------
# src/fidelities.jl
# --------------------------------------------------------------------------------------------------------


def _cis(t):
    return complex(math.cos(t), math.sin(t))


def _angle(z):
    return math.atan2(z.imag, z.real)


def _mod2pi(x):
    return x % (2 * math.pi)


def golden_section_search(f, lo, hi, x_tol):
    """src/fidelities.jl:105-137."""
    if lo > hi:
        raise ValueError(f"x_lower must be less than x_upper ({lo}, {hi})")
    gr = 0.5 * (3.0 - math.sqrt(5.0))
    xm = lo + gr * (hi - lo)
    fm = f(xm)
    while hi - lo >= x_tol:
        if hi - xm > xm - lo:
            xn = xm + gr * (hi - xm)
            fn = f(xn)
            if fn < fm:
                lo, xm, fm = xm, xn, fn
            else:
                hi = xn
        else:
            xn = xm - gr * (xm - lo)
            fn = f(xn)
            if fn < fm:
                hi, xm, fm = xm, xn, fn
            else:
                lo = xn
    return fm, xm


def optimal_calibration(m, theta_tol=1e-9):
    """src/fidelities.jl:81-101 -> (F, [theta1, theta2])."""
    m = [complex(z) for z in m]
    a1 = abs(m[0]) ** 2 + abs(m[1]) ** 2
    b1 = 2 * abs(m[0]) * abs(m[1])
    a2 = abs(m[2]) ** 2 + abs(m[3]) ** 2
    b2 = 2 * abs(m[2]) * abs(m[3])
    p1 = _mod2pi(_angle(m[0]) - _angle(m[1]))
    p2 = _mod2pi(_angle(m[2]) - _angle(m[3]))
    if abs(p2 - p1) <= math.pi:
        pm, D, al = (p1 + p2) / 2, abs(p2 - p1) / 2, (1 if p1 < p2 else -1)
    else:
        pm, D, al = (2 * math.pi + p1 + p2) / 2, math.pi - abs(p2 - p1) / 2, (-1 if p1 < p2 else 1)

    def Jf(dl):
        # max(.,0) guards sqrt against -1e-17 round-off when a == b (|m1| == |m2|); Julia would throw there
        return math.sqrt(max(a1 + b1 * math.cos(dl + D), 0.0)) + math.sqrt(max(a2 + b2 * math.cos(dl - D), 0.0))

    minusJ, d_opt = golden_section_search(lambda dl: -Jf(dl), -D, D, theta_tol)
    t1 = pm + al * d_opt
    t2 = _angle(m[0] + m[1] * _cis(t1)) - _angle(m[2] + m[3] * _cis(t1))
    return -minusJ, [t1, t2]


def basic_calibration(m):
    """src/fidelities.jl:65-69."""
    t0 = _angle(m[0])
    th = [-(_angle(m[1]) - t0), -(_angle(m[2]) - t0)]
    return abs(m[0] + m[1] * _cis(th[0]) + m[2] * _cis(th[1]) + m[3] * _cis(th[0] + th[1])), th


def grid_calibration(m):
    """src/fidelities.jl:72-79."""
    best, tb = -1.0, 0.0
    for t in np.linspace(0, 2 * math.pi, 100):
        v = abs(m[0] + m[1] * _cis(t)) + abs(m[2] + m[3] * _cis(t))
        if v > best:
            best, tb = v, t
    return best, tb


def abs_sum_phase_calibrated(m, calibration="optimal"):
    """src/fidelities.jl:11-40."""
    m = [complex(z) for z in m]
    if calibration == "lms_phase":
        t1 = -_angle(m[0].conjugate() * m[1] + m[2].conjugate() * m[3])
        return abs(m[0] + m[1] * _cis(t1)) + abs(m[2] + m[3] * _cis(t1))
    if calibration == "lms_phase2":
        x1, x2 = math.sqrt(abs(m[0] * m[1])), math.sqrt(abs(m[2] * m[3]))
        eps = np.finfo(float).eps
        if x1 < eps or x2 < eps:
            return abs(m[0]) + abs(m[1]) + abs(m[2]) + abs(m[3])
        t = -_angle(m[0].conjugate() * m[1] / x1 + m[2].conjugate() * m[3] / x2)
        return abs(m[0] + m[1] * _cis(t)) + abs(m[2] + m[3] * _cis(t))
    if calibration == "lms_phase3":
        x1, x2 = abs(m[0]) + abs(m[1]), abs(m[2]) + abs(m[3])
        t = -_angle(m[0].conjugate() * m[1] / x1 + m[2].conjugate() * m[3] / x2)
        return abs(m[0] + m[1] * _cis(t)) + abs(m[2] + m[3] * _cis(t))
    if calibration == "optimal":
        return optimal_calibration(m)[0]
    if calibration == "basic":
        return basic_calibration(m)[0]
    if calibration == "none":
        return abs(sum(m))
    if calibration == "grid":
        return grid_calibration(m)[0]
    return None  # the reference falls through and returns nothing for unknown symbols


def abs_sum_phase_calibrated_grad(m, theta1_opt):
    """src/fidelities.jl:42-46 (gradient of F^2, i.e. 2F * dF/dm)."""
    v1 = m[0] + _cis(theta1_opt) * m[1]
    v2 = m[2] + _cis(theta1_opt) * m[3]
    s = 2 * (abs(v1) + abs(v2))
    return s * np.array([v1 / abs(v1), v1 / abs(v1) * _cis(-theta1_opt), v2 / abs(v2),
                         v2 / abs(v2) * _cis(-theta1_opt)])


def abs_sum_phase_calibrated_rrule(m):
    """src/fidelities.jl:48-56 -> (F, dF_dm) (envelope theorem: theta held at its optimum)."""
    y, th = optimal_calibration(m)
    v1 = m[0] + _cis(th[0]) * m[1]
    v2 = m[2] + _cis(th[0]) * m[3]
    dF = np.array([v1 / abs(v1), v1 / abs(v1) * _cis(-th[0]), v2 / abs(v2), v2 / abs(v2) * _cis(-th[0])])
    return y, dF


def abs_trace_phase_calibrated(M, calibration="optimal"):
    """src/fidelities.jl:9."""
    return abs_sum_phase_calibrated(np.diag(M), calibration)


def infidelity(U_target, Uf, calibration="lms_phase"):
    """src/fidelities.jl:1-7 (4x4 only)."""
    U_target = np.asarray(U_target)
    if U_target.shape != (4, 4):
        raise ValueError("Not supported yet")
    return 1 - abs_trace_phase_calibrated(U_target.conj().T @ np.asarray(Uf), calibration) / 4


def cost_closures(cfg):
    if cfg["cost"] == COST_INFIDELITY:
        return setup_infidelity(cfg["T"], cfg["n"])
    if cfg["cost"] == COST_ABS_TRACE:
        return setup_infidelity_abs_trace(cfg["T"])
    raise ValueError("no built-in cost")


def evaluate(cfg, order=3, u=None, penalty=None):
    """One full fidelity + gradient evaluation through the restated reference path.
    -> (J_total, dJdu, cache).  penalty = (inds_penalty, inds_css, mu) adds the running state penalty
    exactly as examples/ipopt_callbacks_exp.jl:18,27 does (sum over all Nt+1 states)."""
    u = cfg["u"] if u is None else u
    J, dJ = cost_closures(cfg)
    cache = propagate(cfg["A0"], cfg["A"], u, cfg["x0"])
    Jv = J(cache["x"][-1])
    dL = None
    if penalty is not None:
        L, dL = setup_state_penalty(*penalty)
        Jv += sum(L(xk) for xk in cache["x"])
    g = grape_sensitivity(cfg["A0"], cfg["A"], dJ, u, cfg["x0"], cache, dUkdp_order=order, dL_dx=dL)
    return Jv, g.copy(), cache


def f_alg(d, m, nc, q, s, order):
    """Algorithmic flops per slice*pulse, SURVEY.md section 8(d)."""
    pi = {3: 2, 5: 3, 7: 4, 9: 5, 13: 6}[q]
    M = 8.0 * d ** 3
    G = (2 * pi + 2 * s + 2) if order == FRECHET else {1: 0, 2: 2, 3: 5, 4: 10}[order]
    return M * ((pi + s + 4.0 / 3.0) + nc * G) + 8.0 * d * d * m * (2 + nc) + 4.0 * nc * d * d
