"""CPU oracle for the reference's ODE path: propagate_pwc / compute_pwc_gradient,
src/gradient_computations.jl:99-169, driven as in test/test_gradient_computation.jl:41-54.

TEST INFRASTRUCTURE ONLY (same rules as qoc_oracle.py).  A second, independent reference for the exp path: the
states and costates come from a fixed-step Tsit5 integration (dt = 0.1 * slice length, adaptive=false) of
    dx/dt = (A0 + sum_j u_j(t) A_j) x                    examples/models/setup_diffeq_rhs.jl:3-15
    dl/dt = -(A0' + sum_j u_j(t) A_j') l   (backwards)   examples/models/setup_diffeq_rhs.jl:18-32
with piecewise-constant u, and the gradient uses the same truncated-Taylor expm_jacobian! and
_compute_u_sensitivity as the exp path (:159-166).  The reference integrates the 2N-real "c2r" form; integrating the
complex form is the same arithmetic.  A0, A are per UNIT time here (the exp path gets them pre-multiplied by dt).

PARITY PINNING: the reference's test only displays the two gradients side by side (no assertion).  Tsit5 is 5th order:
with h = 0.1 slice the two paths agree to O(h^5) per step -- the tests state that tolerance (1e-6 relative).
The Tsit5 tableau is restated from Tsitouras 2011 (OrdinaryDiffEq's Tsit5ConstantCache); `tableau_check()` verifies row sums,
the order conditions up to 4 and the measured convergence order.
"""
from __future__ import annotations

import numpy as np

import qoc_oracle as o

C2, C3, C4, C5 = 0.161, 0.327, 0.9, 0.9800255409045097
A = {
    (2, 1): 0.161,
    (3, 1): -0.008480655492356989, (3, 2): 0.335480655492357,
    (4, 1): 2.8971530571054935, (4, 2): -6.359448489975075, (4, 3): 4.3622954328695815,
    (5, 1): 5.325864828439257, (5, 2): -11.748883564062828, (5, 3): 7.4955393428898365, (5, 4): -0.09249506636175525,
    (6, 1): 5.86145544294642, (6, 2): -12.92096931784711, (6, 3): 8.159367898576159, (6, 4): -0.071584973281401,
    (6, 5): -0.028269050394068383,
}
B = (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774)
C = (0.0, C2, C3, C4, C5, 1.0)


def tsit5_step(f, x, h):
    """One explicit Tsit5 step for an autonomous right-hand side (u is constant inside a slice)."""
    k = [f(x)]
    for i in range(2, 7):
        xi = x + h * sum(A[(i, j)] * k[j - 1] for j in range(1, i))
        k.append(f(xi))
    return x + h * sum(b * ki for b, ki in zip(B, k))


def tableau_check():
    rows = {i: sum(A[(i, j)] for j in range(1, i)) for i in range(2, 7)}
    c = np.array(C); b = np.array(B)
    return {"row_sum_err": max(abs(rows[i] - C[i - 1]) for i in rows),
            "order_conditions": [abs(b.sum() - 1), abs(b @ c - 0.5), abs(b @ c ** 2 - 1 / 3), abs(b @ c ** 3 - 0.25),
                                 abs(b @ c ** 4 - 0.2)]}


def propagate_pwc(A0, Alist, x0, u, dt_slice, substeps=10):
    """src/gradient_computations.jl:108-132 (saveat the slice boundaries).  Returns x[0..Nt]."""
    x = np.asarray(x0, dtype=np.complex128).copy()
    xs = [x.copy()]
    h = dt_slice / substeps
    for k in range(u.shape[1]):
        X = o.generator(A0, Alist, u[:, k])
        f = lambda y: X @ y
        for _ in range(substeps):
            x = tsit5_step(f, x, h)
        xs.append(x.copy())
    return xs


def compute_pwc_gradient(A0, Alist, dJfinal_dx, xs, u, dt_slice, dUkdp_order=2, substeps=10):
    """src/gradient_computations.jl:135-169.  xs from propagate_pwc.  Returns dJdu (nc x Nt) and the costates."""
    Nt = u.shape[1]
    lam = np.asarray(dJfinal_dx(xs[-1]), dtype=np.complex128)
    lams = [None] * (Nt + 1)
    lams[Nt] = lam.copy()
    h = dt_slice / substeps
    for k in range(Nt - 1, -1, -1):
        Xd = o.generator(A0, Alist, u[:, k]).conj().T
        f = lambda y: -(Xd @ y)
        for _ in range(substeps):
            lam = tsit5_step(f, lam, -h)          # integrating backwards in time
        lams[k] = lam.copy()
    dJdu = np.zeros((len(Alist), Nt))
    for k in range(Nt - 1, -1, -1):
        dU = o.expm_jacobian(A0, Alist, u[:, k], dUkdp_order, dt=dt_slice)     # :160
        for j in range(len(Alist)):
            dJdu[j, k] = o.compute_u_sensitivity(xs[k], lams[k + 1], dU[j])    # :163
    return dJdu, lams
