"""The zz_coupling configuration (C1 / C4) is pinned by no reference-held number, and the oracle takes its input generators
from the product's configs.py: a model-construction slip there would be common-mode.  This file restates the model a SECOND,
structurally different way -- explicit matrix elements in the number basis instead of Kronecker products of ladder matrices,
a hand-written Cox-de Boor recursion instead of scipy's BSpline -- straight from the reference's text
(examples/models/zz_coupling.jl:6-27, src/utils.jl:86-91, examples/zz_coupling_ipopt_exp.jl:14-38,62) and requires the two
constructions to agree.  The other two named models are pinned by the reference's own known answers (tests/test_oracle.py:
cavity overlap 0.999979, bus population 0.937218); they get the same element-wise restatement here anyway."""
import math

import numpy as np

import qoc_oracle as o


def _zz_elements(dt):
    dq = ds = 3
    aq = as_ = 2 * math.pi * 0.2
    chi = 2 * math.pi * 1e-4
    d = dq * ds
    H0 = np.zeros((d, d))
    Tc = np.zeros((d, d))
    for nq in range(dq):
        for ns in range(ds):
            i = ds * nq + ns          # kron(q, s): the first subsystem is the most significant digit (src/utils.jl:40-43)
            # a'a'aa |n> = n (n - 1) |n>,  a'a |n> = n |n>
            H0[i, i] = -aq / 2 * nq * (nq - 1) - as_ / 2 * ns * (ns - 1) - chi * nq * ns
            if nq + 1 < dq:
                Tc[ds * (nq + 1) + ns, i] = math.sqrt(nq + 1)     # a_q' (x) I raises the qubit
    A0 = -1j * H0 * dt
    A1 = -1j * (Tc + Tc.T) * dt
    A2 = -1j * (1j * (Tc - Tc.T)) * dt
    return A0, A1, A2


def _cox_de_boor(knots, i, k, t):
    if k == 0:
        return 1.0 if knots[i] <= t < knots[i + 1] else 0.0
    v = 0.0
    if knots[i + k] > knots[i]:
        v += (t - knots[i]) / (knots[i + k] - knots[i]) * _cox_de_boor(knots, i, k - 1, t)
    if knots[i + k + 1] > knots[i + 1]:
        v += (knots[i + k + 1] - t) / (knots[i + k + 1] - knots[i + 1]) * _cox_de_boor(knots, i + 1, k - 1, t)
    return v


def test_zz_model_two_independent_constructions_agree():
    cfg = o.config_zz(noise=0.0)
    A0, A1, A2 = _zz_elements(0.1)
    assert np.abs(cfg["A0"] - A0).max() < 1e-15
    assert np.abs(cfg["A"][0] - A1).max() < 1e-15 and np.abs(cfg["A"][1] - A2).max() < 1e-15
    # u1 A1 + u2 A2 = -i dt (u Tc + conj(u) Tc') with u = u1 + i u2 (src/utils.jl:86-91)
    u1, u2 = 0.3, -0.7
    Tc = np.zeros((9, 9))
    for nq in range(2):
        for ns in range(3):
            Tc[3 * (nq + 1) + ns, 3 * nq + ns] = math.sqrt(nq + 1)
    uc = u1 + 1j * u2
    assert np.abs(u1 * A1 + u2 * A2 - (-1j * 0.1 * (uc * Tc + np.conj(uc) * Tc.T))).max() < 1e-15
    # computational subspace "00","01","10","11" -> 0-based rows 0, 1, 3, 4; target X (x) I on it, n = 4
    Q = np.zeros((9, 4))
    for col, row in enumerate((0, 1, 3, 4)):
        Q[row, col] = 1.0
    X_I = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0]], dtype=float)   # |q s> -> |1-q, s>
    assert np.abs(cfg["x0"] - Q).max() == 0 and np.abs(cfg["T"] - Q @ X_I).max() == 0 and cfg["n"] == 4
    # pulse: cubic B-splines, 14 uniform breakpoints on [0, 10], basis functions 4..13 (1-based) at the slice midpoints,
    # c0 = [0.01 * ones(10); zeros(10)]
    brk = np.linspace(0.0, 10.0, 14)
    knots = np.concatenate([[0.0] * 3, brk, [10.0] * 3])
    tm = (np.arange(100) + 0.5) * 0.1
    B = np.array([[_cox_de_boor(knots, i, 3, t) for i in range(3, 13)] for t in tm])
    assert np.abs(cfg["B"] - B).max() < 1e-14
    u = np.stack([B @ (0.01 * np.ones(10)), np.zeros(100)])
    assert np.abs(cfg["u"] - u).max() < 1e-16


def test_bus_model_elementwise():
    """examples/models/two_qubit_tunable_bus.jl:7-28 by matrix elements in |n1 n2 nc>."""
    cfg = o.config_bus(Nt=8, tgate=0.28)
    tp = 2 * math.pi
    w1, w2, wc0, al, g = 4.5 * tp, 4.2 * tp, 7.5 * tp, -0.2 * tp, 0.04 * tp
    H0 = np.zeros((27, 27))
    Hc = np.zeros((27, 27))
    idx = lambda a, b, c: 9 * a + 3 * b + c
    for n1 in range(3):
        for n2 in range(3):
            for nc in range(3):
                i = idx(n1, n2, nc)
                H0[i, i] = w1 * n1 + al * n1 * (n1 - 1) + w2 * n2 + al * n2 * (n2 - 1)
                Hc[i, i] = wc0 * nc
                # g (a' + a)(ac' + ac) for each qubit: matrix elements sqrt(n)-type in both factors
                for dq in (-1, 1):
                    for dc in (-1, 1):
                        m1, mc = n1 + dq, nc + dc
                        if 0 <= m1 < 3 and 0 <= mc < 3:
                            H0[idx(m1, n2, mc), i] += g * math.sqrt(max(n1, m1)) * math.sqrt(max(nc, mc))
                        m2 = n2 + dq
                        if 0 <= m2 < 3 and 0 <= mc < 3:
                            H0[idx(n1, m2, mc), i] += g * math.sqrt(max(n2, m2)) * math.sqrt(max(nc, mc))
    dt = 0.28 / 8
    assert np.abs(cfg["A0"] - (-1j * H0 * dt)).max() < 1e-13
    assert np.abs(cfg["A"][0] - (-1j * Hc * dt)).max() < 1e-13
    assert cfg["x0"][idx(1, 1, 0), 0] == 1 and cfg["T"][idx(2, 0, 0), 0] == 1
