"""Input generators for the BASELINE.json configs: the reference's example models, pulse
shapes and target states restated as numpy arrays.  Host-side input preparation only (no propagation, no expm):
shared by bench.py, the tests and the oracle so that every consumer sees identical bits.  All randomness is
numpy default_rng(seed).  Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
import os
from itertools import product as _iproduct

import numpy as np

# --------------------------------------------------------------------------------------------------------
# src/utils.jl
# --------------------------------------------------------------------------------------------------------


def annihilation_op(dim: int) -> np.ndarray:
    """src/utils.jl:66  diagm(1 => [sqrt(k) for k=1:dim-1])"""
    return np.diag(np.sqrt(np.arange(1, dim, dtype=np.float64)), k=1)


def annihilation_ops(*dims: int) -> list[np.ndarray]:
    """src/utils.jl:67-71  kron of identities with a_j at position j (first subsystem most significant)."""
    a_vec = [annihilation_op(n) for n in dims]
    out = []
    for j in range(len(dims)):
        m = np.eye(1)
        for k in range(len(dims)):
            m = np.kron(m, a_vec[k] if k == j else np.eye(dims[k]))
        out.append(m)
    return out


class QuantumBasis:
    """src/utils.jl:35-63.  state_dict maps digit-string labels to 1-based indices, first subsystem most
    significant (kron of the per-subsystem digit strings, :42).  Here indices are 0-based."""

    def __init__(self, dims):
        self.dims = list(dims)
        labels = ["".join(str(x) for x in t) for t in _iproduct(*[range(n) for n in dims])]
        self.state_dict = {lab: i for i, lab in enumerate(labels)}
        self.Ntot = int(np.prod(dims))

    def __call__(self, s):
        if isinstance(s, str):
            return self.state_dict[s]
        return [self.state_dict[x] for x in s]

    def columns(self, cols) -> np.ndarray:
        """qb[:, cols]  (src/utils.jl:47-51): the identity's columns for the given labels."""
        idx = self(cols) if not isinstance(cols, str) else [self(cols)]
        return np.eye(self.Ntot)[:, idx]


def setup_bilinear_matrices(H0, Tc, dt=1.0):
    """src/utils.jl:86-91."""
    H0 = np.asarray(H0, dtype=np.complex128)
    Tc = np.asarray(Tc, dtype=np.complex128)
    A0 = -1j * H0 * dt
    A1 = -1j * (Tc + Tc.conj().T) * dt
    A2 = -1j * (1j * (Tc - Tc.conj().T)) * dt
    return A0, A1, A2


# --------------------------------------------------------------------------------------------------------
# src/parameterized_pulses.jl (only what the named configs need)
# --------------------------------------------------------------------------------------------------------


def cos_envelope(t_plateau, t_rise_fall, t):
    """src/parameterized_pulses.jl:27-35."""
    if t_rise_fall / 2 < t <= t_rise_fall / 2 + t_plateau:
        return 1.0
    if t <= t_rise_fall / 2:
        return 0.5 * (1 - math.cos(2 * math.pi * t / t_rise_fall))
    return 0.5 * (1 - math.cos(2 * math.pi * (t - t_plateau) / t_rise_fall))


# --------------------------------------------------------------------------------------------------------
# examples/models/*.jl  (input generators for the named configs)
# --------------------------------------------------------------------------------------------------------


def model_zz_coupling():
    """examples/models/zz_coupling.jl:6-27 -> (H0, Tc, qb)."""
    dimq = dims = 3
    alpha_q = 2 * math.pi * 0.2
    alpha_s = 2 * math.pi * 0.2
    chi = 2 * math.pi * 1e-4
    a_q, a_s = annihilation_op(dimq), annihilation_op(dims)
    Iq, Is = np.eye(dimq), np.eye(dims)
    Hq = -alpha_q / 2 * np.kron(a_q.T @ a_q.T @ a_q @ a_q, Is)
    Hs = -alpha_s / 2 * np.kron(Iq, a_s.T @ a_s.T @ a_s @ a_s)
    Hint = -chi * np.kron(a_q.T @ a_q, a_s.T @ a_s)
    Tc = np.kron(a_q.T, Is)
    return Hq + Hs + Hint, Tc, QuantumBasis([dimq, dims])


def model_two_qubit_tunable_bus():
    """examples/models/two_qubit_tunable_bus.jl:7-28 -> (H0, Hc, qb)."""
    tp = 2 * math.pi
    w1, w2, wc0 = 4.5 * tp, 4.2 * tp, 7.5 * tp
    al1, al2 = -0.2 * tp, -0.2 * tp
    g1, g2 = 0.04 * tp, 0.04 * tp
    qb = QuantumBasis([3, 3, 3])
    a1, a2, ac = annihilation_ops(3, 3, 3)
    I = np.eye(27)
    n1, n2 = a1.T @ a1, a2.T @ a2
    Hq1 = w1 * n1 + al1 * n1 @ (n1 - I)
    Hq2 = w2 * n2 + al2 * n2 @ (n2 - I)
    Hi1 = g1 * (a1.T + a1) @ (ac.T + ac)
    Hi2 = g2 * (a2.T + a2) @ (ac.T + ac)
    Hc = wc0 * ac.T @ ac
    return Hq1 + Hq2 + Hi1 + Hi2, Hc, qb


def bus_envelope(p, t):
    """examples/two_qubit_tunable_bus.jl:10-18."""
    t_plateau, t_rise_fall, th0, w_phi, A = p
    delta = cos_envelope(t_plateau, t_rise_fall, t)
    phi = th0 + A * delta * math.cos(w_phi * t)
    return math.sqrt(abs(math.cos(math.pi * phi)))


def bus_pulse_parameters(H0, qb):
    """examples/two_qubit_tunable_bus.jl:22-34 -> p0."""
    i1, i2 = qb("110"), qb("200")
    w_th = abs(H0[i1, i1] - H0[i2, i2])
    w_phi = w_th + (-0.002) * 2 * math.pi
    return [300.0, 50.0, 0.25, w_phi, 0.13]


CAVITY_THETA = [3.6348672, 1.1435776, 0.0, 1.7441809, -0.4598031, -0.37506938, -0.27870846,
                0.0, 0.0, 0.0, 0.0, 0.0]


def model_cavity_qubit(N_cavity=12, N_qubit=2):
    """examples/models/cavity_qubit.jl:6-49 -> (H0, Tc, x0, theta) ; all coefficients except xi are zero."""
    xi = 2 * math.pi * (-2.574749e-3)
    a, b = annihilation_op(N_cavity), annihilation_op(N_qubit)
    H0 = xi * np.kron(b.T @ b, a.T @ a)
    Tc = np.kron(b.T, np.eye(N_cavity))
    x0 = np.kron(np.eye(N_qubit)[:, 0], np.ones(N_cavity) / math.sqrt(N_cavity))
    theta = np.zeros(N_cavity)
    nth = min(N_cavity, len(CAVITY_THETA))
    theta[:nth] = CAVITY_THETA[:nth]
    return H0, Tc, x0, theta


def load_pulse_csv(path):
    """test/test_gradient_computation.jl:7-11: whitespace-separated I/Q rows in Hz -> u (2 x N) in GHz."""
    iq = np.loadtxt(path)
    return 1e-9 * iq.T.copy()


def bspline_matrix(tgate=10.0, segment_count=100, nsplines=10):
    """examples/zz_coupling_ipopt_exp.jl:29-38: cubic B-splines (order 4) on nsplines+4 uniform breakpoints,
    evaluated at slice midpoints, columns 4..end-3 (1-based) kept -> (segment_count x nsplines)."""
    from scipy.interpolate import BSpline

    k = 3
    brk = np.linspace(0.0, tgate, nsplines + 4)
    knots = np.concatenate([[brk[0]] * k, brk, [brk[-1]] * k])
    nb = len(knots) - k - 1  # nsplines + 6
    dt = tgate / segment_count
    tm = np.arange(segment_count) * dt + dt / 2
    Bpre = np.zeros((segment_count, nb))
    for i in range(nb):
        c = np.zeros(nb)
        c[i] = 1.0
        Bpre[:, i] = BSpline(knots, c, k, extrapolate=False)(tm)
    return np.nan_to_num(Bpre)[:, 3:nb - 3]


COST_INFIDELITY = 0   # 1 - |tr(T'x)|^2 / n^2        src/penalty_fcns.jl:15-24
COST_ABS_TRACE = 1    # 1 - |tr(T'x)|                test/test_gradient_computation.jl:24
COST_NONE = 2         # terminal costate supplied by the caller (host closure)

def config_zz(seed=0, Nt=100, tgate=10.0, noise=0.05, coeffs=None):
    """C1: examples/zz_coupling_ipopt_exp.jl (d=9, m=4, nc=2, Nt=100, dt=0.1), T = Q_css * (X (x) I), n=4."""
    H0, Tc, qb = model_zz_coupling()
    dt = tgate / Nt
    A0, A1, A2 = setup_bilinear_matrices(H0, Tc, dt)
    Q = qb.columns(["00", "01", "10", "11"])
    css_target = np.kron(np.array([[0, 1], [1, 0]]), np.eye(2))
    T = (Q @ css_target).astype(np.complex128)
    B = bspline_matrix(tgate, Nt, 10)
    if coeffs is None:
        c0 = np.concatenate([0.01 * np.ones(10), np.zeros(10)])
    else:
        c0 = np.asarray(coeffs, dtype=np.float64)
    u = (B @ c0.reshape(2, 10).T).T
    if noise:
        u = u + noise * np.random.default_rng(seed).standard_normal((2, Nt))
    return dict(name="zz", A0=A0, A=[A1, A2], u=np.ascontiguousarray(u), x0=Q.astype(np.complex128), T=T,
                cost=COST_INFIDELITY, n=4, B=B, qb=qb, dt=dt)


def config_zz_batch(batch, seed0=1, Nt=100):
    """C4: `batch` random two-qubit pulses of C1 shape, spline coefficients U(-2pi*0.06, 2pi*0.06)^20
    (the Ipopt box bounds, examples/zz_coupling_ipopt_exp.jl:54-56); pulse b uses seed seed0+b."""
    base = config_zz(noise=0.0, Nt=Nt)
    B = base["B"]
    lim = 2 * math.pi * 0.060
    us = np.zeros((batch, 2, Nt))
    for b in range(batch):
        c = np.random.default_rng(seed0 + b).uniform(-lim, lim, 20)
        us[b] = (B @ c.reshape(2, 10).T).T
    base["u_batch"] = us
    base["name"] = "zz_batch"
    return base


def config_bus(Nt=10000, tgate=350.0):
    """C2: two_qubit_tunable_bus.jl, d=27, m=1, nc=1, x0=|110>, target |200>, midpoint-sampled envelope."""
    H0, Hc, qb = model_two_qubit_tunable_bus()
    dt = tgate / Nt
    p0 = bus_pulse_parameters(H0, qb)
    u = np.array([[bus_envelope(p0, (k + 0.5) * dt) for k in range(Nt)]])
    A0 = -1j * H0 * dt
    A1 = -1j * Hc * dt
    x0 = qb.columns("110").astype(np.complex128)
    T = qb.columns("200").astype(np.complex128)
    return dict(name="bus", A0=A0.astype(np.complex128), A=[A1.astype(np.complex128)], u=u, x0=x0, T=T,
                cost=COST_INFIDELITY, n=1, qb=qb, dt=dt)


def config_cavity(N_cavity=12, Nt=550, csv_path=None):
    """C3: test/test_gradient_computation.jl:7-35 set-up (Tc/2, two state columns, J = 1-|tr(T'x)|)."""
    H0, Tc, _x0, theta = model_cavity_qubit(N_cavity)
    if csv_path is None:
        # the reference's shipped pulse (examples/cavity_qubit_pulse_marina.csv), kept as package data
        csv_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "cavity_qubit_pulse.txt")
    u = load_pulse_csv(csv_path)[:, :Nt]
    A0, A1, A2 = setup_bilinear_matrices(H0, Tc / 2, 1.0)
    nrm = lambda v: v / np.linalg.norm(v)
    ones, zeros = np.ones(N_cavity), np.zeros(N_cavity)
    x0 = np.stack([nrm(np.concatenate([ones, zeros])), nrm(np.concatenate([zeros, ones]))], axis=1)
    T = np.stack([nrm(np.kron([1, 1], np.exp(1j * theta))), nrm(np.concatenate([zeros, ones]).astype(complex))],
                 axis=1)
    return dict(name=f"cavity{N_cavity}", A0=A0, A=[A1, A2], u=np.ascontiguousarray(u),
                x0=x0.astype(np.complex128), T=T.astype(np.complex128), cost=COST_ABS_TRACE, n=2, dt=1.0,
                theta=theta)


def config_synthetic(d, Nt, nc=2, m=4, seed=None):
    """C5: GUE-random Hermitian H0,H1,H2 scaled to ||A0||_1 = 2, ||A_j||_1 = 1; u ~ U(-0.5, 0.5);
    x0 = I[:, :m]; T = first m columns of a Haar-random unitary; seed = d."""
    rng = np.random.default_rng(d if seed is None else seed)

    def gue():
        g = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
        return (g + g.conj().T) / 2

    A0 = -1j * gue()
    A0 *= 2.0 / np.linalg.norm(A0, 1)
    A = []
    for _ in range(nc):
        Aj = -1j * gue()
        A.append(Aj / np.linalg.norm(Aj, 1))
    u = rng.uniform(-0.5, 0.5, (nc, Nt))
    x0 = np.eye(d, dtype=np.complex128)[:, :m]
    g = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
    Q, R = np.linalg.qr(g)
    Q = Q * (np.diag(R) / np.abs(np.diag(R)))[None, :]
    T = np.ascontiguousarray(Q[:, :m])
    return dict(name=f"synth{d}", A0=A0, A=A, u=u, x0=x0, T=T, cost=COST_INFIDELITY, n=m, dt=1.0)


