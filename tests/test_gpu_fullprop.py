"""More than 8 state columns (`-m gpu`): the reference expects full-propagator use, m = d ("if full propagator is computed,
but only subspace is considered", src/penalty_fcns.jl:14).  The library sweeps the columns in chunks of <= 8 that share one
K1 pass; the terminal cost couples them through Omega = tr(T'x) only.  Every shape class, both cost routes (built-in cost on
the device / host closure through lambda_final), batches, the state getters."""
import numpy as np
import pytest

import qoc_oracle as o
import qoc_b200 as q
from qoc_b200 import _lib

pytestmark = pytest.mark.gpu
TOL_J, TOL_G = 1e-10, 1e-8


def full_propagator_case(name):
    if name == "zz9":        # the zz model with all nine columns; the target keeps the computational subspace only
        cfg = o.config_zz()
        d = cfg["A0"].shape[0]
        css = [0, 1, 3, 4]
        T = np.zeros((d, d), dtype=complex)
        T[:, css] = cfg["T"]
        return dict(cfg, x0=np.eye(d, dtype=complex), T=T, n=4)
    if name == "bus27":      # two_qubit_tunable_bus, full 27-column propagator, random unitary target
        cfg = o.config_bus(Nt=300, tgate=10.5)
        d = 27
    elif name == "synth16":  # DMMA class d = 16, m = 16
        cfg = o.config_synthetic(16, 40)
        d = 16
    elif name == "synth32":  # general path, m = 12 (two chunks of 6)
        cfg = o.config_synthetic(32, 24)
        d = 32
    rng = np.random.default_rng(5)
    m = 12 if name == "synth32" else d
    Qm, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return dict(cfg, x0=np.eye(d, dtype=complex)[:, :m].copy(), T=Qm[:, :m].copy(), n=m, cost=o.COST_INFIDELITY)


@pytest.mark.parametrize("name", ["zz9", "bus27", "synth16", "synth32"])
@pytest.mark.parametrize("order", [0, 3])
def test_full_propagator_vs_oracle(name, order):
    cfg = full_propagator_case(name)
    Jo, go, co = o.evaluate(cfg, order=order)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost[1], dUkdp_order=order)
    assert abs(J - Jo) <= TOL_J * max(1.0, abs(Jo))
    assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()
    # the reference-style split calls with a host closure: x[end] (all columns) down, lambda_N up
    Jf, dJf = o.setup_infidelity(cfg["T"], cfg["n"])
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cache)
    assert np.abs(cache.x_final - co["x"][-1]).max() < 1e-11
    assert abs(Jf(cache.x_final) - Jo) <= TOL_J
    g2 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], cache, dUkdp_order=order)
    assert np.abs(g2 - go).max() <= TOL_G * np.abs(go).max()
    # built-in cost through the split calls, and the states / costates of the cache
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cache, Jfinal=cost[0])
    assert abs(cache.J - Jo) <= TOL_J
    g3 = q.grape_sensitivity(cfg["A0"], cfg["A"], cost[1], cfg["u"], cfg["x0"], cache, dUkdp_order=order)
    assert np.abs(g3 - go).max() <= TOL_G * np.abs(go).max()
    assert np.abs(cache.x - co["x"]).max() < 1e-11
    assert np.abs(cache.lam - co["lam"]).max() < 1e-11
    cache.close()


def test_full_propagator_is_unitary_and_batched():
    """m = d: x_N is the propagator itself -> unitary; a batch of pulses equals the pulses one by one."""
    cfg = full_propagator_case("zz9")
    rng = np.random.default_rng(11)
    ub = cfg["u"][None] + 0.02 * rng.standard_normal((5,) + cfg["u"].shape)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    cb = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, batch=5, dUkdp_order=0, store_costates=False)
    Jb, gb = q.evaluate(cb, cfg["A0"], cfg["A"], ub, cfg["x0"], cost[1], dUkdp_order=0)
    q.propagate(cfg["A0"], cfg["A"], ub, cfg["x0"], cb)
    for b in range(5):
        U = cb.x_final[b]
        assert np.abs(U.conj().T @ U - np.eye(9)).max() < 1e-12
        c1 = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=0, store_costates=False)
        J1, g1 = q.evaluate(c1, cfg["A0"], cfg["A"], ub[b], cfg["x0"], cost[1], dUkdp_order=0)
        assert abs(Jb[b] - J1) < 1e-13 and np.abs(gb[b] - g1).max() < 1e-13
        c1.close()
    cb.close()


PEN_CASES = {
    # rows, columns (spanning the chunks; "bus27": chunk 0 has none), mu
    "zz9": ([6, 7, 8], [0, 1, 3, 4, 8], 0.2),            # chunks of 5: columns 0-4 | 5-8
    "bus27": ([20, 25, 26], [9, 13, 26], 0.4),           # chunks of 7: only chunks 1 and 3 carry it
    "synth16": ([3, 15], [0, 7, 8, 15], 0.9),            # first and last column of both chunks (all 16 columns of a
                                                         # unitary would make L constant: rows of x_k have unit norm)
    "synth32": ([1, 31], [5, 6, 11], 0.7),               # general path, chunks of 6
}


@pytest.mark.parametrize("name", ["zz9", "bus27", "synth16", "synth32"])
@pytest.mark.parametrize("order", [0, 3])
def test_full_propagator_with_running_penalty(name, order):
    """setup_state_penalty on a full propagator (the leakage penalty of examples/ipopt_callbacks_exp.jl:18 with m = d): the
    penalty is separable over columns, every chunk carries its own columns of L and dL_dx."""
    cfg = full_propagator_case(name)
    pen = PEN_CASES[name]
    Jo, go, co = o.evaluate(cfg, order=order, penalty=pen)
    J0, g0, _ = o.evaluate(cfg, order=order)
    assert abs(Jo - J0) > 1e-3 and np.abs(go - g0).max() > 1e-6 * np.abs(go).max()
    L, dL = q.setup_state_penalty(*pen)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost[1], dUkdp_order=order, penalty=(L, dL))
    assert abs(J - Jo) <= TOL_J * max(1.0, abs(Jo))
    assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()
    # reference-style split calls, host-closure cost: cache.J is then the running sum of the penalty alone
    c2 = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    Jf, dJf = o.setup_infidelity(cfg["T"], cfg["n"])
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], c2, penalty=(L, dL))
    assert abs(c2.J - sum(L(xk) for xk in co["x"])) <= TOL_J
    g2 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], c2, dUkdp_order=order, dL_dx=dL)
    assert np.abs(g2 - go).max() <= TOL_G * np.abs(go).max()
    assert np.abs(c2.lam - co["lam"]).max() < 1e-10
    cache.close()
    c2.close()
