"""SURVEY.md 8f "next" rows, each to the same parity bar as the hot path:
N1 spline parameterisation + optimiser callbacks (examples/ipopt_callbacks_exp.jl), N3 the ODE path as a second,
independent oracle (src/gradient_computations.jl:99-169), N4 pulse files and compress_states (src/utils.jl:96-109)."""
import os

import numpy as np
import pytest

import qoc_oracle as o
import qoc_ode_oracle as ode
import qoc_b200 as q
from qoc_b200 import callbacks, pulse_io, configs


# ---------------------------------------------------------------- N3: ODE path vs exp path (CPU) --------------------
def test_tsit5_tableau():
    chk = ode.tableau_check()
    assert chk["row_sum_err"] < 1e-14
    assert max(chk["order_conditions"]) < 1e-14


def test_ode_path_agrees_with_exp_path_zz():
    """test/test_gradient_computation.jl:41-54 runs both paths on the same pulse and displays both gradients; here the
    agreement is asserted: fixed-step Tsit5 with h = 0.1 dt is 5th order -> 1e-6 relative on states and gradient."""
    cfg = o.config_zz()
    dt = cfg["dt"]
    A0, A = cfg["A0"] / dt, [a / dt for a in cfg["A"]]          # per unit time for the ODE path
    J, dJ = o.cost_closures(cfg)
    xs = ode.propagate_pwc(A0, A, cfg["x0"], cfg["u"], dt)
    cache = o.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"])
    x_exp = cache["x"]
    assert np.abs(xs[-1] - x_exp[-1]).max() < 1e-7
    g_ode, lams = ode.compute_pwc_gradient(A0, A, dJ, xs, cfg["u"], dt, dUkdp_order=3)
    g_exp = o.grape_sensitivity(cfg["A0"], cfg["A"], dJ, cfg["u"], cfg["x0"], cache, dUkdp_order=3)
    assert np.abs(g_ode - g_exp).max() <= 1e-6 * np.abs(g_exp).max()


# ---------------------------------------------------------------- N4: pulse files, compress_states (CPU) ------------
def test_pulse_file_round_trip(tmp_path, golden_dir):
    path = os.path.join(golden_dir, "cavity_qubit_pulse.txt")
    u = pulse_io.read_pulse(path)
    assert u.shape == (2, 550)
    assert np.array_equal(u, configs.load_pulse_csv(path)[:, :550])
    assert abs(np.abs(u[0] + 1j * u[1]).max() - 0.05797) < 1e-4      # max|u| = 0.05797
    out = tmp_path / "pulse.txt"
    pulse_io.write_pulse(out, u)
    assert np.array_equal(pulse_io.read_pulse(out), u)               # bit-exact with %.17g


def test_compress_states_round_trip():
    rng = np.random.default_rng(3)
    v = ((np.array([0, 1, 2]), np.array([0, 1])), (np.array([3, 4, 5]), np.array([2, 3, 4])))
    x = np.zeros((6, 5), dtype=np.complex128)
    x[np.ix_(*v[0])] = rng.standard_normal((3, 2)) + 1j * rng.standard_normal((3, 2))
    x[np.ix_(*v[1])] = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    xc = pulse_io.compress_states(x, v)
    assert xc.shape == (6, 3)
    assert np.array_equal(pulse_io.decompress_states(xc, v), x)


# ---------------------------------------------------------------- N1: chain rule and constraints (CPU) --------------
def test_constraints_and_jacobian():
    c = np.random.default_rng(0).standard_normal(20)
    g = callbacks.constraints(c, 10, 2)
    cm = c.reshape(2, 10).T
    assert np.allclose(g, [np.linalg.norm(cm), np.linalg.norm(np.diff(cm, axis=0))])
    Jac = callbacks.constraints_jacobian(c, 10, 2)
    h = 1e-6
    fd = np.array([(callbacks.constraints(c + h * e, 10, 2) - callbacks.constraints(c - h * e, 10, 2)) / (2 * h)
                   for e in np.eye(20)]).T
    assert np.abs(Jac - fd).max() < 1e-8


def test_spline_chain_rule_against_finite_differences_of_the_oracle():
    """dJdc = B' * transpose(dJdu) (examples/ipopt_callbacks_exp.jl:28) is the gradient of c -> J(transpose(B*c))."""
    cfg = o.config_zz()
    B = configs.bspline_matrix()
    c = np.concatenate([0.01 * np.ones(10), np.zeros(10)]) + 0.02 * np.random.default_rng(1).standard_normal(20)

    def J_of(cv):
        u = (B @ cv.reshape(2, 10).T).T
        return o.evaluate(cfg, order=0, u=u)[0]

    u = (B @ c.reshape(2, 10).T).T
    _, dJdu, _ = o.evaluate(cfg, order=0, u=u)
    dJdc = (B.T @ dJdu.T).T.reshape(-1)
    h = 1e-6
    for i in (0, 4, 9, 13, 19):
        e = np.zeros(20); e[i] = h
        fd = (J_of(c + e) - J_of(c - e)) / (2 * h)
        assert abs(fd - dJdc[i]) <= 1e-7 * max(1.0, np.abs(dJdc).max())


# ---------------------------------------------------------------- N1 on the GPU ------------------------------------
@pytest.mark.gpu
def test_ipopt_callbacks_against_oracle():
    cfg = o.config_zz()
    B = configs.bspline_matrix()
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    f, g, f_grad, g_jac, nu, ng, nx, nc, cache = callbacks.setup_ipopt_callbacks(
        cfg["A0"], cfg["A"][0], cfg["A"][1], cfg["x0"], np.zeros((2, 100)), cost, None, B, dUkdp_order=3)
    assert (nu, ng, nc) == (2, 2, 20)
    c = np.concatenate([0.01 * np.ones(10), np.zeros(10)]) + 0.02 * np.random.default_rng(1).standard_normal(20)
    u = (B @ c.reshape(2, 10).T).T
    Jo, dJdu_o, _ = o.evaluate(cfg, order=3, u=u)
    dJdc_o = (B.T @ dJdu_o.T).T.reshape(-1)
    J = f(c)
    out = np.zeros(nc)
    f_grad(c, out)
    assert abs(J - Jo) <= 1e-10 * max(1.0, abs(Jo))
    assert np.abs(out - dJdc_o).max() <= 1e-8 * np.abs(dJdc_o).max()
    # f_grad at a point f has not seen re-runs f first (examples/ipopt_callbacks_exp.jl:22-25)
    c2 = c + 0.01
    out2 = f_grad(c2)
    u2 = (B @ c2.reshape(2, 10).T).T
    d2 = (B.T @ o.evaluate(cfg, order=3, u=u2)[1].T).T.reshape(-1)
    assert np.abs(out2 - d2).max() <= 1e-8 * np.abs(d2).max()
    # fused device path: both skinny products on the device
    Jf, gf = f.f_and_grad(c)
    assert abs(Jf - Jo) <= 1e-10 * max(1.0, abs(Jo))
    assert np.abs(gf - dJdc_o).max() <= 1e-8 * np.abs(dJdc_o).max()
    rows, cols, vals = np.zeros(ng * nc), np.zeros(ng * nc), np.zeros(ng * nc)
    g_jac(c, "Structure", rows, cols, None)
    assert rows[0] == 1 and rows[-1] == 2 and cols[nc - 1] == nc
    g_jac(c, "Values", None, None, vals)
    assert np.allclose(vals, callbacks.constraints_jacobian(c, 10, 2).reshape(-1))


@pytest.mark.gpu
def test_optimiser_loop_reduces_infidelity():
    """A complete optimiser loop around the path (L-BFGS-B with the reference's box bounds): the zz gate infidelity
    drops by orders of magnitude from the reference's starting point (examples/zz_coupling_ipopt_exp.jl:62)."""
    cfg = o.config_zz()
    B = configs.bspline_matrix()
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    f, *_ = callbacks.setup_ipopt_callbacks(cfg["A0"], cfg["A"][0], cfg["A"][1], cfg["x0"], np.zeros((2, 100)), cost, None, B,
                                            dUkdp_order=0)
    c0 = np.concatenate([0.01 * np.ones(10), np.zeros(10)])
    J0, _ = f.f_and_grad(c0)
    bound = 2 * np.pi * 0.060
    res = callbacks.minimize_lbfgs(f.f_and_grad, c0, bounds=[(-bound, bound)] * 20, maxiter=60)
    assert res.fun < 0.05 * J0
    # the optimum is a genuine one for the oracle too
    u = (B @ res.x.reshape(2, 10).T).T
    assert abs(o.evaluate(cfg, order=0, u=u)[0] - res.fun) <= 1e-10
