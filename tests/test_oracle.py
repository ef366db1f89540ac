"""Pins the oracle (oracle/qoc_oracle.py and oracle/qoc_ref.c) against every known answer / fixture the reference
holds for the hot path, before anything is allowed to trust it.  CPU only."""
import os

import numpy as np
import pytest
import scipy.linalg as sla

import qoc_oracle as o
import qoc_ref


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


# ---- the reference's two example known answers (examples/cavity_qubit.jl:80-81, two_qubit_tunable_bus.jl:66-67)
def test_cavity_known_answer():
    H0, Tc, x0, theta = o.model_cavity_qubit(12)
    cfg = o.config_cavity(12, Nt=550)
    cache = o.propagate(cfg["A0"], cfg["A"], cfg["u"], x0.astype(complex))
    tgt = np.kron([1, 0], np.exp(1j * theta))
    tgt /= np.linalg.norm(tgt)
    ov = abs(np.vdot(tgt, cache["x"][-1][:, 0]))
    assert abs(ov - 0.999979) < 1e-6            # "Should be about 0.999979"
    assert abs(ov - 0.9999786609318578) < 1e-13  # SURVEY F5 value


def test_bus_known_answer(golden_dir):
    cfg = o.config_bus(Nt=2000)  # PWC midpoint rule converges O(dt^2) to the ODE value 0.937218
    c = o.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"])
    pop = abs(np.vdot(cfg["T"], c["x"][-1])) ** 2
    assert abs(pop - 0.937218) < 3e-3
    ka = load(golden_dir, "known_answers")
    assert abs(float(ka["bus_population"]) - 0.937218) < 1e-4     # Nt = 1e4: 0.9372909
    assert abs(float(ka["bus_population"]) - 0.9372908966534538) < 1e-11
    # model pins
    H0, Hc, qb = o.model_two_qubit_tunable_bus()
    assert abs(np.linalg.norm(H0, 1) - 106.017) < 1e-3 and abs(np.linalg.norm(Hc, 1) - 94.248) < 1e-3
    assert qb("110") == 12 and qb("200") == 18


# ---- test/test_fidelities.jl known answers
@pytest.mark.parametrize("m,val,basic", [
    ([1, 1j, 1j, 1], 2.8284271, 2.0),
    ([1, 0.1j, 0.1j, 1], 2.0099751, 0.2),
    (list(np.exp(1j * np.array([1., 2, 3, 4]))), 4.0, 4.0),
])
def test_fidelity_kats(m, val, basic):
    assert abs(o.abs_sum_phase_calibrated(m) - val) < 1e-6
    assert abs(o.abs_sum_phase_calibrated(m, "basic") - basic) < 1e-6
    assert abs(o.abs_sum_phase_calibrated(m, "grid") - val) < 1e-3
    J = lambda th: abs(m[0] + m[1] * np.exp(1j * th[0]) + m[2] * np.exp(1j * th[1]) + m[3] * np.exp(1j * (th[0] + th[1])))
    assert abs(J(o.optimal_calibration(m)[1]) - o.abs_sum_phase_calibrated(m)) < 1e-8


def test_fidelity_kats_more():
    m = np.exp(1j * np.array([1, 2, -2.5, -1.7]))
    assert abs(o.abs_sum_phase_calibrated(m) - 3.995001) < 1e-6
    th = o.optimal_calibration(m)[1]
    assert np.allclose(th, [5.383258515112539, 3.6000220820575084], atol=1e-4)
    m = np.exp(1j * np.array([2.5, 2.5, 1.5, -2.5]))
    assert abs(o.abs_sum_phase_calibrated(m) - 3.365883939061934) < 1e-8
    m = [0.65 - 0.75j, -0.4 + 0.8j, -0.4 + 0.1j, 0.7]
    assert abs(o.abs_sum_phase_calibrated(m) - 2.9787244710195484) < 1e-8
    assert abs(o.optimal_calibration(m, 1e-15)[0] - 2.9787244710195484) < 1e-12
    # infidelity entry point (src/fidelities.jl:1-7)
    U = np.diag(np.exp(1j * np.array([0.3, 1.1, -0.4, 0.4])))
    assert abs(o.infidelity(np.eye(4), U, "optimal")) < 1e-9
    with pytest.raises(ValueError):
        o.infidelity(np.eye(3), np.eye(3))


def test_fidelity_gradient_vs_fd():
    # test/test_fidelities.jl:130-148 (FiniteDifferences, rtol 1e-6) on F^2
    rng = np.random.default_rng(100)
    for _ in range(50):
        m = rng.random(4) * np.exp(2j * np.pi * rng.random(4))
        F, th = o.optimal_calibration(m, 1e-12)
        ga = o.abs_sum_phase_calibrated_grad(m, th[0])
        h = 1e-6
        for i in range(4):
            for dz, part in ((h, "re"), (1j * h, "im")):
                mp, mm = m.copy(), m.copy()
                mp[i] += dz
                mm[i] -= dz
                fd = (o.abs_sum_phase_calibrated(mp) ** 2 - o.abs_sum_phase_calibrated(mm) ** 2) / (2 * h)
                an = ga[i].real if part == "re" else ga[i].imag
                assert abs(fd - an) < 2e-5 * max(1.0, abs(an))


# ---- test/test_expm_jacobian.jl: truncated Taylor Jacobian vs finite differences of exp
def test_expm_jacobian_thresholds():
    rng = np.random.default_rng(0)
    A0, A1, A2 = (0.05 * rng.standard_normal((3, 3)) for _ in range(3))
    u = np.array([1.0, 2.0])

    def fd(dt):
        out = []
        for j in range(2):
            h = 1e-6
            up, um = u.copy(), u.copy()
            up[j] += h
            um[j] -= h
            out.append((sla.expm(dt * (A0 + up[0] * A1 + up[1] * A2)) - sla.expm(dt * (A0 + um[0] * A1 + um[1] * A2))) / (2 * h))
        return out

    for dt, t3, t4 in ((1.0, 4e-4, 3e-5), (0.25, 2e-6, 3e-8)):
        ref = fd(dt)
        e = {od: np.sqrt(sum(np.linalg.norm(a - b) ** 2 for a, b in zip(o.expm_jacobian(A0, [A1, A2], u, od, dt), ref)))
             for od in (1, 2, 3, 4)}
        assert e[3] < t3 and e[4] < t4 and e[1] > e[2] > e[3] > e[4]


# ---- test/test_penalty_fcns.jl: analytic cost gradients == Zygote-convention gradient
def _zygote_fd(J, x, h=1e-6):
    g = np.zeros_like(x)
    for idx in np.ndindex(*x.shape):
        for dz in (h, 1j * h):
            xp, xm = x.copy(), x.copy()
            xp[idx] += dz
            xm[idx] -= dz
            v = (J(xp) - J(xm)) / (2 * h)
            g[idx] += v if dz == h else 1j * v
    return g


def test_cost_gradients_zygote_convention():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((9, 4)) + 1j * rng.standard_normal((9, 4))
    Q, _ = np.linalg.qr(rng.standard_normal((9, 8)) + 1j * rng.standard_normal((9, 8)))
    T = Q[:, :4]
    for J, dJ in (o.setup_infidelity(T), o.setup_infidelity(T, 4), o.setup_infidelity_abs_trace(T),
                  o.setup_infidelity_zcalibrated(T), o.setup_state_penalty([6, 7, 8], [0, 1, 2, 3], 0.22)):
        assert np.allclose(dJ(x), _zygote_fd(J, x), atol=2e-6)
    L, dL = o.setup_state_penalty([6, 7, 8], [0, 1, 2, 3], 0.22)
    x0 = np.arange(1.0, 82).reshape(9, 9, order="F")  # test_penalty_fcns.jl:11
    assert L(x0) == pytest.approx(0.22 * np.linalg.norm(x0[np.ix_([6, 7, 8], [0, 1, 2, 3])]) ** 2, rel=1e-15)
    with pytest.raises(ValueError):
        o.setup_infidelity_zcalibrated(T[:, :3])


# ---- expm restatement vs scipy; Frechet restatement vs scipy and vs the block-triangular identity
@pytest.mark.parametrize("d,scale", [(9, 0.2), (27, 6.0), (27, 12.0), (16, 1.5), (5, 0.01), (12, 0.9), (3, 0.1)])
def test_expm_and_frechet(d, scale):
    rng = np.random.default_rng(d)
    G = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
    X = -1j * (G + G.conj().T)
    X *= scale / np.linalg.norm(X, 1)
    E = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
    assert np.abs(o.expm_higham2005(X) - sla.expm(X)).max() < 1e-13
    R, L = o.expm_frechet_sps(X, E)
    R2, L2 = sla.expm_frechet(X, E)
    R3, L3 = o.expm_frechet_blocktri(X, E)
    assert np.abs(R - R2).max() < 1e-13
    assert np.abs(L - L2).max() < 1e-12 * np.abs(L2).max()
    assert np.abs(L - L3).max() < 1e-12 * np.abs(L3).max()


def test_expm_nonnormal_input():
    # the reference's own Jacobian test uses real non-normal matrices
    rng = np.random.default_rng(3)
    X = rng.standard_normal((6, 6))
    assert np.abs(o.expm_higham2005(X) - sla.expm(X)).max() < 1e-12


# ---- gradient vs finite differences of propagate (test/test_gradient_computation.jl:97-98)
def test_gradient_vs_finite_differences():
    cfg = o.config_zz()
    J0, g0, _ = o.evaluate(cfg, order=0)
    Jf, _ = o.cost_closures(cfg)
    for (j, k) in ((0, 0), (1, 37), (0, 99), (1, 50)):
        h = 1e-6
        up, um = cfg["u"].copy(), cfg["u"].copy()
        up[j, k] += h
        um[j, k] -= h
        fd = (Jf(o.propagate(cfg["A0"], cfg["A"], up, cfg["x0"])["x"][-1]) -
              Jf(o.propagate(cfg["A0"], cfg["A"], um, cfg["x0"])["x"][-1])) / (2 * h)
        assert abs(fd - g0[j, k]) < 1e-8 * max(1, abs(g0[j, k]) / 1e-2)
    assert abs(J0 - 0.9862874005098565) < 1e-13
    assert g0[0, 0] == pytest.approx(-0.0232113407, abs=1e-9)
    # Taylor orders converge to the exact derivative
    errs = [np.abs(o.evaluate(cfg, order=od)[1] - g0).max() / np.abs(g0).max() for od in (1, 2, 3, 4)]
    assert errs[0] == pytest.approx(3.078e-3, rel=1e-2) and errs[2] == pytest.approx(2.414e-6, rel=1e-2)
    assert errs[0] > errs[1] > errs[2] > errs[3] and errs[3] < 2e-7


def test_stale_cache_and_dimension_errors():
    cfg = o.config_zz()
    cache = o.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"])
    with pytest.raises(RuntimeError, match="Cache data from other control signal u"):
        o.grape_sensitivity(cfg["A0"], cfg["A"], lambda x: x, cfg["u"] + 1e-3, cfg["x0"], cache)
    with pytest.raises(ValueError, match="incompatiable"):
        o.setup_grape_cache(cfg["A0"], np.zeros((8, 4), complex), (2, 100))


# ---- committed golden fixtures reproduce (guards against silent oracle edits)
@pytest.mark.parametrize("name,cfgf,order,pen", [
    ("zz_order3", lambda: o.config_zz(), 3, None),
    ("zz_order0", lambda: o.config_zz(), 0, None),
    ("zz_penalty_order4", lambda: o.config_zz(), 4, ([6, 7, 8], [0, 1, 2, 3], 0.22)),
    ("cavity12_nt100_order3", lambda: o.config_cavity(12, Nt=100), 3, None),
    ("bus_nt500_order0", lambda: o.config_bus(Nt=500, tgate=17.5), 0, None),
    ("synth16_nt64_order0", lambda: o.config_synthetic(16, 64), 0, None),
])
def test_golden_reproduce(golden_dir, name, cfgf, order, pen):
    gd = load(golden_dir, name)
    cfg = cfgf()
    assert np.array_equal(cfg["u"], gd["u"])
    J, g, cache = o.evaluate(cfg, order=order, penalty=pen)
    assert abs(J - float(gd["J"])) < 1e-13
    assert np.abs(g - gd["dJdu"]).max() < 1e-12 * max(1.0, np.abs(gd["dJdu"]).max())


# ---- the C restatement (timed CPU baseline) agrees with the numpy restatement
@pytest.mark.parametrize("cfgf,order,pen", [
    (lambda: o.config_zz(), 3, None), (lambda: o.config_zz(), 0, None), (lambda: o.config_zz(), 1, None),
    (lambda: o.config_zz(), 4, ([6, 7, 8], [0, 1, 2, 3], 0.22)),
    (lambda: o.config_cavity(12, Nt=60), 3, None), (lambda: o.config_cavity(12, Nt=60), 0, None),
    (lambda: o.config_bus(Nt=120, tgate=4.2), 0, None), (lambda: o.config_synthetic(16, 32), 2, None),
])
def test_c_port_matches_numpy_oracle(cfgf, order, pen):
    cfg = cfgf()
    J, g, cache = o.evaluate(cfg, order=order, penalty=pen)
    for nth in (1, 2):
        out = qoc_ref.ref_eval(cfg, order=order, nthreads=nth, penalty=pen, want_cache=True)
        assert abs(out["J"] - J) < 1e-13
        assert np.abs(out["dJdu"] - g).max() < 1e-11 * max(np.abs(g).max(), 1e-3)
        assert np.abs(out["Uk"] - cache["Uk"]).max() < 1e-13
        assert np.abs(out["x"] - cache["x"]).max() < 1e-12
        assert np.abs(out["lam"] - cache["lam"]).max() < 1e-12


def test_f_alg_worked_values():
    # SURVEY.md 8(d) worked values (Frechet mode)
    assert o.f_alg(9, 4, 2, 7, 0, 0) == pytest.approx(158.7e3, rel=2e-3)
    assert o.f_alg(27, 1, 1, 13, 1, 0) == pytest.approx(3.85e6, rel=3e-3)
    assert o.f_alg(24, 2, 2, 7, 0, 0) == pytest.approx(2.84e6, rel=3e-3)
    assert o.f_alg(256, 4, 2, 13, 0, 0) == pytest.approx(4.74e9, rel=3e-3)
    assert o.f_alg(9, 4, 2, 7, 0, 3) == pytest.approx(100e3, rel=2e-2)


def test_expm_restatements_against_binary128_ground_truth():
    """SURVEY 8c oracle stack: the reference's expm lives in an un-vendored dependency, so the oracle's own Higham-2005 and
    Al-Mohy-Higham (Frechet) restatements are pinned against an independent binary128 computation (oracle/qoc_quad.c:
    Taylor series + squaring in __float128, central difference with step 2^-40 for the derivative) on slices of every
    named model, on a non-normal generator, and over the norm regimes that select every Pade degree."""
    import qoc_quad
    cases = []
    for cfg in (o.config_zz(), o.config_bus(Nt=40, tgate=1.4), o.config_cavity(12, Nt=30)):
        for k in (0, cfg["u"].shape[1] // 2, cfg["u"].shape[1] - 1):
            cases.append((o.generator(cfg["A0"], cfg["A"], cfg["u"][:, k]), cfg["A"][0]))
    rng = np.random.default_rng(11)
    for d, scale in ((6, 0.05), (6, 0.5), (8, 3.0), (5, 12.0), (7, 45.0)):
        cases.append((scale * rng.standard_normal((d, d)) / d + 0j, rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))))
    for X, E in cases:
        U, L = qoc_quad.expm_quad(X, E)
        sU, sL = max(1.0, np.abs(U).max()), max(1.0, np.abs(L).max())
        assert np.abs(o.expm_higham2005(X) - U).max() < 5e-14 * sU
        Us, Ls = o.expm_frechet_sps(X, E)
        assert np.abs(Us - U).max() < 5e-14 * sU
        assert np.abs(Ls - L).max() < 2e-13 * sL
        assert np.abs(o.expm_frechet_blocktri(X, E)[1] - L).max() < 2e-13 * sL
