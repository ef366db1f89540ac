"""Parity tests proper (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against
(1) the committed golden fixtures, (2) the oracle on the same seeded inputs at sizes it finishes in seconds,
(3) size-independent properties at BASELINE.json's full sizes.

Tolerances (north star):
  |dJ| <= 1e-10 * max(1, |J|)                       (relative on the fidelity, well-posed near J -> 0)
  |dg_jk| <= 1e-8 * max_jk |g_jk|  per component     (same-order Taylor mode vs the reference formula;
                                                      exact-Frechet mode vs the oracle's exact Frechet derivative)
U_k, states and costates: 1e-12 absolute (unit-norm quantities).
"""
import os

import numpy as np
import pytest

import qoc_oracle as o
import qoc_ref
import qoc_b200 as q
from qoc_b200 import _lib

pytestmark = pytest.mark.gpu

TOL_J = 1e-10
TOL_G = 1e-8


def cost_of(cfg):
    if cfg["cost"] == o.COST_INFIDELITY:
        return q.setup_infidelity(cfg["T"], cfg["n"])
    return q.setup_infidelity_abs_trace(cfg["T"])


def gpu_eval(cfg, order, u=None, batch=1):
    u = cfg["u"] if u is None else u
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], u.shape[-2:], batch=batch, dUkdp_order=order)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], cost_of(cfg)[1], dUkdp_order=order)
    return J, g, cache


def assert_parity(J, g, Jref, gref):
    assert abs(J - Jref) <= TOL_J * max(1.0, abs(Jref))
    assert np.abs(g - gref).max() <= TOL_G * np.abs(gref).max()


def test_library_is_the_cuda_one():
    lib = _lib.load()
    assert lib.qoc_version() >= 100
    import torch
    assert torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10


# ---- (1) golden fixtures ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,cfgf,order", [
    ("zz_order1", lambda: o.config_zz(), 1), ("zz_order2", lambda: o.config_zz(), 2),
    ("zz_order3", lambda: o.config_zz(), 3), ("zz_order4", lambda: o.config_zz(), 4),
    ("zz_order0", lambda: o.config_zz(), 0),
    ("cavity12_nt100_order3", lambda: o.config_cavity(12, Nt=100), 3),
    ("cavity12_nt550_order0", lambda: o.config_cavity(12, Nt=550), 0),
    ("bus_nt500_order0", lambda: o.config_bus(Nt=500, tgate=17.5), 0),
    ("synth16_nt64_order0", lambda: o.config_synthetic(16, 64), 0),
])
def test_golden(golden_dir, name, cfgf, order):
    gd = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = cfgf()
    assert np.array_equal(cfg["u"], gd["u"])
    J, g, cache = gpu_eval(cfg, order)
    assert_parity(J, g, float(gd["J"]), gd["dJdu"])
    assert np.abs(cache.x[-1] - gd["x_final"]).max() < 1e-12
    if "Uk" in gd:
        assert np.abs(cache.Uk_vec - gd["Uk"]).max() < 1e-12
        assert np.abs(cache.x - gd["x"]).max() < 1e-12
        assert np.abs(cache.lam - gd["lam"]).max() < 1e-12


def test_known_answers_full_size(golden_dir):
    """The reference's two example known answers, computed by the CUDA path at full size."""
    ka = np.load(os.path.join(golden_dir, "known_answers.npz"))
    # cavity: examples/cavity_qubit.jl:80-81  "Should be about 0.999979"
    H0, Tc, x0, theta = o.model_cavity_qubit(12)
    cfg = o.config_cavity(12, Nt=550)
    cache = q.propagate(cfg["A0"], cfg["A"], cfg["u"], x0.astype(complex).reshape(-1, 1))
    tgt = np.kron([1, 0], np.exp(1j * theta))
    tgt /= np.linalg.norm(tgt)
    ov = abs(np.vdot(tgt, cache.x_final[:, 0]))
    assert abs(ov - 0.999979) < 1e-6 and abs(ov - float(ka["cavity_overlap"])) < 1e-12
    # bus: examples/two_qubit_tunable_bus.jl:66-67  "Should be something like 0.937218", Nt = 1e4 PWC
    b = o.config_bus(Nt=10000)
    cb = q.propagate(b["A0"], b["A"], b["u"], b["x0"])
    pop = abs(np.vdot(b["T"], cb.x_final)) ** 2
    assert abs(pop - 0.937218) < 1e-4 and abs(pop - float(ka["bus_population"])) < 1e-10
    assert np.abs(cb.x_final - ka["bus_x_final"]).max() < 1e-11


# ---- (2) oracle on the same seeded inputs ----------------------------------------------------------------------------
@pytest.mark.parametrize("d,nt,nc,m,order", [
    (3, 17, 1, 1, 0), (4, 9, 2, 4, 3), (8, 33, 2, 8, 0), (9, 100, 2, 4, 4), (12, 40, 3, 2, 0), (13, 21, 1, 3, 2),
    (16, 50, 2, 4, 0), (17, 30, 2, 1, 3), (20, 25, 1, 5, 0), (21, 19, 2, 2, 1), (24, 60, 2, 2, 0), (25, 14, 1, 4, 0), (23, 14, 2, 4, 0),
    (27, 80, 1, 1, 0), (28, 31, 1, 7, 3), (24, 31, 2, 7, 3), (27, 1, 1, 1, 0), (9, 2, 2, 4, 0), (25, 14, 2, 4, 0),
])
def test_random_shapes_vs_oracle(d, nt, nc, m, order):
    """Ragged sizes across every shape class, incl. d not a multiple of 4/8, Nt = 1, m up to 8."""
    cfg = o.config_synthetic(d, nt, nc=nc, m=m, seed=1000 + d * 7 + nt)
    Jo, go, co = o.evaluate(cfg, order=order)
    J, g, cache = gpu_eval(cfg, order)
    assert_parity(J, g, Jo, go)
    assert np.abs(cache.Uk_vec - co["Uk"]).max() < 1e-12
    assert np.abs(cache.x - co["x"]).max() < 1e-12
    assert np.abs(cache.lam - co["lam"]).max() < 1e-12


def test_unsupported_sizes_fail_loudly():
    """Sizes no kernel covers must raise, never fall back to anything on the CPU."""
    for d, nc, m in ((600, 1, 1), (16, 9, 2), (16, 1, 65)):
        cfg = o.config_synthetic(d, 2, nc=nc, m=m, seed=1) if (d < 100 and m <= 64) else None
        if cfg is None:
            with pytest.raises(q.QOCError) as ei:
                c = q.setup_grape_cache(np.zeros((d, d), complex), np.zeros((d, m), complex), (nc, 2))
                c._ensure(np.zeros((d, d), complex), [np.zeros((d, d), complex)] * nc, np.zeros((d, m), complex))
        else:
            with pytest.raises(q.QOCError) as ei:
                gpu_eval(cfg, 0)
        assert ei.value.status == _lib.ERR_UNSUPPORTED


# ---- general path (d > 28, or a working set that does not fit shared memory): csrc/qoc_gpath.cuh --------------------
@pytest.mark.parametrize("d,nt,nc,m,order", [
    (28, 9, 2, 3, 0),      # d <= 28 but 17 + 3 nc - 2 matrices do not fit on chip -> general path
    (29, 11, 1, 2, 0), (32, 20, 2, 4, 0), (33, 7, 2, 1, 3), (40, 16, 2, 2, 4), (50, 6, 3, 8, 0), (64, 12, 2, 4, 0),
    (80, 8, 2, 2, 0), (100, 4, 1, 1, 2), (128, 5, 1, 2, 0), (130, 3, 2, 2, 1), (256, 2, 1, 4, 0),
])
def test_general_path_vs_oracle(d, nt, nc, m, order):
    cfg = o.config_synthetic(d, nt, nc=nc, m=m, seed=2000 + d + nt)
    Jo, go, co = o.evaluate(cfg, order=order)
    J, g, cache = gpu_eval(cfg, order)
    assert_parity(J, g, Jo, go)
    assert np.abs(cache.Uk_vec - co["Uk"]).max() < 1e-12
    assert np.abs(cache.x - co["x"]).max() < 1e-12
    assert np.abs(cache.lam - co["lam"]).max() < 1e-12


def test_general_path_cavity_truncations():
    """BASELINE.json configs[2]: cavity-qubit with cavity truncation 20 and 40 levels (d = 40, 80), CSV pulse."""
    for N_cav, nt in ((20, 120), (40, 40)):
        cfg = o.config_cavity(N_cav, Nt=nt)
        for order in (3, 0):
            Jo, go, _ = o.evaluate(cfg, order=order)
            J, g, _ = gpu_eval(cfg, order)
            assert_parity(J, g, Jo, go)


def test_general_path_forced_on_small_sizes(monkeypatch):
    """The same configs the shared-memory path serves, pushed through the general path (QOC_FORCE_GPATH=1):
    penalty, host-closure costate, Taylor orders, several squarings, batches."""
    monkeypatch.setenv("QOC_FORCE_GPATH", "1")
    cfg = o.config_zz()
    for order in (0, 3, 4):
        Jo, go, co = o.evaluate(cfg, order=order)
        J, g, cache = gpu_eval(cfg, order)
        assert_parity(J, g, Jo, go)
        assert cache.launch_count() > 3   # really the multi-launch path
    pen = ([6, 7, 8], [0, 1, 2, 3], 0.22)
    Jo, go, co = o.evaluate(cfg, order=4, penalty=pen)
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 100), dUkdp_order=4)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1],
                      dUkdp_order=4, penalty=q.setup_state_penalty(*pen))
    assert_parity(J, g, Jo, go)
    assert np.abs(cache.lam - co["lam"]).max() < 1e-12
    # host closure (lambda_final) path
    c2 = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 100))
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], c2)
    Jf, dJf = o.setup_infidelity(cfg["T"], cfg["n"])
    g2 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], c2, dUkdp_order=3)
    _, go3, _ = o.evaluate(cfg, order=3)
    assert np.abs(g2 - go3).max() <= TOL_G * np.abs(go3).max()
    # several squarings
    s = o.config_synthetic(12, 24, nc=2, m=3, seed=5)
    s["A0"] = s["A0"] * 11
    s["A"] = [a * 11 for a in s["A"]]
    for order in (0, 3):
        Jo, go, co = o.evaluate(s, order=order)
        J, g, cache = gpu_eval(s, order)
        assert abs(J - Jo) <= TOL_J and np.abs(g - go).max() <= TOL_G * np.abs(go).max()
    # batch
    cb = o.config_zz_batch(5)
    Jb, gb, _ = gpu_eval(cb, 3, u=cb["u_batch"], batch=5)
    for b in (0, 4):
        Jo, go, _ = o.evaluate(cb, order=3, u=cb["u_batch"][b])
        assert_parity(Jb[b], gb[b], Jo, go)


@pytest.mark.parametrize("scale", [1e-3, 0.3, 3.0, 11.0, 40.0])
def test_norm_regimes_vs_oracle(scale):
    """||X_k||_1 from tiny to large: 0..4 squarings, both threshold tables (Taylor mode 5.4, Frechet mode 4.74)."""
    cfg = o.config_synthetic(12, 24, nc=2, m=3, seed=5)
    cfg["A0"] = cfg["A0"] * scale
    cfg["A"] = [a * scale for a in cfg["A"]]
    for order in (0, 3):
        Jo, go, co = o.evaluate(cfg, order=order)
        J, g, cache = gpu_eval(cfg, order)
        assert abs(J - Jo) <= TOL_J
        # at large norm the truncated Taylor gradient is huge and ill-conditioned: compare relative to its scale
        assert np.abs(g - go).max() <= TOL_G * max(np.abs(go).max(), 1e-300)
        assert np.abs(cache.Uk_vec - co["Uk"]).max() < 1e-11


def test_nonnormal_generators_need_pivoting():
    """Real non-normal matrices as in the reference's test/test_expm_jacobian.jl, and a rotation generator whose
    Pade denominator has a (near-)zero leading entry: exercises the partial pivoting of the in-kernel inverse."""
    rng = np.random.default_rng(0)
    d = 6
    A0 = (0.7 * rng.standard_normal((d, d))).astype(complex)
    A1 = (0.5 * rng.standard_normal((d, d))).astype(complex)
    cfg = dict(A0=A0, A=[A1], u=rng.uniform(-1, 1, (1, 12)), x0=np.eye(d, 2, dtype=complex),
               T=np.eye(d, 2, dtype=complex), cost=o.COST_INFIDELITY, n=2)
    for order in (0, 3):
        Jo, go, co = o.evaluate(cfg, order=order)
        J, g, cache = gpu_eval(cfg, order)
        assert_parity(J, g, Jo, go)
        assert np.abs(cache.Uk_vec - co["Uk"]).max() < 1e-11 * max(1, np.abs(co["Uk"]).max())
    G = np.array([[0, np.pi], [-np.pi, 0]], dtype=complex)   # exp(-G/2) has zero diagonal
    cfg = dict(A0=G, A=[np.array([[0, 1], [1, 0]], dtype=complex) * 1j], u=np.array([[0.0, 0.3, -0.2]]),
               x0=np.eye(2, dtype=complex), T=np.eye(2, dtype=complex), cost=o.COST_INFIDELITY, n=2)
    Jo, go, co = o.evaluate(cfg, order=0)
    J, g, cache = gpu_eval(cfg, 0)
    assert_parity(J, g, Jo, go)
    assert np.abs(cache.Uk_vec - co["Uk"]).max() < 1e-12


def test_reference_style_calls_and_host_closure(golden_dir):
    """Mirrors test/test_gradient_computation.jl:27-35: propagate, then grape_sensitivity with an arbitrary host
    closure dJfinal_dx (here the oracle's closure) -- the lambda_final path of the C ABI."""
    cfg = o.config_cavity(12, Nt=100)
    gd = np.load(os.path.join(golden_dir, "cavity12_nt100_order3.npz"))
    Jf, dJf = o.setup_infidelity_abs_trace(cfg["T"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 100))
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cache)
    x = cache.x
    assert abs(Jf(x[-1]) - float(gd["J"])) < TOL_J
    g = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], cache, dUkdp_order=3)
    assert np.abs(g - gd["dJdu"]).max() <= TOL_G * np.abs(gd["dJdu"]).max()
    # the same through the device-side built-in cost, and with another order on the same cache
    g2 = q.grape_sensitivity(cfg["A0"], cfg["A"], q.setup_infidelity_abs_trace(cfg["T"])[1], cfg["u"], cfg["x0"], cache, dUkdp_order=3)
    assert np.abs(g2 - g).max() < 1e-13
    g4 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], cache, dUkdp_order=4)
    _, go4, _ = o.evaluate(cfg, order=4)
    assert np.abs(g4 - go4).max() <= TOL_G * np.abs(go4).max()
    # stale cache error, src/gradient_computations.jl:37-39
    with pytest.raises(q.QOCError, match="Cache data from other control signal u"):
        q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"] + 1e-9, cfg["x0"], cache)


def test_gradient_vs_finite_differences_of_gpu_propagate():
    """test/test_gradient_computation.jl:97-98: exact-Frechet gradient vs central differences of J through the
    CUDA propagate itself (validates the exact mode independently of the oracle's Frechet code)."""
    cfg = o.config_zz()
    J, g, cache = gpu_eval(cfg, 0)
    Jf = q.setup_infidelity(cfg["T"], cfg["n"])[0]
    pc = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 100))
    for (j, k) in ((0, 0), (1, 37), (0, 99), (1, 63)):
        h = 1e-6
        up, um = cfg["u"].copy(), cfg["u"].copy()
        up[j, k] += h
        um[j, k] -= h
        Jp = q.propagate(cfg["A0"], cfg["A"], up, cfg["x0"], pc, Jfinal=Jf).J
        Jm = q.propagate(cfg["A0"], cfg["A"], um, cfg["x0"], pc, Jfinal=Jf).J
        assert abs((Jp - Jm) / (2 * h) - g[j, k]) < 2e-9
    # order 4 -> exact convergence
    _, g4, _ = gpu_eval(cfg, 4)
    assert np.abs(g4 - g).max() / np.abs(g).max() == pytest.approx(9.77e-8, rel=0.02)


def test_batch_equals_individual_pulses():
    cfg = o.config_zz_batch(37)
    ub = cfg["u_batch"]
    Jb, gb, _ = gpu_eval(cfg, 3, u=ub, batch=37)
    for b in (0, 5, 36):
        Jo, go, _ = o.evaluate(cfg, order=3, u=ub[b])
        assert_parity(Jb[b], gb[b], Jo, go)
    J1, g1, _ = gpu_eval(cfg, 3, u=ub[11])
    assert abs(J1 - Jb[11]) < 1e-14 and np.abs(g1 - gb[11]).max() < 1e-15


def test_c_port_agrees_on_gpu_box():
    """The timed CPU baseline (oracle/qoc_ref.c) computes the same thing as the CUDA path."""
    cfg = o.config_bus(Nt=300, tgate=10.5)
    J, g, _ = gpu_eval(cfg, 0)
    out = qoc_ref.ref_eval(cfg, order=0)
    assert_parity(J, g, out["J"], out["dJdu"])


# ---- (3) size-independent properties at full size ----------------------------------------------------------------------
def test_bus_full_size_properties():
    """C2 at Nt = 1e4: unitarity of every U_k, norm preservation along the whole trajectory, costate norm
    preservation, the exact-gradient identity sum_k over a doubled-slice refinement, and segment independence
    (the parallel scan must give the same answer as a different segmentation = a shorter problem's prefix)."""
    cfg = o.config_bus(Nt=10000)
    J, g, cache = gpu_eval(cfg, 0)
    U = cache.Uk_vec
    err = np.abs(np.einsum("kji,kjl->kil", U.conj(), U) - np.eye(27)).max()
    assert err < 5e-13
    x = cache.x
    assert np.abs(np.linalg.norm(x[:, :, 0], axis=1) - 1).max() < 1e-11
    lam = cache.lam
    assert np.abs(np.linalg.norm(lam[:, :, 0], axis=1) - np.linalg.norm(lam[-1, :, 0])).max() < 1e-11
    # oracle on a strided sample of slices: U_k and the exact Frechet contraction
    for k in range(0, 10000, 997):
        X = o.generator(cfg["A0"], cfg["A"], cfg["u"][:, k])
        R, L = o.expm_frechet_sps(X, cfg["A"][0])
        assert np.abs(U[k] - R).max() < 1e-12
        assert abs(o.compute_u_sensitivity(x[k], lam[k + 1], L) - g[0, k]) <= TOL_G * np.abs(g).max()
    # prefix property: the first 2500 slices evaluated as their own problem give the same states
    c2 = q.propagate(cfg["A0"], cfg["A"], cfg["u"][:, :2500], cfg["x0"])
    assert np.abs(c2.x_final[:, 0] - x[2500, :, 0]).max() < 1e-11
    # J from the C restatement's forward sweep (serial order) vs the parallel scan
    ref = qoc_ref.ref_eval(cfg, order=0, want_grad=False)
    assert abs(J - ref["J"]) <= TOL_J


def test_linearity_in_terminal_costate():
    """grape_sensitivity is linear in lambda_final: g(a*l1 + b*l2) = a*g(l1) + b*g(l2)."""
    cfg = o.config_synthetic(16, 200, nc=2, m=4)
    rng = np.random.default_rng(2)
    l1 = rng.standard_normal((16, 4)) + 1j * rng.standard_normal((16, 4))
    l2 = rng.standard_normal((16, 4)) + 1j * rng.standard_normal((16, 4))
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 200), dUkdp_order=0)
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cache)
    gs = [q.grape_sensitivity(cfg["A0"], cfg["A"], (lambda l: (lambda x: l))(l), cfg["u"], cfg["x0"], cache, dUkdp_order=0)
          for l in (l1, l2, 0.3 * l1 - 1.7 * l2)]
    assert np.abs(gs[2] - (0.3 * gs[0] - 1.7 * gs[1])).max() < 1e-12 * np.abs(gs[2]).max()


def test_zz_batch_4096_properties():
    """C4 shape: 4096 pulses in one call; spot-check pulses against the oracle and check J in [0, 1]."""
    cfg = o.config_zz_batch(4096)
    ub = cfg["u_batch"]
    Jb, gb, cache = gpu_eval(cfg, 0, u=ub, batch=4096)
    assert np.all(np.isfinite(Jb)) and np.all(Jb > -1e-12) and np.all(Jb < 1 + 1e-12)
    for b in (0, 1234, 4095):
        Jo, go, _ = o.evaluate(cfg, order=0, u=ub[b])
        assert_parity(Jb[b], gb[b], Jo, go)
    assert cache.launch_count() >= 3
    cache.close()


@pytest.mark.parametrize("order", [0, 3])
def test_zz_batch_4096_every_pulse_vs_c_restatement(order):
    """C4 in full: every one of the 4096 pulses (J and all 200 gradient components each) against the C restatement of the
    reference (oracle/qoc_ref.c, which agrees with the numpy oracle to 1e-15: tests/test_oracle.py), both gradient modes."""
    cfg = o.config_zz_batch(4096)
    ub = cfg["u_batch"]
    Jb, gb, cache = gpu_eval(cfg, order, u=ub, batch=4096)
    worst_J = worst_g = 0.0
    for b in range(4096):
        r = qoc_ref.ref_eval(cfg, order=order, nthreads=1, u=ub[b])
        worst_J = max(worst_J, abs(Jb[b] - r["J"]) / max(1.0, abs(r["J"])))
        worst_g = max(worst_g, np.abs(gb[b] - r["dJdu"]).max() / np.abs(r["dJdu"]).max())
    print(f"order {order}: worst |dJ| = {worst_J:.2e}, worst rel |dg| = {worst_g:.2e} over 4096 pulses")
    assert worst_J <= TOL_J and worst_g <= TOL_G
    cache.close()


# ---- running state penalty (src/penalty_fcns.jl:1-11; affine costate recurrence, gradient_computations.jl:47-57) -------
def test_state_penalty_golden_and_variants(golden_dir):
    """test/test_gradient_computation.jl:105-132 set-up (zz_coupling, guard rows "20","21","22", order 4)."""
    gd = np.load(os.path.join(golden_dir, "zz_penalty_order4.npz"))
    cfg = o.config_zz()
    rows, cols, mu = [6, 7, 8], [0, 1, 2, 3], 0.22
    pen = q.setup_state_penalty(rows, cols, mu)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 100), dUkdp_order=4)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost[1], dUkdp_order=4, penalty=pen)
    assert_parity(J, g, float(gd["J"]), gd["dJdu"])
    # reference-style two-step call with host closures for both the cost and the penalty
    Jf, dJf = o.setup_infidelity(cfg["T"], cfg["n"])
    c2 = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 100))
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], c2, penalty=pen)
    x = c2.x
    Jtot = Jf(x[-1]) + sum(pen[0](xk) for xk in x)          # Jfinal(x[end]) + sum(L, x)
    assert abs(Jtot - float(gd["J"])) <= TOL_J * max(1, abs(Jtot))
    assert abs(c2.J - sum(pen[0](xk) for xk in x)) < 1e-12   # device-side running sum of the penalty alone
    g2 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], c2, dUkdp_order=4, dL_dx=pen[1])
    assert np.abs(g2 - gd["dJdu"]).max() <= TOL_G * np.abs(gd["dJdu"]).max()
    # exact-Frechet mode with penalty, longer pulse split over many segments, vs the oracle
    cfg = o.config_synthetic(12, 700, nc=2, m=3, seed=9)
    pen2 = ([1, 5, 11], [0, 2], 0.37)
    Jo, go, co = o.evaluate(cfg, order=0, penalty=pen2)
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 700), dUkdp_order=0)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1],
                      dUkdp_order=0, penalty=q.setup_state_penalty(*pen2))
    assert_parity(J, g, Jo, go)
    assert np.abs(cache.lam - co["lam"]).max() < 1e-11


# ---- time-segment sharding through the C ABI phase API (one GPU, virtual ranks; real ranks: bench.py --shard time) -----
def test_time_sharding_virtual_ranks_on_one_gpu():
    """P contiguous time segments, one all-gather of the rank propagators, redundant boundary algebra.
    Here the P ranks are emulated sequentially on one GPU (a single exchange step, no kernel waits on another)."""
    import torch
    from qoc_b200 import sharding
    cfg = o.config_bus(Nt=1003, tgate=35.105)        # 1003 slices: uneven segments
    P = 4
    Jo, go, co = o.evaluate(cfg, order=0)
    engines, S = [], []
    for r in range(P):
        lo, hi = sharding.time_partition(1003, P, r)
        e = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], hi - lo, cfg["x0"].shape[1], 0, order=0)
        S.append(e.phase1(cfg["u"][:, lo:hi]).clone())
        engines.append(e)
    dev = engines[0].device
    x = torch.as_tensor(cfg["x0"]).to(dev)
    starts = []
    for r in range(P):
        starts.append(x)
        x = S[r] @ x
    assert np.abs(x.cpu().numpy() - co["x"][-1]).max() < 1e-11
    J, lam = sharding._builtin_cost_torch(q.setup_infidelity(cfg["T"], cfg["n"])[1], x)
    assert abs(J - Jo) <= TOL_J
    g = np.zeros_like(go)
    for r in range(P - 1, -1, -1):
        lo, hi = sharding.time_partition(1003, P, r)
        xe = engines[r].forward(starts[r])
        gl, lam_next = engines[r].backward(lam)
        g[:, lo:hi] = gl.cpu().numpy()
        assert np.abs(lam_next.cpu().numpy() - co["lam"][lo]).max() < 1e-11
        lam = S[r].conj().t() @ lam
    assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()
    # and the evaluator class itself with world_size 1
    ev = sharding.TimeShardedEvaluator(sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], 1003, 1, 0, order=0), cfg["x0"],
                                       q.setup_infidelity(cfg["T"], cfg["n"])[1], 1003)
    J1, g1 = ev.evaluate(cfg["u"])
    assert_parity(J1, g1, Jo, go)


def test_time_sharding_phase2_on_device_virtual_ranks():
    """The one-call phase 2 (qoc_shard_phase2_device): boundary algebra over the all-gathered rank propagators on the device,
    two-level local scan from (x_start, lambda_end), sweeps.  P virtual ranks on one GPU, uneven segments."""
    import torch
    from qoc_b200 import sharding
    cfg = o.config_bus(Nt=1003, tgate=35.105)
    P = 3
    Jo, go, co = o.evaluate(cfg, order=0)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])[1]
    engines = []
    dev = torch.device("cuda", 0)
    S_all = torch.empty((P, 27, 27), dtype=torch.complex128, device=dev)
    for r in range(P):
        lo, hi = sharding.time_partition(1003, P, r)
        e = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], hi - lo, cfg["x0"].shape[1], 0, order=0)
        e.set_builtin_cost(cost, cfg["x0"])
        u_dev = torch.from_numpy(np.ascontiguousarray(cfg["u"][:, lo:hi].T)).to(dev)
        S_all[r].copy_(e.phase1_cm(u_dev))
        engines.append(e)
    g = np.zeros_like(go)
    for r in range(P):
        lo, hi = sharding.time_partition(1003, P, r)
        J, gl = engines[r].phase2(S_all, P, r)
        assert abs(float(J.cpu()[0]) - Jo) <= TOL_J
        g[:, lo:hi] = gl.cpu().numpy().T
    assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()


@pytest.mark.parametrize("d,scale", [(7, 0.1), (16, 0.2), (20, 0.05), (24, 0.25), (27, 0.04), (27, 0.2), (28, 0.12)])
def test_low_pade_degrees_every_shape_class(d, scale):
    """||X_k||_1 below the [5/5] / [7/7] switch points (0.2 / 0.783 Frechet, 0.25 / 0.95 Taylor) in every shape class:
    the low-degree Pade forms and their structured Frechet derivatives against the oracle (which selects its own degree)."""
    cfg = o.config_synthetic(d, 12, nc=2, m=2, seed=77 + d)
    cfg["A0"] = cfg["A0"] * scale
    cfg["A"] = [a * scale for a in cfg["A"]]
    for order in (0, 3):
        Jo, go, co = o.evaluate(cfg, order=order)
        J, g, cache = gpu_eval(cfg, order)
        assert_parity(J, g, Jo, go)
        assert np.abs(cache.Uk_vec - co["Uk"]).max() < 1e-13
        if order == 0:   # exact Frechet derivative itself
            dU_o = np.array([[o.expm_frechet_sps(o.generator(cfg["A0"], cfg["A"], cfg["u"][:, k]), a)[1] for a in cfg["A"]]
                             for k in range(cfg["u"].shape[1])])
            assert np.abs(cache.dUkdu - dU_o).max() < 1e-13


def test_time_sharding_general_path_virtual_ranks():
    """Time-segment sharding on the general path (d > 28): rank propagators through the batched GEMM chain, boundary algebra
    on the device, two-level sweeps from (x_start, lambda_end).  P = 3 virtual ranks on one GPU, uneven segments."""
    import torch
    from qoc_b200 import sharding
    cfg = o.config_synthetic(32, 61, nc=2, m=3, seed=9)
    P = 3
    Jo, go, co = o.evaluate(cfg, order=0)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])[1] if cfg["cost"] == o.COST_INFIDELITY else q.setup_infidelity_abs_trace(cfg["T"])[1]
    dev = torch.device("cuda", 0)
    S_all = torch.empty((P, 32, 32), dtype=torch.complex128, device=dev)
    engines = []
    for r in range(P):
        lo, hi = sharding.time_partition(61, P, r)
        e = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], hi - lo, cfg["x0"].shape[1], 0, order=0)
        e.set_builtin_cost(cost, cfg["x0"])
        u_dev = torch.from_numpy(np.ascontiguousarray(cfg["u"][:, lo:hi].T)).to(dev)
        S_all[r].copy_(e.phase1_cm(u_dev))
        engines.append(e)
    # the rank propagators themselves
    for r in range(P):
        lo, hi = sharding.time_partition(61, P, r)
        Sr = np.eye(32, dtype=complex)
        for k in range(lo, hi):
            Sr = co["Uk"][k] @ Sr
        assert np.abs(S_all[r].cpu().numpy().T - Sr).max() < 1e-12
    g = np.zeros_like(go)
    for r in range(P):
        lo, hi = sharding.time_partition(61, P, r)
        J, gl = engines[r].phase2(S_all, P, r)
        assert abs(float(J.cpu()[0]) - Jo) <= TOL_J
        g[:, lo:hi] = gl.cpu().numpy().T
    assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()


@pytest.mark.parametrize("d,nc,m,scale,sym", [(27, 1, 1, 6.0, True), (27, 2, 3, 3.0, True), (28, 1, 2, 11.0, False),
                                               (24, 2, 2, 2.5, True), (20, 3, 1, 4.0, False), (16, 2, 4, 9.0, True),
                                               (12, 1, 2, 2.2, True), (9, 2, 4, 5.0, False), (5, 1, 1, 30.0, True)])
def test_real_hamiltonian_fast_path_every_shape_class(d, nc, m, scale, sym, monkeypatch):
    """Generators with an exactly zero real plane (X = -i H dt, H real: symmetric as in the tunable-bus model, and
    non-symmetric, which the ABI does not forbid) above the low-degree switch take K1's real-plane instantiation.
    Against the oracle, against the general complex instantiation (QOC_NO_REALH=1), and U_k / dU_k themselves."""
    rng = np.random.default_rng(300 + d)
    def realH():
        H = rng.standard_normal((d, d))
        return (H + H.T) / 2 if sym else H
    H0 = realH(); H0 *= scale / np.abs(H0).sum(axis=0).max()
    A = []
    for _ in range(nc):
        Hj = realH(); A.append(-1j * Hj / np.abs(Hj).sum(axis=0).max())
    x0 = np.eye(d, m, dtype=complex)
    Tq, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    cfg = dict(A0=-1j * H0, A=A, u=rng.uniform(-0.5, 0.5, (nc, 20)), x0=x0, T=Tq[:, :m].copy(), cost=o.COST_INFIDELITY, n=m)
    Jo, go, co = o.evaluate(cfg, order=0)
    J, g, cache = gpu_eval(cfg, 0)
    assert_parity(J, g, Jo, go)
    tolU = 1e-12 * max(1.0, np.abs(co["Uk"]).max())
    assert np.abs(cache.Uk_vec - co["Uk"]).max() < tolU
    dU = np.array(cache.dUkdu)
    monkeypatch.setenv("QOC_NO_REALH", "1")
    J2, g2, cache2 = gpu_eval(cfg, 0)
    assert_parity(J, g, J2, g2)   # (non-symmetric H: the state norm grows and |J| is huge -- relative, as in assert_parity)
    assert np.abs(cache.Uk_vec - cache2.Uk_vec).max() < tolU
    assert np.abs(dU - np.array(cache2.dUkdu)).max() < 1e-11 * max(1.0, np.abs(dU).max())
    if sym:   # the symmetric case inverts N through the real SPD matrix N N^dagger; "2" keeps the complex in-kernel inverse
        monkeypatch.setenv("QOC_NO_REALH", "2")
        J4, g4, cache4 = gpu_eval(cfg, 0)
        assert_parity(J4, g4, Jo, go)
        assert np.abs(cache4.Uk_vec - co["Uk"]).max() < tolU
        assert np.abs(dU - np.array(cache4.dUkdu)).max() < 1e-11 * max(1.0, np.abs(dU).max())
    # propagate-only (expm without Jacobians) goes through the same instantiation
    monkeypatch.delenv("QOC_NO_REALH")
    cache3 = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=0)
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cache3)
    assert np.abs(cache3.x - co["x"]).max() < 1e-12 * max(1.0, np.abs(co["x"]).max())


@pytest.mark.parametrize("d,nc,m,scale", [(27, 1, 1, 6.0), (20, 2, 3, 3.0), (12, 1, 2, 9.0), (9, 2, 4, 5.0)])
def test_real_hamiltonian_fast_path_taylor_orders(d, nc, m, scale, monkeypatch):
    """The reference's truncated Taylor Jacobian (orders 1..4) on the real-plane instantiation: against the oracle's same-order
    formula and against the general complex instantiation."""
    rng = np.random.default_rng(900 + d)
    def symH():
        H = rng.standard_normal((d, d))
        return (H + H.T) / 2
    H0 = symH(); H0 *= scale / np.abs(H0).sum(axis=0).max()
    A = []
    for _ in range(nc):
        Hj = symH(); A.append(-1j * Hj / np.abs(Hj).sum(axis=0).max())
    Tq, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    cfg = dict(A0=-1j * H0, A=A, u=rng.uniform(-0.5, 0.5, (nc, 16)), x0=np.eye(d, m, dtype=complex), T=Tq[:, :m].copy(),
               cost=o.COST_INFIDELITY, n=m)
    for order in (1, 2, 3, 4):
        Jo, go, co = o.evaluate(cfg, order=order)
        J, g, cache = gpu_eval(cfg, order)
        assert_parity(J, g, Jo, go)
        dU = np.array(cache.dUkdu)
        monkeypatch.setenv("QOC_NO_REALH", "1")
        J2, g2, cache2 = gpu_eval(cfg, order)
        monkeypatch.delenv("QOC_NO_REALH")
        assert_parity(J, g, J2, g2)
        assert np.abs(dU - np.array(cache2.dUkdu)).max() < 1e-11 * max(1.0, np.abs(dU).max())


@pytest.mark.parametrize("d,nt,nc,m,order", [(4, 9, 2, 4, 3), (8, 33, 2, 8, 0), (9, 40, 2, 4, 0), (9, 20, 2, 4, 4), (6, 15, 5, 2, 0)])
def test_small_dimension_kernel_and_dmma_classes_agree(d, nt, nc, m, order, monkeypatch):
    """d <= 9 is served by the warp-per-slice kernel k1s_kernel (nc <= 4); the DMMA shape classes it replaces stay reachable
    (QOC_NO_K1S=1, and nc > 4).  Both against the oracle and against each other, U_k and dU_k/du_j included."""
    cfg = o.config_synthetic(d, nt, nc=nc, m=m, seed=4000 + 13 * d + nt)
    Jo, go, co = o.evaluate(cfg, order=order)
    J, g, cache = gpu_eval(cfg, order)
    assert_parity(J, g, Jo, go)
    assert np.abs(cache.Uk_vec - co["Uk"]).max() < 1e-12
    dU = np.array(cache.dUkdu)
    monkeypatch.setenv("QOC_NO_K1S", "1")
    J2, g2, cache2 = gpu_eval(cfg, order)
    assert_parity(J2, g2, Jo, go)
    assert np.abs(cache2.Uk_vec - co["Uk"]).max() < 1e-12
    assert np.abs(dU - np.array(cache2.dUkdu)).max() < 1e-12 * max(1.0, np.abs(dU).max())


@pytest.mark.parametrize("d,nc,m", [(7, 2, 3), (9, 4, 4), (3, 1, 1)])
@pytest.mark.parametrize("scale", [1e-3, 0.3, 3.0, 11.0, 40.0])
def test_small_dimension_kernel_norm_regimes(d, nc, m, scale):
    """k1s_kernel over every Pade degree ([5/5], [7/7], [13/13]) and 0..4 squarings, exact-Frechet and Taylor-3 mode,
    nc up to its limit of 4, ragged d (masked strips)."""
    cfg = o.config_synthetic(d, 14, nc=nc, m=m, seed=77 + d)
    cfg["A0"] = cfg["A0"] * scale
    cfg["A"] = [a * scale for a in cfg["A"]]
    for order in (0, 3):
        Jo, go, co = o.evaluate(cfg, order=order)
        J, g, cache = gpu_eval(cfg, order)
        assert abs(J - Jo) <= TOL_J
        assert np.abs(g - go).max() <= TOL_G * max(np.abs(go).max(), 1e-300)
        assert np.abs(cache.Uk_vec - co["Uk"]).max() < 1e-11
        assert np.abs(cache.x - co["x"]).max() < 1e-11


def test_propagators_and_jacobians_against_binary128_ground_truth():
    """U_k and the exact dU_k/du_j of the CUDA path against exp / Frechet derivative computed in binary128
    (oracle/qoc_quad.c), for the three K1 forms: k1s_kernel (zz, d = 9), the real-symmetric instantiation (bus, d = 27),
    the general complex low-degree instantiation (cavity, d = 24)."""
    import qoc_quad
    for cfg in (o.config_zz(), o.config_bus(Nt=40, tgate=1.4), o.config_cavity(12, Nt=30)):
        J, g, cache = gpu_eval(cfg, 0)
        Uk, dU = np.array(cache.Uk_vec), np.array(cache.dUkdu)
        nt = cfg["u"].shape[1]
        for k in (0, nt // 2, nt - 1):
            X = o.generator(cfg["A0"], cfg["A"], cfg["u"][:, k])
            for j, Aj in enumerate(cfg["A"]):
                U, L = qoc_quad.expm_quad(X, Aj)
                assert np.abs(Uk[k] - U).max() < 1e-13
                assert np.abs(dU[k, j] - L).max() < 1e-13 * max(1.0, np.abs(L).max())


@pytest.mark.parametrize("order", [0, 3])
def test_time_sharding_small_dimension_kernel_virtual_ranks(order):
    """Time-segment sharding with k1s_kernel feeding the phase API (d = 9, m = 4, nc = 2, uneven segments, both gradient modes):
    one-call phase 2 over the gathered rank propagators, P virtual ranks on one GPU."""
    import torch
    from qoc_b200 import sharding
    cfg = o.config_synthetic(9, 301, nc=2, m=4, seed=123)
    P = 3
    Jo, go, co = o.evaluate(cfg, order=order)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])[1]
    dev = torch.device("cuda", 0)
    S_all = torch.empty((P, 9, 9), dtype=torch.complex128, device=dev)
    engines = []
    for r in range(P):
        lo, hi = sharding.time_partition(301, P, r)
        e = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], hi - lo, cfg["x0"].shape[1], 0, order=order)
        e.set_builtin_cost(cost, cfg["x0"])
        u_dev = torch.from_numpy(np.ascontiguousarray(cfg["u"][:, lo:hi].T)).to(dev)
        S_all[r].copy_(e.phase1_cm(u_dev))
        engines.append(e)
    g = np.zeros_like(go)
    for r in range(P):
        lo, hi = sharding.time_partition(301, P, r)
        J, gl = engines[r].phase2(S_all, P, r)
        assert abs(float(J.cpu()[0]) - Jo) <= TOL_J
        g[:, lo:hi] = gl.cpu().numpy().T
    assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()


@pytest.mark.parametrize("d,nt,order", [(32, 91, 0), (40, 60, 3)])
def test_time_sharding_general_path_with_a_host_closure_cost(d, nt, order):
    """The three-call phase API (phase 1 / forward / backward) on the GENERAL path with a cost that is an arbitrary host closure
    (here the z-calibrated infidelity's host twin and the oracle's trace infidelity): qoc_shard_forward_device /
    qoc_shard_backward_device route through the two-level sweeps of qoc_gpath.cuh (round 1 launched the shared-memory kernels
    with d > 28: ADVICE.md).  P = 3 virtual ranks on one GPU."""
    import torch
    from qoc_b200 import sharding
    cfg = o.config_synthetic(d, nt, nc=2, m=4, seed=3)
    P = 3
    for cost_o, cost_h in ((o.setup_infidelity(cfg["T"], cfg["n"]), o.setup_infidelity(cfg["T"], cfg["n"])),
                           (o.setup_infidelity_zcalibrated(cfg["T"]), q.setup_infidelity_zcalibrated(cfg["T"], device=False))):
        co = o.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape)
        xs = o.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], co)["x"]
        Jo = cost_o[0](xs[-1])
        go = o.grape_sensitivity(cfg["A0"], cfg["A"], cost_o[1], cfg["u"], cfg["x0"], co, dUkdp_order=order).copy()
        engines, S = [], []
        for r in range(P):
            lo, hi = sharding.time_partition(nt, P, r)
            e = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], hi - lo, 4, 0, order=order)
            S.append(e.phase1(cfg["u"][:, lo:hi]).clone())
            engines.append(e)
        x = torch.as_tensor(cfg["x0"]).to(engines[0].device)
        starts = []
        for r in range(P):
            starts.append(x)
            x = S[r] @ x
        xN = x.cpu().numpy()
        assert np.abs(xN - xs[-1]).max() < 1e-11
        assert abs(float(cost_h[0](xN)) - Jo) <= TOL_J
        lam = torch.as_tensor(np.asarray(cost_h[1](xN), dtype=np.complex128)).to(engines[0].device)
        g = np.zeros_like(go)
        for r in range(P - 1, -1, -1):
            lo, hi = sharding.time_partition(nt, P, r)
            engines[r].forward(starts[r])
            gl, _ = engines[r].backward(lam)
            g[:, lo:hi] = gl.cpu().numpy()
            lam = S[r].conj().t() @ lam
        # (the z-calibrated gradient is defined to ~1e-7 by the reference's own golden-section search: tests/test_gpu_zcal.py)
        tol = 1e-6 if cost_h[1] is not cost_o[1] else TOL_G
        assert np.abs(g - go).max() <= tol * np.abs(go).max()
