"""The z-calibrated infidelity (src/penalty_fcns.jl:27-42, src/fidelities.jl:11-56, 81-137) evaluated ON THE DEVICE
(QOC_COST_ZCAL): m = diag(T'x) reduced in the sweep kernels, golden-section search and rrule per pulse in one thread,
lambda_N formed on the device.  Run with `-m gpu`; every call goes through the C ABI.

Checks: (1) every known answer of test/test_fidelities.jl:19-84 reproduced by the device search; (2) J and dJ/du against
the oracle's closures on the zz model, every K1 / sweep code path (small-dimension kernel, DMMA classes, general path);
(3) a 4096-pulse batch against the oracle on a sample and against the host-closure route on all of it."""
import numpy as np
import pytest

import qoc_oracle as o
import qoc_b200 as q
from qoc_b200 import _lib

pytestmark = pytest.mark.gpu
TOL_J, TOL_G = 1e-10, 1e-8
cis = lambda t: np.exp(1j * np.asarray(t, dtype=float))

# test/test_fidelities.jl: (m, abs_sum_phase_calibrated(m))
KATS = [
    (np.array([1, 1j, 1j, 1]), 2.8284271),                         # :19
    (np.array([1, 0.1j, 0.1j, 1]), 2.0099751),                     # :30
    (cis([1, 2, 3, 4]), 4.0),                                      # :40
    (cis([1, 2, -2.5, -1.7]), 3.995001),                           # :53
    (cis([2.5, 2.5, 1.5, -2.5]), 3.365883939061934),               # :72-74
    (np.array([0.65 - 0.75j, -0.4 + 0.8j, -0.4 + 0.1j, 0.7]), 2.9787244710195484),   # :80-84
]


def _identity_problem(mvec, d=4):
    """A problem whose final state has diag(T'x_N) = mvec exactly: U_k = exp(0) = I, x0 = Diagonal(mvec), T = I."""
    A0 = np.zeros((d, d), dtype=complex)
    rng = np.random.default_rng(3)
    H = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
    A = [-1j * (H + H.conj().T) * 0.05]
    x0 = np.zeros((d, 4), dtype=complex)
    x0[:4, :4] = np.diag(mvec)
    T = np.eye(d, 4, dtype=complex)
    return A0, A, x0, T


@pytest.mark.parametrize("kat", range(len(KATS)))
@pytest.mark.parametrize("d", [4, 16, 32])   # small-dimension kernel | DMMA class | general path
def test_device_search_reproduces_reference_known_answers(kat, d):
    mvec, F_ref = KATS[kat]
    A0, A, x0, T = _identity_problem(mvec, d)
    u = np.zeros((1, 6))
    Jf, dJf = q.setup_infidelity_zcalibrated(T)
    cache = q.setup_grape_cache(A0, x0, u.shape, dUkdp_order=0)
    J, g = q.evaluate(cache, A0, A, u, x0, dJf, dUkdp_order=0)
    # the reference's `@test a ≈ b` is isapprox with rtol = sqrt(eps) = 1.5e-8; three values are quoted to 7-8 digits only
    tol = 2e-7 if F_ref in (2.8284271, 2.0099751, 3.995001) else 1.5e-8 * F_ref
    assert abs(np.sqrt(16 * (1 - J)) - F_ref) < tol
    # the oracle's restatement of the search on the same four numbers: same bracket updates -> same F to rounding
    assert abs(np.sqrt(16 * (1 - J)) - o.abs_sum_phase_calibrated(mvec)) < 1e-12
    # gradient: the oracle with its own closures
    Jo, dJo = o.setup_infidelity_zcalibrated(T)
    co = o.setup_grape_cache(A0, x0, u.shape)
    o.propagate(A0, A, u, x0, co)
    go = o.grape_sensitivity(A0, A, dJo, u, x0, co, dUkdp_order=0)
    # Tolerance: the reference locates theta by golden-section search on a flat maximum, which cannot resolve the maximiser
    # better than ~sqrt(eps) (the last comparisons are decided by rounding noise of cos / sqrt, which differ between libm
    # implementations); dF_dm depends on theta to first order, so the gradient is only defined to ~1e-8..1e-7 relative by
    # the reference's own algorithm.  Its test compares the rrule with finite differences at rtol 1e-6
    # (test/test_fidelities.jl:130-148): the same bar here.
    assert np.abs(g - go).max() <= 1e-6 * max(np.abs(go).max(), 1e-3)
    cache.close()


@pytest.mark.parametrize("order", [0, 3])
@pytest.mark.parametrize("force_dmma", [False, True])
def test_zz_zcalibrated_vs_oracle(order, force_dmma, monkeypatch):
    if force_dmma:
        monkeypatch.setenv("QOC_NO_K1S", "1")
    cfg = o.config_zz()
    Jo, dJo = o.setup_infidelity_zcalibrated(cfg["T"])
    co = o.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape)
    xs = o.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], co)["x"]
    Jref = Jo(xs[-1])
    gref = o.grape_sensitivity(cfg["A0"], cfg["A"], dJo, cfg["u"], cfg["x0"], co, dUkdp_order=order)
    Jf, dJf = q.setup_infidelity_zcalibrated(cfg["T"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], dJf, dUkdp_order=order)
    assert abs(J - Jref) <= TOL_J * max(1.0, abs(Jref))
    assert np.abs(g - gref).max() <= TOL_G * np.abs(gref).max()
    # the reference-style split calls: propagate with the built-in cost (J on the device), then grape_sensitivity
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cache, Jfinal=Jf)
    assert abs(cache.J - Jref) <= TOL_J
    g2 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], cache, dUkdp_order=order)
    assert np.abs(g2 - gref).max() <= TOL_G * np.abs(gref).max()
    # and the reference's own route (host closures: x[end] down, terminal costate up) gives the same numbers
    Jh, dJh = q.setup_infidelity_zcalibrated(cfg["T"], device=False)
    g3 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJh, cfg["u"], cfg["x0"], cache, dUkdp_order=order)
    assert np.abs(g3 - g2).max() <= 1e-12 * np.abs(gref).max() + 1e-15
    cache.close()


def test_wrong_column_count_is_the_reference_error():
    cfg = o.config_zz()
    with pytest.raises(ValueError, match="Only works for two-qubit gates"):
        q.setup_infidelity_zcalibrated(cfg["T"][:, :3])
    # and at the C ABI: a three-column problem cannot select the cost (src/penalty_fcns.jl:28-30)
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"][:, :3], cfg["u"].shape)
    cache._ensure(cfg["A0"], cfg["A"], cfg["x0"][:, :3])
    lib = _lib.load()
    T3 = np.asfortranarray(cfg["T"][:, :3])
    rc = lib.qoc_set_cost(cache.handle, _lib.COST_ZCAL, T3.ctypes.data_as(_lib.C.POINTER(_lib.C.c_double)), 4)
    assert rc == _lib.ERR_DIMENSION
    cache.close()


def test_zz_batch_4096_zcalibrated_on_device():
    """C4 with the z-calibrated cost: 4096 golden-section searches on the device, no x_final round trip."""
    nb = 4096
    cfg = o.config_zz_batch(nb)
    ub = cfg["u_batch"]
    Jf, dJf = q.setup_infidelity_zcalibrated(cfg["T"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], ub.shape[-2:], batch=nb, dUkdp_order=3, store_costates=False)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], ub, cfg["x0"], dJf, dUkdp_order=3)
    assert J.shape == (nb,) and g.shape == ub.shape and np.all(np.isfinite(J)) and np.all(np.isfinite(g))
    # oracle on a sample of the pulses
    Jo, dJo = o.setup_infidelity_zcalibrated(cfg["T"])
    for b in (0, 1, 777, 2048, 4095):
        co = o.setup_grape_cache(cfg["A0"], cfg["x0"], ub[b].shape)
        xs = o.propagate(cfg["A0"], cfg["A"], ub[b], cfg["x0"], co)["x"]
        gref = o.grape_sensitivity(cfg["A0"], cfg["A"], dJo, ub[b], cfg["x0"], co, dUkdp_order=3)
        assert abs(J[b] - Jo(xs[-1])) <= TOL_J
        assert np.abs(g[b] - gref).max() <= TOL_G * np.abs(gref).max()
    # every pulse: the host-closure route of the product (x[end] down, vectorised host search, costates up)
    Jh, dJh = q.setup_infidelity_zcalibrated(cfg["T"], device=False)
    q.propagate(cfg["A0"], cfg["A"], ub, cfg["x0"], cache)
    assert np.abs(J - np.asarray(Jh(cache.x_final))).max() <= TOL_J   # x_final: (batch, d, 4)
    gh = q.grape_sensitivity(cfg["A0"], cfg["A"], dJh, ub, cfg["x0"], cache, dUkdp_order=3)
    assert np.abs(g - gh).max() <= TOL_G * np.abs(gh).max()
    cache.close()
