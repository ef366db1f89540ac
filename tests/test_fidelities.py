"""Product-side fidelity functions (quantumoptimalcontrol.jl_b200/fidelities.py) against the reference's known answers
(test/test_fidelities.jl) and against the oracle's scalar restatement; CPU only."""
import numpy as np
import pytest

import qoc_oracle as o
import qoc_b200 as q


def cis(t):
    return np.exp(1j * np.asarray(t, dtype=float))


KATS = [  # (m, optimal, basic)   test/test_fidelities.jl:17-84
    ([1, 1j, 1j, 1], 2.8284271, 2.0),
    ([1, 0.1j, 0.1j, 1], 2.0099751, 0.2),
    (cis([1, 2, 3, 4]), 4.0, 4.0),
    (cis([1, 2, -2.5, -1.7]), 3.995001, None),
    (cis([2.5, 2.5, 1.5, -2.5]), 3.365883939061934, None),
    ([0.65 - 0.75j, -0.4 + 0.8j, -0.4 + 0.1j, 0.7], 2.9787244710195484, None),
]


@pytest.mark.parametrize("m,val,basic", KATS)
def test_known_answers(m, val, basic):
    assert q.abs_sum_phase_calibrated(m) == pytest.approx(val, abs=2e-7)
    assert q.abs_sum_phase_calibrated(m, "grid") == pytest.approx(val, abs=1e-3)
    if basic is not None:
        assert q.abs_sum_phase_calibrated(m, "basic") == pytest.approx(basic, abs=1e-7)
    F, th = q.optimal_calibration(m)
    Jm = abs(m[0] + m[1] * cis(th[0]) + m[2] * cis(th[1]) + m[3] * cis(th[0] + th[1]))
    assert Jm == pytest.approx(F, abs=1e-8)


def test_theta_known_answer_and_batch():
    m = cis([1, 2, -2.5, -1.7])
    assert np.allclose(q.optimal_calibration(m)[1], [5.383258515112539, 3.6000220820575084], atol=1e-4)
    # the whole KAT table as one batch == element by element
    M = np.array([np.asarray(k[0], dtype=complex) for k in KATS])
    Fb, thb = q.optimal_calibration(M)
    for i, k in enumerate(KATS):
        assert Fb[i] == pytest.approx(k[1], abs=2e-7)


def test_matches_oracle_on_random_batch():
    rng = np.random.default_rng(0)
    M = rng.random((500, 4)) * cis(2 * np.pi * rng.random((500, 4)))  # test_fidelities.jl:110
    for cal in ("optimal", "basic", "none", "grid", "lms_phase", "lms_phase2", "lms_phase3"):
        vb = q.abs_sum_phase_calibrated(M, cal)
        vo = np.array([o.abs_sum_phase_calibrated(m, cal) for m in M])
        assert np.abs(vb - vo).max() < 1e-12, cal
    F, dF = q.abs_sum_phase_calibrated_rrule(M)
    for i in (0, 17, 499):
        Fo, dFo = o.abs_sum_phase_calibrated_rrule(M[i])
        assert abs(F[i] - Fo) < 1e-12 and np.abs(dF[i] - dFo).max() < 5e-7  # theta is only located to 1e-9 (golden section)
    # :optimal is at least as good as :grid and never much better  (test_fidelities.jl:116-117)
    Fg = q.abs_sum_phase_calibrated(M, "grid")
    assert np.all(F - Fg > -1e-12) and np.all(F - Fg < 4e-3)


def test_zcalibrated_cost_and_infidelity():
    rng = np.random.default_rng(1)
    Q, _ = np.linalg.qr(rng.standard_normal((9, 8)) + 1j * rng.standard_normal((9, 8)))
    T = Q[:, :4]
    x = rng.standard_normal((9, 4)) + 1j * rng.standard_normal((9, 4))
    J, dJ = q.setup_infidelity_zcalibrated(T)
    Jo, dJo = o.setup_infidelity_zcalibrated(T)
    assert J(x) == pytest.approx(Jo(x), abs=1e-13)
    assert np.abs(dJ(x) - dJo(x)).max() < 5e-7
    xb = np.stack([x, 0.5 * x + 0.1, x[::-1]])
    assert np.allclose(J(xb), [Jo(v) for v in xb], atol=1e-13)
    assert np.abs(dJ(xb)[2] - dJo(xb[2])).max() < 5e-7
    with pytest.raises(ValueError, match="four columns"):
        q.setup_infidelity_zcalibrated(T[:, :3])
    U = np.diag(cis([0.3, 1.1, -0.4, 0.4]))
    assert abs(q.infidelity(np.eye(4), U, "optimal")) < 1e-9
    assert q.infidelity(np.eye(4), U) == pytest.approx(o.infidelity(np.eye(4), U), abs=1e-13)
    with pytest.raises(ValueError, match="Not supported yet"):
        q.infidelity(np.eye(3), np.eye(3))
