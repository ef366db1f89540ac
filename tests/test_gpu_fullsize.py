"""Parity at the sizes BASELINE.json names (run with `-m gpu` on a B200), through the C ABI.

The checker for the large cases is the C restatement oracle/qoc_ref.c (the numpy oracle needs minutes at d = 80); the two
agree to 1e-15 on every config they share (tests/test_oracle.py).  Where even the C restatement is too slow (Nt = 1e5) the
checks are size-independent properties plus the binary128 bound on the reassociation error of the parallel scans
(SURVEY.md F7): the serial product of the SAME U_k accumulated in __float128 (oracle/qoc_quad.c) is the ground truth,
so what is measured is purely the effect of the scan's association order and of double rounding.

Tolerances (north star): |dJ| <= 1e-10 max(1, |J|), |dg_jk| <= 1e-8 max|g|.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import qoc_oracle as o
import qoc_quad
import qoc_ref
import qoc_b200 as q

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_J, TOL_G = 1e-10, 1e-8


def cost_of(cfg):
    return q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == o.COST_INFIDELITY else q.setup_infidelity_abs_trace(cfg["T"])


def gpu_eval(cfg, order, store_costates=False):
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order, store_costates=store_costates)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost_of(cfg)[1], dUkdp_order=order)
    return J, g, cache


# ---- C3 at the full pulse length, every truncation BASELINE.json names -------------------------------------------------
@pytest.mark.parametrize("ncav", [12, 20, 40])
@pytest.mark.parametrize("order", [0, 3])
def test_cavity_full_pulse_vs_c_restatement(ncav, order):
    cfg = o.config_cavity(ncav, Nt=550)
    r = qoc_ref.ref_eval(cfg, order=order)
    J, g, cache = gpu_eval(cfg, order)
    assert abs(J - r["J"]) <= TOL_J * max(1.0, abs(r["J"]))
    assert np.abs(g - r["dJdu"]).max() <= TOL_G * np.abs(r["dJdu"]).max()
    cache.close()


# ---- C5 at Nt >= 1e4 (d = 16 on-chip path, d = 32 / 64 general path), d = 128 at a reduced length ---------------------
@pytest.mark.parametrize("d,nt", [(16, 20000), (32, 10000), (64, 10000), (128, 1000)])
def test_synthetic_long_pulse_vs_c_restatement(d, nt):
    cfg = o.config_synthetic(d, nt)
    r = qoc_ref.ref_eval(cfg, order=0)
    J, g, cache = gpu_eval(cfg, 0)
    assert abs(J - r["J"]) <= TOL_J * max(1.0, abs(r["J"]))
    assert np.abs(g - r["dJdu"]).max() <= TOL_G * np.abs(r["dJdu"]).max()
    cache.close()


# ---- reassociation error of the scans, bounded against binary128 ------------------------------------------
@pytest.mark.parametrize("d,nt", [(16, 20000), (27, 10000), (40, 4000)])
def test_scan_reassociation_error_vs_binary128(d, nt):
    """x_N from the CUDA path (segment products + two-level boundary walk) against the SERIAL product of the same U_k in
    binary128; next to it the serial product in double, i.e. what the reference's loop (src/gradient_computations.jl:27-29)
    itself loses.  The scan may not lose more than a small multiple of the serial loop, and the extrapolation to Nt = 1e5
    (errors grow at most linearly in Nt) has to stay below the 1e-10 bar on J."""
    cfg = o.config_bus(Nt=nt, tgate=0.035 * nt) if d == 27 else o.config_synthetic(d, nt)
    cache = q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"])
    U = cache.Uk_vec
    x_gpu = cache.x_final.reshape(cfg["x0"].shape)
    x_quad = qoc_quad.chain_quad(U, cfg["x0"])
    x = np.array(cfg["x0"], dtype=np.complex128)
    for k in range(nt):
        x = U[k] @ x
    e_scan = np.abs(x_gpu - x_quad).max()
    e_serial = np.abs(x - x_quad).max()
    print(f"d={d} Nt={nt}: scan error {e_scan:.2e}, serial double loop {e_serial:.2e}")
    assert e_scan <= max(4.0 * e_serial, 2e-13)
    # J = 1 - |tr(T'x)|^2/n^2: |dJ| <= 2 |T| |dx| / n, T has unit-norm columns
    assert 2.0 * e_scan * (1e5 / nt) <= TOL_J
    cache.close()


# ---- BASELINE sizes, size-independent properties ----------------------------------------------------------------------
def test_long_pulse_1e5_properties():
    """C5 d = 16, Nt = 1e5: unitarity of the propagated columns, prefix consistency (the first half of the pulse evaluated on
    its own reproduces the state at Nt/2), gradient of the halves vs the whole through the costate at the cut."""
    d, nt = 16, 100000
    cfg = o.config_synthetic(d, nt)
    cache = q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"])
    xf = cache.x_final
    G = xf.conj().T @ xf
    assert np.abs(G - np.eye(G.shape[0])).max() < 1e-10
    half = dict(cfg)
    half["u"] = np.ascontiguousarray(cfg["u"][:, : nt // 2])
    c2 = q.propagate(half["A0"], half["A"], half["u"], half["x0"])
    x_mid = cache.x[nt // 2]
    assert np.abs(c2.x_final - x_mid).max() < 1e-10
    cache.close(); c2.close()


# ---- real ranks: time-segment sharding over 2 GPUs vs one GPU vs the C restatement -------------------------------------
def test_time_sharding_on_two_real_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (the virtual-rank tests in test_gpu_parity.py cover the same code on one)")
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tests", "two_rank_parity.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "TWO_RANK_PARITY_OK" in r.stdout


# ---- streamed Jacobians (what lets C5 d = 256, Nt = 1e5 run on one 180 GB device) -----------------------------------------
@pytest.mark.parametrize("d,nt,order", [(32, 5000, 0), (64, 2500, 3), (40, 700, 0)])
def test_streamed_jacobians_equal_stored_jacobians(d, nt, order, monkeypatch):
    """QOC_STREAM_JAC=1 forces the mode a handle picks by itself when U_k plus every dU_k/du_j would not fit the device:
    dU_k/du_j of one chunk at a time, re-formed after the sweeps and contracted at once.  Same arithmetic, same order of
    operations per slice -> the gradient must agree with the stored-Jacobian evaluation to rounding (1e-12), several chunks."""
    cfg = o.config_synthetic(d, nt) if d != 40 else o.config_cavity(20, Nt=550)
    J0, g0, c0 = gpu_eval(cfg, order)
    c0.close()
    monkeypatch.setenv("QOC_STREAM_JAC", "1")
    J1, g1, c1 = gpu_eval(cfg, order)
    assert abs(J1 - J0) <= 1e-13
    assert np.abs(g1 - g0).max() <= 1e-12 * np.abs(g0).max()
    # the split calls (propagate, then grape_sensitivity with a host closure) go through the same streamed pass
    Jf, dJf = o.cost_closures(cfg)
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], c1)
    g2 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], c1, dUkdp_order=order)
    assert np.abs(g2 - g0).max() <= 1e-12 * np.abs(g0).max()
    with pytest.raises(q.QOCError):   # nothing is stored: the Jacobian getter says so
        _ = c1.dUkdu
    c1.close()


def test_streamed_jacobians_time_sharded_in_library(monkeypatch):
    import torch
    from qoc_b200 import sharding
    cfg = o.config_synthetic(32, 6000)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    J0, g0, c0 = gpu_eval(cfg, 0)
    c0.close()
    monkeypatch.setenv("QOC_STREAM_JAC", "1")
    ngpu = torch.cuda.device_count()
    sh = sharding.InProcessSharded(cfg["A0"], cfg["A"], cfg["x0"], cost[1], cfg["u"].shape, [r % ngpu for r in range(3)], kind="time",
                                   dUkdp_order=0)
    J, g = sh.evaluate(cfg["u"])
    assert abs(J - J0) <= 1e-10 and np.abs(g - g0).max() <= 1e-8 * np.abs(g0).max()
    sh.close()


# ---- control bounds: the general path without its one synchronisation, CUDA-graph capture --------------------------------
def test_control_bounds_make_the_general_path_asynchronous_and_capturable():
    import ctypes as C
    import torch
    from qoc_b200 import _lib
    from qoc_b200.grape import _u_arr
    lib = _lib.load()
    cfg = o.config_synthetic(32, 300)
    J0, g0, cache = gpu_eval(cfg, 0)
    dp = C.POINTER(C.c_double)
    ub = np.array([0.5, 0.5])    # config_synthetic draws u ~ U(-0.5, 0.5)
    assert lib.qoc_set_control_bounds(cache.handle, ub.ctypes.data_as(dp)) == 0
    J1, g1 = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost_of(cfg)[1], dUkdp_order=0)
    assert abs(J1 - J0) <= 1e-13 and np.abs(g1 - g0).max() <= 1e-12 * np.abs(g0).max()
    # device-resident evaluation captured in a CUDA graph and replayed on new inputs
    dev = torch.device("cuda", 0)
    d_u = torch.from_numpy(_u_arr(cfg["u"], cache)).to(dev)
    d_J = torch.zeros(1, dtype=torch.float64, device=dev)
    d_g = torch.zeros(d_u.shape, dtype=torch.float64, device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        rc = lib.qoc_eval_device(cache.handle, C.c_void_p(d_u.data_ptr()), C.c_void_p(d_J.data_ptr()), C.c_void_p(d_g.data_ptr()),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    u2 = 0.9 * cfg["u"]
    d_u.copy_(torch.from_numpy(_u_arr(u2, cache)))
    graph.replay()
    torch.cuda.synchronize()
    c2 = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=0, store_costates=False)
    J2, g2 = q.evaluate(c2, cfg["A0"], cfg["A"], u2, cfg["x0"], cost_of(cfg)[1], dUkdp_order=0)
    assert abs(float(d_J.cpu()[0]) - J2) <= 1e-12
    assert np.abs(d_g.cpu().numpy().T - g2).max() <= 1e-12 * np.abs(g2).max()
    c2.close()
    # a broken promise is reported, not silently under-scaled
    with pytest.raises(q.QOCError, match="exceeds the bound"):
        q.evaluate(cache, cfg["A0"], cfg["A"], 1.5 * cfg["u"], cfg["x0"], cost_of(cfg)[1], dUkdp_order=0)
    cache.close()


# ---- running state penalty on the general path: two-level (parallel) sweeps, any row ------------------------------------------
@pytest.mark.parametrize("d,nt,m,rows,cols,order", [(32, 300, 3, [1, 5, 31], [0, 2], 0), (40, 150, 2, [3, 39], [0, 1], 3),
                                                     (80, 60, 2, [2, 64, 65, 79], [1], 0),
                                                     # five and more columns: the penalty accumulator used to share a slot with
                                                     # the overlap of column 4 (found by tools/stress_parity.py)
                                                     (33, 40, 5, [21, 28], [3], 2), (44, 12, 7, [38, 40], [1, 2], 0),
                                                     (28, 15, 8, [2], [0, 7], 0)])
def test_state_penalty_general_path_vs_oracle(d, nt, m, rows, cols, order):
    """src/penalty_fcns.jl:1-11 with src/gradient_computations.jl:47-49, :55-57 at d > 28: the costate recurrence is affine, a
    segment is summarised by (Q_seg, c_seg); rows >= 64 included (round 1's serial sweep carried a 64-bit row mask)."""
    cfg = o.config_synthetic(d, nt, nc=2, m=m, seed=d)
    pen = (rows, cols, 0.41)
    Jo, go, co = o.evaluate(cfg, order=order, penalty=pen)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost[1], dUkdp_order=order, penalty=q.setup_state_penalty(*pen))
    assert abs(J - Jo) <= TOL_J * max(1.0, abs(Jo))
    assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()
    assert np.abs(cache.lam - co["lam"]).max() < 1e-11
    # the reference's split calls with host closures for cost and penalty (examples/ipopt_callbacks_exp.jl:16-18, :27)
    Jf, dJf = o.setup_infidelity(cfg["T"], cfg["n"])
    L, dL = q.setup_state_penalty(*pen)
    c2 = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape)      # no built-in cost on this handle
    q.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], c2, penalty=(L, dL))
    x = c2.x
    assert abs(Jf(x[-1]) + sum(L(xk) for xk in x) - Jo) <= TOL_J * max(1.0, abs(Jo))
    assert abs(c2.J - sum(L(xk) for xk in x)) < 1e-11                  # device-side running sum of the penalty alone
    g2 = q.grape_sensitivity(cfg["A0"], cfg["A"], dJf, cfg["u"], cfg["x0"], c2, dUkdp_order=order, dL_dx=dL)
    assert np.abs(g2 - go).max() <= TOL_G * np.abs(go).max()
    cache.close(); c2.close()


def test_state_penalty_small_case_forced_onto_the_general_path(golden_dir, monkeypatch):
    monkeypatch.setenv("QOC_FORCE_GPATH", "1")
    gd = np.load(os.path.join(golden_dir, "zz_penalty_order4.npz"))
    cfg = o.config_zz()
    pen = q.setup_state_penalty([6, 7, 8], [0, 1, 2, 3], 0.22)
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (2, 100), dUkdp_order=4)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], dUkdp_order=4, penalty=pen)
    assert abs(J - float(gd["J"])) <= TOL_J and np.abs(g - gd["dJdu"]).max() <= TOL_G * np.abs(gd["dJdu"]).max()
    cache.close()
