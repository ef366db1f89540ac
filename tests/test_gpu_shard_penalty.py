"""Running state penalty (src/penalty_fcns.jl:1-11) under TIME sharding: the affine costate recurrence of
src/gradient_computations.jl:47-57 split over ranks through qoc_shard_forward_device / qoc_shard_affine_device /
qoc_shard_backward_device.  P virtual ranks on one GPU with the boundary algebra written out here (independent of
quantumoptimalcontrol.jl_b200/sharding.py, whose own version is tested under gloo on CPU and with two real ranks), against the
oracle's serial evaluation.  Every sweep family: small-dimension (d <= 9), shared-memory first generation, general path,
streamed Jacobians.  Run with `-m gpu`."""
import numpy as np
import pytest

import qoc_oracle as o
import qoc_b200 as q
from qoc_b200 import _lib, sharding

pytestmark = pytest.mark.gpu
TOL_J, TOL_G = 1e-10, 1e-8


def _dL(pen, x):
    import torch
    rows = torch.as_tensor(pen.rows, device=x.device, dtype=torch.long)
    cols = torch.as_tensor(pen.cols, device=x.device, dtype=torch.long)
    out = torch.zeros_like(x)
    out[rows[:, None], cols[None, :]] = 2 * pen.mu * x[rows][:, cols]
    return out, float(pen.mu * torch.sum(torch.abs(x[rows][:, cols]) ** 2))


CASES = [
    # d, nt, nc, m, order, P, penalty (rows, cols, mu), eager/streamed
    (4, 301, 2, 4, 3, 3, ([2, 3], [0, 1, 3], 0.8), False),        # small-dimension sweeps (K3S)
    (9, 400, 3, 4, 0, 4, ([0, 8], [1, 2], 1.3), False),
    (16, 333, 2, 3, 0, 3, ([9, 12, 15], [0, 2], 0.37), False),    # shared-memory first-generation K2 / K3
    (27, 257, 2, 4, 3, 2, ([20, 26], [0, 1, 2, 3], 0.5), False),
    (32, 91, 2, 4, 0, 3, ([1, 30, 31], [0, 3], 0.9), False),      # general path, two-level sweeps
    (40, 60, 2, 5, 3, 3, ([7, 39], [0, 1, 4], 2.0), False),       # ... five columns
    (72, 48, 1, 2, 0, 2, ([64, 70, 71], [0, 1], 0.6), False),     # ... penalised rows >= 64 (byte mask)
    (32, 120, 2, 4, 0, 3, ([5, 6], [1, 2], 1.1), True),           # streamed Jacobians
]


@pytest.mark.parametrize("d,nt,nc,m,order,P,pen,streamed", CASES)
def test_time_sharded_running_penalty(d, nt, nc, m, order, P, pen, streamed, monkeypatch):
    import torch
    if streamed:
        monkeypatch.setenv("QOC_STREAM_JAC", "1")
    cfg = o.config_synthetic(d, nt, nc=nc, m=m, seed=11)
    Jo, go, co = o.evaluate(cfg, order=order, penalty=pen)
    J0, g0, _ = o.evaluate(cfg, order=order)
    assert abs(Jo - J0) > 1e-4     # the penalty is not negligible here
    L, _ = q.setup_state_penalty(*pen)
    Jf, dJf = o.cost_closures(cfg)
    engines, S = [], []
    for r in range(P):
        lo, hi = sharding.time_partition(nt, P, r)
        e = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], hi - lo, m, 0, order=order, penalty=L)
        S.append(e.phase1(cfg["u"][:, lo:hi]).clone())
        engines.append(e)
    dev = engines[0].device
    x = torch.as_tensor(cfg["x0"]).to(dev)
    starts = []
    for r in range(P):
        starts.append(x)
        x = S[r] @ x
    xN = x.cpu().numpy()
    assert np.abs(xN - co["x"][-1]).max() < 1e-11
    # phase 2a on every rank: forward, affine term, local sum of L
    cs, J = [], float(Jf(xN))
    for r in range(P):
        xe = engines[r].forward(starts[r])
        assert torch.abs(xe - (starts[r + 1] if r + 1 < P else x)).max() < 1e-11
        c, Jp = engines[r].affine()
        cs.append(c.clone())
        J += float(Jp.cpu()[0])
        if r >= 1:
            J -= _dL(L, starts[r])[1]
    assert abs(J - Jo) <= TOL_J * max(1.0, abs(Jo))
    # phase 2b: boundary costates walked down from the last rank, local backward sweeps
    lam = torch.as_tensor(np.asarray(dJf(xN), dtype=np.complex128)).to(dev)
    g = np.zeros_like(go)
    for r in range(P - 1, -1, -1):
        lo, hi = sharding.time_partition(nt, P, r)
        gl, ls = engines[r].backward(lam)
        g[:, lo:hi] = gl.cpu().numpy()
        # the affine relation the exchange relies on, against the backward sweep's own lambda_start
        pred = S[r].conj().t() @ lam + cs[r]
        assert torch.abs(pred - ls).max() <= 1e-11 * max(1.0, float(torch.abs(ls).max()))
        lam = pred - _dL(L, starts[r])[0]
    assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()
    # the same pulse through the product's evaluator (one rank: forward / affine / backward and its own bookkeeping)
    e1 = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], nt, m, 0, order=order, penalty=L)
    ev = sharding.TimeShardedEvaluator(e1, cfg["x0"], (Jf, dJf), nt)
    J1, g1 = ev.evaluate(cfg["u"])
    assert abs(J1 - Jo) <= TOL_J * max(1.0, abs(Jo)) and np.abs(g1 - go).max() <= TOL_G * np.abs(go).max()
    for e in engines + [e1]:
        e.cache.close()


def test_phase_api_errors_with_penalty():
    """The one-call phase 2 does not carry the penalty (QOC_ERR_UNSUPPORTED, no silent fallback); the affine call needs the
    forward call first and a penalty to exist."""
    import ctypes as C
    import torch
    cfg = o.config_synthetic(16, 64, nc=2, m=3, seed=1)
    L, _ = q.setup_state_penalty([3, 4], [0, 1], 0.5)
    lib = _lib.load()
    e = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], 64, 3, 0, order=0, penalty=L)
    e.phase1(cfg["u"])
    c = torch.empty((3, 16), dtype=torch.complex128, device=e.device)
    rc = lib.qoc_shard_affine_device(e.cache.handle, C.c_void_p(c.data_ptr()), None, None)
    assert rc == _lib.ERR_STALE_CACHE
    S_all = torch.zeros((1, 16, 16), dtype=torch.complex128, device=e.device)
    with pytest.raises(q.QOCError) as ei:
        e.phase2(S_all, 1, 0)
    assert ei.value.status == _lib.ERR_UNSUPPORTED
    e.cache.close()
    e0 = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], 64, 3, 0, order=0)
    e0.phase1(cfg["u"])
    e0.forward(torch.as_tensor(cfg["x0"]).to(e0.device))
    rc = lib.qoc_shard_affine_device(e0.cache.handle, C.c_void_p(c.data_ptr()), None, None)
    assert rc == _lib.ERR_INVALID
    e0.cache.close()


@pytest.mark.parametrize("d,nt,nc,m,order,P,pen,streamed", CASES)
@pytest.mark.parametrize("threads", ["1", "0"])
def test_in_library_time_sharded_running_penalty(d, nt, nc, m, order, P, pen, streamed, threads, monkeypatch):
    """qoc_create_sharded(QOC_SHARD_TIME) with a penalty: the second exchange (c_p records) inside the library, rank threads or
    one enqueueing thread; on one GPU the ranks are virtual (event hand-over), with >= P GPUs each rank has its own device
    (peer stores + flags)."""
    import torch
    if streamed:
        monkeypatch.setenv("QOC_STREAM_JAC", "1")
    monkeypatch.setenv("QOC_SHARD_THREADS", threads)
    cfg = o.config_synthetic(d, nt, nc=nc, m=m, seed=11)
    Jo, go, _ = o.evaluate(cfg, order=order, penalty=pen)
    ngpu = torch.cuda.device_count()
    cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == o.COST_INFIDELITY else q.setup_infidelity_abs_trace(cfg["T"])
    sh = sharding.InProcessSharded(cfg["A0"], cfg["A"], cfg["x0"], cost[1], cfg["u"].shape, [r % ngpu for r in range(P)], kind="time",
                                   dUkdp_order=order, penalty=q.setup_state_penalty(*pen))
    for _ in range(3):   # repeated evaluations: the epochs of the two exchanges keep advancing
        J, g = sh.evaluate(cfg["u"])
        assert abs(J - Jo) <= TOL_J * max(1.0, abs(Jo))
        assert np.abs(g - go).max() <= TOL_G * np.abs(go).max()
    sh.close()
