"""Host-side multi-GPU logic under gloo with world_size 2 on CPU: the partitioning, the boundary algebra and the two
exchanges of quantumoptimalcontrol.jl_b200/sharding.py, driven by an ORACLE-backed stand-in engine (tests may use the
oracle; the product engine is CudaSegmentEngine and is exercised by the gpu tests / bench on >= 2 GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import qoc_oracle as o
import qoc_b200 as q
from qoc_b200 import sharding


class OracleSegmentEngine:
    """phase1 / forward / backward of one time segment computed by the numpy oracle (test double)."""
    device = torch.device("cpu")

    def __init__(self, A0, A, order=0, penalty=None):
        self.A0, self.A, self.order = A0, A, order
        self.penalty = penalty[0] if isinstance(penalty, tuple) else penalty   # q.setup_state_penalty(...)[0]: rows, cols, mu
        if self.penalty is not None:
            self.L, self.dL = o.setup_state_penalty(self.penalty.rows, self.penalty.cols, self.penalty.mu)

    def phase1(self, u_local):
        self.u = np.asarray(u_local)
        self.Uk = [o.expm_higham2005(o.generator(self.A0, self.A, self.u[:, k])) for k in range(self.u.shape[1])]
        S = np.eye(self.A0.shape[0], dtype=complex)
        for U in self.Uk:
            S = U @ S
        return torch.as_tensor(S)

    def forward(self, x_start):
        x = x_start.numpy()
        self.x = [x]
        for U in self.Uk:
            self.x.append(U @ self.x[-1])
        return torch.as_tensor(self.x[-1])

    def affine(self):
        """c_p = lambda_start for lambda_end = 0 (the recurrence of src/gradient_computations.jl:47-57 on the local states),
        and the local sum of L over the nt_local + 1 local states."""
        lam = self.dL(self.x[-1])
        for k in range(len(self.Uk) - 1, -1, -1):
            lam = self.Uk[k].conj().T @ lam + self.dL(self.x[k])
        return torch.as_tensor(lam), torch.as_tensor([sum(self.L(xk) for xk in self.x)], dtype=torch.float64)

    def backward(self, lam_end):
        lam = lam_end.numpy()
        nt = len(self.Uk)
        g = np.zeros((len(self.A), nt))
        if self.penalty is not None:   # as the reference does for lambda_N (:47-49)
            lam = lam + self.dL(self.x[-1])
        for k in range(nt - 1, -1, -1):
            X = o.generator(self.A0, self.A, self.u[:, k])
            dU = [o.expm_frechet_sps(X, Aj)[1] for Aj in self.A] if self.order == 0 else \
                o.expm_jacobian(self.A0, self.A, self.u[:, k], self.order)
            for j in range(len(self.A)):
                g[j, k] = o.compute_u_sensitivity(self.x[k], lam, dU[j])
            lam = self.Uk[k].conj().T @ lam
            if self.penalty is not None:
                lam = lam + self.dL(self.x[k])
        return torch.as_tensor(g), torch.as_tensor(lam)


PEN = ([5, 6, 7], [0, 2], 0.37)   # rows, columns, mu of the running penalty in the time_pen case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q_out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if mode == "time":
            cfg = o.config_synthetic(8, 37, nc=2, m=3, seed=4)      # 37 slices: ranks own 18 and 19
            for cost, order in ((q.setup_infidelity(cfg["T"], cfg["n"])[1], 0),
                                (o.setup_infidelity_zcalibrated(np.eye(8, 4, dtype=complex)) if False else
                                 o.setup_infidelity(cfg["T"], cfg["n"]), 3)):
                ev = sharding.TimeShardedEvaluator(OracleSegmentEngine(cfg["A0"], cfg["A"], order), cfg["x0"], cost, 37)
                J, g = ev.evaluate(cfg["u"])
                q_out.put((rank, "time", order, J, g, (ev.lo, ev.hi)))
        elif mode == "time_pen":
            cfg = o.config_synthetic(8, 37, nc=2, m=3, seed=4)
            for order in (0, 3):
                pen = q.setup_state_penalty(*PEN)
                ev = sharding.TimeShardedEvaluator(OracleSegmentEngine(cfg["A0"], cfg["A"], order, pen), cfg["x0"],
                                                   o.setup_infidelity(cfg["T"], cfg["n"]), 37)
                J, g = ev.evaluate(cfg["u"])
                q_out.put((rank, "time_pen", order, J, g, (ev.lo, ev.hi)))
        else:
            cfg = o.config_zz_batch(5)
            ub = cfg["u_batch"]

            def eval_local(u_block):
                Js, gs = [], []
                for u in u_block:
                    J, g, _ = o.evaluate(cfg, order=3, u=u)
                    Js.append(J)
                    gs.append(g)
                return np.array(Js), np.array(gs).reshape((len(Js),) + ub.shape[1:])

            J, g, blk = sharding.evaluate_batch_sharded(eval_local, ub)
            q_out.put((rank, "batch", 3, J, g, blk))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(mode, world=2):
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    for _ in range(world * (2 if mode in ("time", "time_pen") else 1)):
        res.append(q_out.get(timeout=180))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_partitions():
    assert [sharding.time_partition(10000, 8, r) for r in range(8)][3] == (3750, 5000)
    parts = [sharding.block_partition(37, 2, r) for r in range(2)]
    assert parts == [(0, 18), (18, 37)]
    for n, w in ((4096, 8), (5, 2), (3, 4), (1, 2)):
        cover = [sharding.block_partition(n, w, r) for r in range(w)]
        assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))


def test_time_sharded_world2_gloo():
    cfg = o.config_synthetic(8, 37, nc=2, m=3, seed=4)
    res = _run("time")
    for order in (0, 3):
        Jo, go, _ = o.evaluate(cfg, order=order)
        got = [r for r in res if r[2] == order]
        assert sorted(r[0] for r in got) == [0, 1]
        assert sorted(r[5] for r in got) == [(0, 18), (18, 37)]
        for r in got:  # every rank holds the full, identical answer
            assert abs(r[3] - Jo) < 1e-12
            assert np.abs(r[4] - go).max() < 1e-11 * max(1.0, np.abs(go).max())


@pytest.mark.parametrize("world", [2, 3])
def test_time_sharded_running_penalty_gloo(world):
    """The second exchange (affine terms c_p, partial sums of L): src/gradient_computations.jl:47-57 split over the ranks."""
    cfg = o.config_synthetic(8, 37, nc=2, m=3, seed=4)
    res = _run("time_pen", world)
    for order in (0, 3):
        Jo, go, _ = o.evaluate(cfg, order=order, penalty=PEN)
        J0, g0, _ = o.evaluate(cfg, order=order)
        assert abs(Jo - J0) > 1e-3 and np.abs(go - g0).max() > 1e-4   # the penalty matters in this case
        got = [r for r in res if r[2] == order]
        assert sorted(r[0] for r in got) == list(range(world))
        for r in got:
            assert abs(r[3] - Jo) < 1e-12
            assert np.abs(r[4] - go).max() < 1e-11 * max(1.0, np.abs(go).max())


def test_batch_sharded_world2_gloo():
    cfg = o.config_zz_batch(5)
    res = _run("batch")
    assert sorted(r[5] for r in res) == [(0, 2), (2, 5)]
    for r in res:
        for b in range(5):
            Jo, go, _ = o.evaluate(cfg, order=3, u=cfg["u_batch"][b])
            assert abs(r[3][b] - Jo) < 1e-13 and np.abs(r[4][b] - go).max() < 1e-13


def test_single_process_fallbacks():
    """world_size 1 (no process group): same code path without collectives."""
    cfg = o.config_synthetic(6, 11, nc=1, m=2, seed=2)
    ev = sharding.TimeShardedEvaluator(OracleSegmentEngine(cfg["A0"], cfg["A"], 0), cfg["x0"],
                                       o.setup_infidelity(cfg["T"], cfg["n"]), 11)
    J, g = ev.evaluate(cfg["u"])
    Jo, go, _ = o.evaluate(cfg, order=0)
    assert abs(J - Jo) < 1e-13 and np.abs(g - go).max() < 1e-12
