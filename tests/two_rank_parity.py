"""Launched by tests/test_gpu_fullsize.py::test_time_sharding_on_two_real_ranks under torchrun (one process per GPU, NCCL):
ONE pulse time-segment sharded over the ranks against (a) the single-GPU evaluation and (b) the C restatement of the
reference, on the on-chip path (bus d = 27) and the general path (synthetic d = 32), exact-Frechet and Taylor-3."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import qoc_oracle as o  # noqa: E402
import qoc_ref  # noqa: E402
import qoc_b200 as q  # noqa: E402
from qoc_b200 import sharding  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    cases = [(o.config_bus(Nt=2000, tgate=70.0), 0), (o.config_bus(Nt=2001, tgate=70.035), 3),
             (o.config_synthetic(32, 403), 0), (o.config_synthetic(16, 5000), 0), (o.config_cavity(12, Nt=550), 3)]
    for cfg, order in cases:
        u = cfg["u"]
        nc, nt = u.shape
        lo, hi = sharding.time_partition(nt, world, rank)
        cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == o.COST_INFIDELITY else q.setup_infidelity_abs_trace(cfg["T"])
        eng = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], hi - lo, cfg["x0"].shape[1], lr, order=order)
        ev = sharding.TimeShardedEvaluator(eng, cfg["x0"], cost[1], nt)
        J, g = ev.evaluate(u)
        cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (nc, nt), device=lr, dUkdp_order=order, store_costates=False)
        J1, g1 = q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], cost[1], dUkdp_order=order)
        r = qoc_ref.ref_eval(cfg, order=order, nthreads=4)
        for name, Jr, gr in (("single GPU", J1, g1), ("C restatement", r["J"], r["dJdu"])):
            dJ, dg = abs(J - Jr), np.abs(g - gr).max() / np.abs(gr).max()
            print(f"rank {rank} {cfg['name']} nt={nt} order={order} vs {name}: |dJ|={dJ:.2e} rel|dg|={dg:.2e}", flush=True)
            assert dJ <= 1e-10 * max(1.0, abs(Jr)) and dg <= 1e-8
        cache.close()
    # running state penalty: the second exchange (affine terms c_p) over NCCL, on-chip path and general path
    for cfg, order, pen in ((o.config_synthetic(16, 1001, nc=2, m=3, seed=5), 0, ([9, 12, 15], [0, 2], 0.37)),
                            (o.config_synthetic(32, 403), 3, ([1, 30, 31], [0, 3], 0.9)),
                            (o.config_synthetic(6, 999, nc=2, m=4, seed=2), 0, ([4, 5], [0, 1, 2], 1.5))):
        u = cfg["u"]
        nc, nt = u.shape
        lo, hi = sharding.time_partition(nt, world, rank)
        L = q.setup_state_penalty(*pen)
        cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == o.COST_INFIDELITY else q.setup_infidelity_abs_trace(cfg["T"])
        eng = sharding.CudaSegmentEngine(cfg["A0"], cfg["A"], hi - lo, cfg["x0"].shape[1], lr, order=order, penalty=L)
        ev = sharding.TimeShardedEvaluator(eng, cfg["x0"], cost[1], nt)
        J, g = ev.evaluate(u)
        Jo, go, _ = o.evaluate(cfg, order=order, penalty=pen)
        dJ, dg = abs(J - Jo), np.abs(g - go).max() / np.abs(go).max()
        print(f"rank {rank} {cfg['name']} nt={nt} order={order} running penalty vs oracle: |dJ|={dJ:.2e} rel|dg|={dg:.2e}", flush=True)
        assert dJ <= 1e-10 * max(1.0, abs(Jo)) and dg <= 1e-8
        eng.cache.close()
    dist.barrier()
    if rank == 0:
        print("TWO_RANK_PARITY_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
