"""Developer aid: per-stage device times (K1 / K2 / K3) of the small single-pulse workloads, default sweeps vs QOC_OLD_SWEEPS=1
(first-generation, non-cooperative boundary scan).  usage: python tools/k2_floor.py <workload> [frechet|taylor3]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qoc_b200 as q

wl = sys.argv[1]; mode = sys.argv[2] if len(sys.argv) > 2 else "frechet"
order = 0 if mode == "frechet" else 3
cfg, u, batch, desc = bench.build_workload(wl, 0, mode)
nc, nt = u.shape[-2], u.shape[-1]
cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (nc, nt), batch=batch, dUkdp_order=order, store_costates=False)
cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == 0 else q.setup_infidelity_abs_trace(cfg["T"])
for r in range(5):
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], cost[1], dUkdp_order=order)
t0 = time.perf_counter()
for r in range(200):
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], cost[1], dUkdp_order=order)
wall = (time.perf_counter() - t0) / 200
cache.set_profiling(True)
st = np.zeros(3)
for r in range(20):
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], cost[1], dUkdp_order=order)
    st += np.array(cache.stage_ms())
print(desc, mode, "OLD" if os.environ.get("QOC_OLD_SWEEPS") == "1" else "new", "e2e wall %.1f us" % (1e6 * wall),
      "k1 %.1f k2 %.1f k3 %.1f us" % tuple(1e3 * st / 20), "launches", cache.launch_count())
