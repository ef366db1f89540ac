"""One (or a few) fidelity+gradient evaluations of a bench workload through the C ABI: the command profiled under ncu.
usage: python tools/run_once.py <workload> [frechet|taylor3] [reps]     (workload names as in bench.py --workload)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import qoc_b200 as q  # noqa: E402


def main():
    wl = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "frechet"
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    order = 0 if mode == "frechet" else 3
    cfg, u, batch, desc = bench.build_workload(wl, 0, mode)
    nc, nt = u.shape[-2], u.shape[-1]
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], (nc, nt), batch=batch, dUkdp_order=order, store_costates=False)
    cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == 0 else q.setup_infidelity_abs_trace(cfg["T"])
    for r in range(reps):
        t0 = time.perf_counter()
        J, g = q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], cost[1], dUkdp_order=order)
        dt = time.perf_counter() - t0
        print(f"{desc} [{mode}] rep {r}: J={np.atleast_1d(J)[0]:.15f} |g|max={np.abs(g).max():.6e} launches={cache.launch_count()} "
              f"wall={1e3 * dt:.3f} ms alg_flops={cache.alg_flops():.4e}", flush=True)


if __name__ == "__main__":
    main()
