"""Developer smoke check on a B200: CUDA path vs oracle on small cases (not a test; tests/ has the real ones)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qoc_oracle as o
import qoc_b200 as q

def run(cfg, order, name):
    J, dJ = o.cost_closures(cfg)
    t = time.time(); Jo, go, co = o.evaluate(cfg, order=order); to = time.time() - t
    cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == o.COST_INFIDELITY else q.setup_infidelity_abs_trace(cfg["T"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    t = time.time(); Jg, gg = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost[1], dUkdp_order=order); tg = time.time() - t
    U = cache.Uk_vec; dU = cache.dUkdu; X = cache.x; LAM = cache.lam
    eU = np.abs(U - co["Uk"]).max()
    eX = np.abs(X - co["x"]).max(); eL = np.abs(LAM - co["lam"]).max()
    print(f"{name} order={order}: J gpu={Jg:.15f} oracle={Jo:.15f} |dJ|={abs(Jg-Jo):.2e} maxU={eU:.2e} maxX={eX:.2e} maxLam={eL:.2e} "
          f"grad rel={np.abs(gg-go).max()/np.abs(go).max():.2e} launches={cache.launch_count()} t_oracle={to:.2f}s t_gpu={tg:.3f}s flops={cache.alg_flops():.3e}")

if __name__ == "__main__":
    run(o.config_zz(), 3, "zz")
    run(o.config_zz(), 0, "zz")
    run(o.config_zz(), 4, "zz")
    run(o.config_cavity(12, Nt=100), 3, "cavity12")
    run(o.config_cavity(12, Nt=550), 0, "cavity12")
    run(o.config_bus(Nt=400, tgate=14.0), 0, "bus400")
    run(o.config_synthetic(16, 300), 0, "synth16")
    run(o.config_synthetic(20, 64, nc=1, m=1), 2, "synth20")
