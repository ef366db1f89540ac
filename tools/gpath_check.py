"""Developer check of the general (d > 28) path on a B200: forced on small d against the oracle, then native sizes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qoc_oracle as o
import qoc_b200 as q

def run(cfg, order, name, pen=None):
    Jo, go, co = o.evaluate(cfg, order=order, penalty=pen)
    cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == o.COST_INFIDELITY else q.setup_infidelity_abs_trace(cfg["T"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    t = time.time()
    Jg, gg = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost[1], dUkdp_order=order,
                        penalty=None if pen is None else q.setup_state_penalty(*pen))
    tg = time.time() - t
    U = cache.Uk_vec; X = cache.x; LAM = cache.lam
    print(f"{name} order={order}: |dJ|={abs(Jg-Jo):.2e} maxU={np.abs(U-co['Uk']).max():.2e} maxX={np.abs(X-co['x']).max():.2e} "
          f"maxLam={np.abs(LAM-co['lam']).max():.2e} grad rel={np.abs(gg-go).max()/np.abs(go).max():.2e} launches={cache.launch_count()} t={tg:.3f}s", flush=True)

if __name__ == "__main__":
    if os.environ.get("QOC_FORCE_GPATH") == "1":
        run(o.config_zz(), 0, "zz(forced)")
        run(o.config_zz(), 3, "zz(forced)")
        run(o.config_zz(), 4, "zz(forced)", pen=([6, 7, 8], [0, 1, 2, 3], 0.22))
        run(o.config_bus(Nt=200, tgate=7.0), 0, "bus200(forced)")
        run(o.config_cavity(12, Nt=100), 3, "cavity12(forced)")
        s = o.config_synthetic(12, 24, nc=2, m=3, seed=5); s["A0"] = s["A0"] * 11; s["A"] = [a * 11 for a in s["A"]]
        run(s, 0, "synth12 x11 (forced, squarings)")
        run(s, 3, "synth12 x11 (forced, squarings)")
    else:
        run(o.config_synthetic(32, 20), 0, "synth32")
        run(o.config_cavity(20, Nt=60), 3, "cavity20 d=40")
        run(o.config_cavity(20, Nt=60), 0, "cavity20 d=40")
        run(o.config_synthetic(64, 12), 0, "synth64")
        run(o.config_cavity(40, Nt=24), 0, "cavity40 d=80")
        run(o.config_synthetic(128, 6, nc=1, m=2), 0, "synth128")
