"""Short profiling driver: a few device-resident evaluations of a workload (for ncu; never a bench number)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qoc_b200 as q
from qoc_b200 import configs, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "bus"
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
order = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if name == "bus":
    cfg = configs.config_bus(Nt=nt, tgate=350.0 * nt / 10000); u = cfg["u"]; batch = 1
elif name == "zz_batch":
    cfg = configs.config_zz_batch(nt); u = cfg["u_batch"]; batch = nt
else:
    cfg = configs.config_cavity(12, Nt=nt); u = cfg["u"]; batch = 1
cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], u.shape[-2:], batch=batch, dUkdp_order=order, store_costates=False)
cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == 0 else q.setup_infidelity_abs_trace(cfg["T"])
for _ in range(reps):
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], cost[1], dUkdp_order=order)
print("J", J if batch == 1 else J[:3])
