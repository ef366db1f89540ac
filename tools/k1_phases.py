"""Developer aid: cycles between consecutive compute-warp barriers of K1's CTA 0 (one line per barrier of a slice period)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qoc_b200 as q
from qoc_b200 import configs, _lib
name = sys.argv[2] if len(sys.argv) > 2 else "bus"
if name == "zz_batch":
    cfg = configs.config_zz_batch(4096); u = cfg["u_batch"]; batch = 4096
elif name == "cavity":
    cfg = configs.config_cavity(12, Nt=550); u = cfg["u"]; batch = 1
else:
    cfg = configs.config_bus(Nt=10000); u = cfg["u"]; batch = 1
cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], u.shape[-2:], batch=batch, dUkdp_order=0, store_costates=False)
cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == 0 else q.setup_infidelity_abs_trace(cfg["T"])
q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], cost[1], dUkdp_order=0)
lib = _lib.load()
n = 8
out = np.zeros(16 * n + 4096, dtype=np.int64)
fn = lib.qoc_debug_k1_timeline
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 5   # 1 = jac, +4 = barrier stamps, +2 = skip inverse
assert fn(cache.handle, out.ctypes.data, n, flags) == 0
st = out[16 * n:]
st = st[st != 0]
dt = np.diff(st)
# find the period: autocorrelation on the sequence of deltas
best = None
for P in range(20, 80):
    if len(dt) > 4 * P:
        a, b = dt[P:3 * P], dt[2 * P:4 * P]
        err = np.abs(a - b).sum() / a.sum()
        if best is None or err < best[0]: best = (err, P)
P = best[1]
print("barriers per slice period:", P, "mismatch", round(best[0], 3))
seg = dt[2 * P:3 * P]
seg2 = dt[3 * P:4 * P]
print("period cycles:", seg.sum(), seg2.sum())
print(" ".join(f"{int(x)}" for x in seg))
