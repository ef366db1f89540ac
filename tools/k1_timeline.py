"""Developer aid: prints the per-slice hand-off timeline of K1's CTA 0 (cycles), compute warps vs service warp."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qoc_b200 as q
from qoc_b200 import configs, _lib
order = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cfg = configs.config_bus(Nt=10000)
cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order, store_costates=False)
J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], dUkdp_order=order)
lib = _lib.load()
n = 12
out = np.zeros((n, 16), dtype=np.int64)
fn = lib.qoc_debug_k1_timeline
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
want_jac = int(sys.argv[2]) if len(sys.argv) > 2 else 1
assert fn(cache.handle, out.ctypes.data, n, want_jac) == 0
names = {0: "c:start", 1: "c:part1 done", 2: "c:next N built", 3: "c:reached Ninv wait", 4: "c:tail done", 5: "c:Q done",
         8: "s:loop top", 9: "s:reached N wait", 10: "s:inv done"}
d = out[3:n - 1]
print("period (cycles):", np.diff(d[:, 0]))
print("compute: part1", np.median(d[:, 1] - d[:, 0]), "buildX+N(next)", np.median(d[:, 2] - d[:, 1]),
      "tail incl. wait", np.median(d[:, 4] - d[:, 2]), "Q", np.median(d[:, 5] - d[:, 4]))
print("service: inverse incl. wait for N", np.median(d[:, 10] - d[:, 9]), " loop period", np.median(np.diff(d[:, 10])))
# when did inverse(k) finish relative to compute reaching the Ninv(k) wait?  (>0: compute waited)
print("inverse finished - compute reached wait:", (d[:, 10] - d[:, 3]))
