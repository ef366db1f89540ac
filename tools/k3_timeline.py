"""Developer aid: per-step timeline of K3N's recurrence warp 0 in CTA 0 (cycles): wait for the operand, barrier, compute."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qoc_b200 as q
from qoc_b200 import configs, _lib
cfg = configs.config_bus(Nt=10000); u = cfg["u"]
cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], u.shape[-2:], batch=1, dUkdp_order=0, store_costates=False)
q.evaluate(cache, cfg["A0"], cfg["A"], u, cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], dUkdp_order=0)
lib = _lib.load()
n = 134
out = np.zeros(4 * n, dtype=np.int64)
fn = lib.qoc_debug_k3_timeline
fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert fn(cache.handle, out.ctypes.data, n) == 0
o = out[out != 0]
dt = np.diff(o)
print("backward steps stamped", len(o), "median cycles/step", np.median(dt), "min", dt.min(), "max", dt.max())
