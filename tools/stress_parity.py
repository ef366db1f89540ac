"""Randomised parity sweep (developer aid, GPU): many small random problems through every K1 form against the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qoc_oracle as o
import qoc_b200 as q

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 150
worst = dict(J=0.0, g=0.0, U=0.0)
fails = 0
t0 = time.time()
for it in range(n):
    d = int(rng.integers(2, 29)); nc = int(rng.integers(1, 4)); m = int(rng.integers(1, min(d, 5) + 1)); nt = int(rng.integers(1, 40))
    order = int(rng.integers(0, 5)); kind = rng.choice(["complex", "real_sym", "real_nonsym"]); scale = float(rng.choice([0.05, 0.3, 1.0, 3.0, 8.0, 20.0]))
    def H():
        if kind == "complex":
            A = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)); return (A + A.conj().T) / 2
        A = rng.standard_normal((d, d)); return (A + A.T) / 2 if kind == "real_sym" else A
    H0 = H(); H0 = H0 * (scale / np.abs(H0).sum(axis=0).max())
    A = [(-1j * Hj / np.abs(Hj).sum(axis=0).max()) for Hj in (H() for _ in range(nc))]
    Tq, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    cfg = dict(A0=(-1j * H0).astype(complex), A=[a.astype(complex) for a in A], u=rng.uniform(-0.5, 0.5, (nc, nt)),
               x0=np.eye(d, m, dtype=complex), T=Tq[:, :m].copy(), cost=o.COST_INFIDELITY, n=m)
    Jo, go, co = o.evaluate(cfg, order=order)
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], dUkdp_order=order)
    eJ = abs(J - Jo) / max(1.0, abs(Jo)); eg = np.abs(g - go).max() / max(np.abs(go).max(), 1e-300)
    eU = np.abs(cache.Uk_vec - co["Uk"]).max() / max(1.0, np.abs(co["Uk"]).max())
    worst["J"] = max(worst["J"], eJ); worst["g"] = max(worst["g"], eg); worst["U"] = max(worst["U"], eU)
    if eJ > 1e-10 or eg > 1e-8 or eU > 1e-11:
        fails += 1
        print("FAIL", dict(d=d, nc=nc, m=m, nt=nt, order=order, kind=kind, scale=scale), "eJ %.1e eg %.1e eU %.1e" % (eJ, eg, eU))
print("cases", n, "fails", fails, "worst", {k: "%.1e" % v for k, v in worst.items()}, "time %.0fs" % (time.time() - t0))
