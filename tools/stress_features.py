"""Randomised parity sweep over the round-2 features (developer aid, GPU): z-calibrated device cost, in-library time sharding
over virtual ranks, streamed Jacobians, batches with column chunks, control bounds.  usage: stress_features.py [seed] [n]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qoc_oracle as o
import qoc_b200 as q
from qoc_b200 import sharding

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
FEATS = sys.argv[3].split(",") if len(sys.argv) > 3 else ["zcal", "shard", "stream", "batch", "shardpen", "chunkpen"]
fails = 0
worst = {}
t0 = time.time()


def problem(d, nc, m, nt, scale):
    def H():
        A = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)); return (A + A.conj().T) / 2
    H0 = H(); H0 = H0 * (scale / np.abs(H0).sum(axis=0).max())
    A = [(-1j * Hj / np.abs(Hj).sum(axis=0).max()) for Hj in (H() for _ in range(nc))]
    Tq, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
    return dict(A0=(-1j * H0).astype(complex), A=[a.astype(complex) for a in A], u=rng.uniform(-0.5, 0.5, (nc, nt)),
                x0=np.eye(d, m, dtype=complex), T=Tq[:, :m].copy(), cost=o.COST_INFIDELITY, n=m)


def check(tag, info, J, g, Jo, go, tol_g=1e-8):
    global fails
    eJ = abs(J - Jo) / max(1.0, abs(Jo)); eg = np.abs(g - go).max() / max(np.abs(go).max(), 1e-300)
    w = worst.setdefault(tag, [0.0, 0.0]); w[0] = max(w[0], eJ); w[1] = max(w[1], eg)
    if eJ > 1e-10 or eg > tol_g:
        fails += 1
        print("FAIL", tag, info, "eJ %.1e eg %.1e" % (eJ, eg))


for it in range(n):
    feat = rng.choice(FEATS)
    order = int(rng.choice([0, 3]))
    scale = float(rng.choice([0.05, 0.3, 1.0, 3.0]))
    os.environ.pop("QOC_STREAM_JAC", None)
    os.environ.pop("QOC_SHARD_THREADS", None)
    if feat == "zcal":
        d = int(rng.integers(4, 41)); nc = int(rng.integers(1, 3)); nt = int(rng.integers(2, 30))
        cfg = problem(d, nc, 4, nt, scale)
        Jo_, dJo = o.setup_infidelity_zcalibrated(cfg["T"])
        co = o.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape)
        xs = o.propagate(cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], co)["x"]
        go = o.grape_sensitivity(cfg["A0"], cfg["A"], dJo, cfg["u"], cfg["x0"], co, dUkdp_order=order).copy()
        cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
        J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], q.setup_infidelity_zcalibrated(cfg["T"])[1], dUkdp_order=order)
        check(feat, dict(d=d, nc=nc, nt=nt, order=order, scale=scale), J, g, Jo_(xs[-1]), go, tol_g=1e-6)   # (golden-section theta: 1e-7, see tests)
        cache.close()
    elif feat == "shard":
        d = int(rng.integers(2, 41)); nc = int(rng.integers(1, 3)); m = int(rng.integers(1, min(d, 8) + 1)); P = int(rng.integers(2, 5))
        nt = int(rng.integers(4 * P, 120))
        cfg = problem(d, nc, m, nt, scale)
        Jo, go, _ = o.evaluate(cfg, order=order)
        sh = sharding.InProcessSharded(cfg["A0"], cfg["A"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], cfg["u"].shape, [0] * P,
                                       kind="time", dUkdp_order=order)
        J, g = sh.evaluate(cfg["u"])
        check(feat, dict(d=d, nc=nc, m=m, nt=nt, P=P, order=order, scale=scale), J, g, Jo, go)
        sh.close()
    elif feat == "shardpen":   # running penalty under in-library time sharding (second exchange), every sweep family
        d = int(rng.integers(2, 41)); nc = int(rng.integers(1, 3)); m = int(rng.integers(1, min(d, 8) + 1)); P = int(rng.integers(2, 5))
        nt = int(rng.integers(4 * P, 120))
        cfg = problem(d, nc, m, nt, scale)
        rows = sorted(rng.choice(d, size=int(rng.integers(1, d + 1)), replace=False).tolist())
        cols = sorted(rng.choice(m, size=int(rng.integers(1, m + 1)), replace=False).tolist())
        pen = (rows, cols, float(rng.uniform(0.1, 2.0)))
        Jo, go, _ = o.evaluate(cfg, order=order, penalty=pen)
        if rng.random() < 0.3:
            os.environ["QOC_STREAM_JAC"] = "1"
        os.environ["QOC_SHARD_THREADS"] = str(int(rng.integers(0, 2)))
        sh = sharding.InProcessSharded(cfg["A0"], cfg["A"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], cfg["u"].shape, [0] * P,
                                       kind="time", dUkdp_order=order, penalty=q.setup_state_penalty(*pen))
        for rep in range(2):
            J, g = sh.evaluate(cfg["u"])
            check(feat, dict(d=d, nc=nc, m=m, nt=nt, P=P, order=order, scale=scale, pen=pen, rep=rep), J, g, Jo, go)
        sh.close()
    elif feat == "chunkpen":   # running penalty with more than 8 state columns (column chunks), single pulses and batches
        d = int(rng.integers(9, 41)); nc = int(rng.integers(1, 3)); m = int(rng.integers(9, d + 1)); nt = int(rng.integers(2, 40))
        nb = int(rng.integers(1, 4))
        cfg = problem(d, nc, m, nt, scale)
        rows = sorted(rng.choice(d, size=int(rng.integers(1, d + 1)), replace=False).tolist())
        cols = sorted(rng.choice(m, size=int(rng.integers(1, m + 1)), replace=False).tolist())
        pen = (rows, cols, float(rng.uniform(0.1, 2.0)))
        ub = rng.uniform(-0.5, 0.5, (nb, nc, nt))
        cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, batch=nb, dUkdp_order=order, store_costates=False)
        Jb, gb = q.evaluate(cache, cfg["A0"], cfg["A"], ub if nb > 1 else ub[0], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1],
                            dUkdp_order=order, penalty=q.setup_state_penalty(*pen))
        Jb, gb = np.atleast_1d(Jb), (gb if nb > 1 else gb[None])
        for b in range(nb):
            Jo, go, _ = o.evaluate(cfg, order=order, u=ub[b], penalty=pen)
            check(feat, dict(d=d, nc=nc, m=m, nt=nt, nb=nb, b=b, order=order, scale=scale, pen=pen), Jb[b], gb[b], Jo, go)
        cache.close()
    elif feat == "stream":
        d = int(rng.integers(29, 41)); nc = int(rng.integers(1, 3)); m = int(rng.integers(1, 6)); nt = int(rng.integers(4, 40))
        cfg = problem(d, nc, m, nt, scale)
        Jo, go, _ = o.evaluate(cfg, order=order)
        os.environ["QOC_STREAM_JAC"] = "1"
        cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
        J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], dUkdp_order=order)
        check(feat, dict(d=d, nc=nc, m=m, nt=nt, order=order, scale=scale), J, g, Jo, go)
        cache.close()
    else:
        d = int(rng.integers(2, 13)); nc = int(rng.integers(1, 4)); m = int(rng.integers(1, d + 1)); nt = int(rng.integers(2, 40)); nb = int(rng.integers(2, 6))
        cfg = problem(d, nc, m, nt, scale)
        ub = rng.uniform(-0.5, 0.5, (nb, nc, nt))
        cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, batch=nb, dUkdp_order=order, store_costates=False)
        Jb, gb = q.evaluate(cache, cfg["A0"], cfg["A"], ub, cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], dUkdp_order=order)
        for b in range(nb):
            Jo, go, _ = o.evaluate(cfg, order=order, u=ub[b])
            check(feat, dict(d=d, nc=nc, m=m, nt=nt, nb=nb, b=b, order=order, scale=scale), Jb[b], gb[b], Jo, go)
        cache.close()
print("cases", n, "fails", fails, "worst (J, g) per feature", {k: ("%.1e" % v[0], "%.1e" % v[1]) for k, v in worst.items()}, "time %.0fs" % (time.time() - t0))
