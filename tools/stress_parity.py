"""Randomised parity sweep (developer aid, GPU): many small random problems through every K1 form against the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qoc_oracle as o
import qoc_b200 as q

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 150
# regime: "all" d = 2..28 (every K1 form) | "small" d = 2..9, nc <= 4, m <= d incl. column chunks (k1s / k3s) | "gpath" d = 29..48
regime = sys.argv[3] if len(sys.argv) > 3 else "all"
worst = dict(J=0.0, g=0.0, U=0.0)
fails = 0
t0 = time.time()
for it in range(n):
    d = int(rng.integers(2, 29)); nc = int(rng.integers(1, 4)); m = int(rng.integers(1, min(d, 5) + 1)); nt = int(rng.integers(1, 40))
    if regime == "small":
        d = int(rng.integers(2, 10)); nc = int(rng.integers(1, 5)); m = int(rng.integers(1, d + 1)); nt = int(rng.integers(1, 60))
    elif regime == "gpath":
        d = int(rng.integers(29, 49)); nc = int(rng.integers(1, 3)); m = int(rng.integers(1, 13)); nt = int(rng.integers(4, 30))
    pen = None
    if m <= 8 and rng.random() < 0.3:   # running state penalty on random rows / columns
        rows = sorted(set(int(r) for r in rng.integers(0, d, size=int(rng.integers(1, 4)))))
        cols = sorted(set(int(c) for c in rng.integers(0, m, size=int(rng.integers(1, 3)))))
        pen = (rows, cols, float(rng.uniform(0.05, 0.6)))
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
    Jo, go, co = o.evaluate(cfg, order=order, penalty=pen)
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order)
    try:
        J, g = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], q.setup_infidelity(cfg["T"], cfg["n"])[1], dUkdp_order=order,
                          penalty=None if pen is None else q.setup_state_penalty(*pen))
    except q.QOCError as e:
        # a problem that overflows in the oracle too (non-normal generators at large norm: J = -inf) is the library's NOT_FINITE
        # status doing its job, not a disagreement
        if not np.isfinite(Jo) or not np.all(np.isfinite(go)):
            print("both non-finite (oracle J", Jo, "), library:", str(e))
            cache.close()
            continue
        fails += 1
        print("ERROR", dict(d=d, nc=nc, m=m, nt=nt, order=order, kind=kind, scale=scale, pen=pen), str(e), "| oracle J", Jo,
              "finite oracle gradient:", bool(np.all(np.isfinite(go))))
        cache.close()
        continue
    eJ = abs(J - Jo) / max(1.0, abs(Jo)); eg = np.abs(g - go).max() / max(np.abs(go).max(), 1e-300)
    eU = np.abs(cache.Uk_vec - co["Uk"]).max() / max(1.0, np.abs(co["Uk"]).max())
    worst["J"] = max(worst["J"], eJ); worst["g"] = max(worst["g"], eg); worst["U"] = max(worst["U"], eU)
    if eJ > 1e-10 or eg > 1e-8 or eU > 1e-11:
        fails += 1
        print("FAIL", dict(d=d, nc=nc, m=m, nt=nt, order=order, kind=kind, scale=scale, pen=pen), "eJ %.1e eg %.1e eU %.1e" % (eJ, eg, eU))
        if pen is not None:   # where does J differ: the terminal cost or the running sum?
            xs = cache.x
            L = q.setup_state_penalty(*pen)[0]
            Jc = q.setup_infidelity(cfg["T"], cfg["n"])[0](xs[-1])
            per = [L(xk) for xk in xs]
            pero = [o.setup_state_penalty(*pen)[0](xk) for xk in co["x"]]
            print("   J gpu %.12f oracle %.12f | host cost %.12f + host sum L %.12f = %.12f | oracle sum L %.12f | states max err %.1e" %
                  (J, Jo, Jc, sum(per), Jc + sum(per), sum(pero), np.abs(xs - co["x"]).max()))
    cache.close()
print("regime", regime, "cases", n, "fails", fails, "worst", {k: "%.1e" % v for k, v in worst.items()}, "time %.0fs" % (time.time() - t0))
