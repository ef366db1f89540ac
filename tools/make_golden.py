"""Generates tests/golden/*.npz from the oracle (numpy restatement of the reference; the Julia reference itself
cannot run in this image).  Inputs are the BASELINE.json configs at sizes the oracle finishes in seconds.
Run: python tools/make_golden.py     (committed outputs are what the tests compare against)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qoc_oracle as o  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dump(name, cfg, order, penalty=None, keep_states=False):
    J, g, cache = o.evaluate(cfg, order=order, penalty=penalty)
    d = dict(J=J, dJdu=g, x_final=cache["x"][-1], u=cfg["u"], order=order)
    if keep_states:
        d.update(Uk=cache["Uk"], x=cache["x"], lam=cache["lam"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "J =", repr(J), "max|g| =", np.abs(g).max())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    zz = o.config_zz()
    for od in (1, 2, 3, 4, 0):
        dump(f"zz_order{od}", zz, od, keep_states=(od == 3))
    dump("zz_penalty_order4", zz, 4, penalty=([6, 7, 8], [0, 1, 2, 3], 0.22))
    cav = o.config_cavity(12, Nt=100)   # test/test_gradient_computation.jl:27-35
    dump("cavity12_nt100_order3", cav, 3, keep_states=True)
    dump("cavity12_nt550_order0", o.config_cavity(12, Nt=550), 0)
    dump("bus_nt500_order0", o.config_bus(Nt=500, tgate=17.5), 0)
    dump("synth16_nt64_order0", o.config_synthetic(16, 64), 0)
    # the reference's two example known answers, at full size
    c = o.config_cavity(12, Nt=550)
    H0, Tc, x0, theta = o.model_cavity_qubit(12)
    cache = o.propagate(c["A0"], c["A"], c["u"], x0.astype(complex))
    tgt = np.kron([1, 0], np.exp(1j * theta))
    tgt = tgt / np.linalg.norm(tgt)
    ov = abs(np.vdot(tgt, cache["x"][-1][:, 0]))
    b = o.config_bus(Nt=10000)
    cb = o.propagate(b["A0"], b["A"], b["u"], b["x0"])
    pop = abs(np.vdot(b["T"], cb["x"][-1])) ** 2
    np.savez(os.path.join(OUT, "known_answers.npz"), cavity_overlap=ov, bus_population=pop,
             bus_x_final=cb["x"][-1])
    print("cavity overlap", repr(ov), "bus population", repr(pop))
