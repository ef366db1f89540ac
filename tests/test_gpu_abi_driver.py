"""The C ABI without Python in the loop (`-m gpu`): tests/abi_driver.c is compiled with gcc against include/qoc_b200.h and the
in-tree shared library, and run on the zz_coupling golden case -- the reference's call sequence (setup_grape_cache ->
propagate -> grape_sensitivity with a host closure and with the built-in cost -> stale-u error) and the in-library
multi-GPU evaluation (qoc_create_sharded), on as many GPUs as the box has (virtual ranks on one GPU otherwise).
Also: the in-library sharded evaluation from Python against the single-GPU evaluation on the BASELINE shapes."""
import os
import subprocess

import numpy as np
import pytest

import qoc_oracle as o
import qoc_b200 as q
from qoc_b200 import _lib, sharding

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_case(tmp, cfg, order):
    d, m = cfg["x0"].shape
    nc, nt = cfg["u"].shape
    J, g, cache = o.evaluate(cfg, order=order)
    Jf, dJf = o.cost_closures(cfg)
    lam = dJf(cache["x"][-1])
    with open(os.path.join(tmp, "dims.txt"), "w") as fh:
        fh.write(f"{d} {m} {nc} {nt} {order} {cfg['cost']} {cfg['n']}\n")
    f = lambda a: np.asfortranarray(np.asarray(a, dtype=np.complex128)).tobytes(order="F")
    open(os.path.join(tmp, "A0.bin"), "wb").write(f(cfg["A0"]))
    open(os.path.join(tmp, "A.bin"), "wb").write(b"".join(f(a) for a in cfg["A"]))
    open(os.path.join(tmp, "x0.bin"), "wb").write(f(cfg["x0"]))
    open(os.path.join(tmp, "T.bin"), "wb").write(f(cfg["T"]))
    open(os.path.join(tmp, "lam.bin"), "wb").write(f(lam))
    open(os.path.join(tmp, "u.bin"), "wb").write(np.ascontiguousarray(cfg["u"].T).tobytes())
    open(os.path.join(tmp, "J.bin"), "wb").write(np.array([J]).tobytes())
    open(os.path.join(tmp, "g.bin"), "wb").write(np.ascontiguousarray(g.T).tobytes())


@pytest.mark.parametrize("case", ["zz_order3", "cavity12_order0", "synth32_order0"])
def test_c_driver_through_the_abi(tmp_path, case):
    import torch
    _lib.load()
    if case == "zz_order3":
        cfg, order = o.config_zz(), 3
    elif case == "cavity12_order0":
        cfg, order = o.config_cavity(12, Nt=120), 0
    else:
        cfg, order = o.config_synthetic(32, 48), 0
    _write_case(str(tmp_path), cfg, order)
    exe = str(tmp_path / "abi_driver")
    libdir = os.path.dirname(_lib.lib_path())
    subprocess.run(["gcc", "-O1", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_driver.c"), "-o", exe,
                    "-L", libdir, "-lqoc_b200", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    ngpu = torch.cuda.device_count()
    env = dict(os.environ)
    nranks = 3
    if ngpu < nranks:
        env["ABI_DRIVER_ONE_GPU"] = "1"
    r = subprocess.run([exe, str(tmp_path), str(nranks)], capture_output=True, text=True, env=env, timeout=600)
    print(r.stdout)
    print(r.stderr)
    assert r.returncode == 0 and "ABI_DRIVER_OK" in r.stdout


@pytest.mark.parametrize("name,order,nranks", [("bus", 0, 2), ("bus", 3, 4), ("synth32", 0, 3), ("synth16", 0, 4), ("cavity20", 3, 2)])
def test_in_library_time_sharding_matches_single_gpu(name, order, nranks):
    import torch
    cfg = {"bus": lambda: o.config_bus(Nt=3001, tgate=105.035), "synth32": lambda: o.config_synthetic(32, 603),
           "synth16": lambda: o.config_synthetic(16, 4000), "cavity20": lambda: o.config_cavity(20, Nt=550)}[name]()
    cost = q.setup_infidelity(cfg["T"], cfg["n"]) if cfg["cost"] == o.COST_INFIDELITY else q.setup_infidelity_abs_trace(cfg["T"])
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], cfg["u"].shape, dUkdp_order=order, store_costates=False)
    J1, g1 = q.evaluate(cache, cfg["A0"], cfg["A"], cfg["u"], cfg["x0"], cost[1], dUkdp_order=order)
    ngpu = torch.cuda.device_count()
    devs = [r % ngpu for r in range(nranks)]
    sh = sharding.InProcessSharded(cfg["A0"], cfg["A"], cfg["x0"], cost[1], cfg["u"].shape, devs, kind="time", dUkdp_order=order)
    for _ in range(2):   # the second evaluation reuses every buffer
        J, g = sh.evaluate(cfg["u"])
        assert abs(J - J1) <= 1e-10 * max(1.0, abs(J1))
        assert np.abs(g - g1).max() <= 1e-8 * np.abs(g1).max()
    assert sh.last_ms() > 0
    sh.close(); cache.close()


def test_in_library_batch_sharding_matches_single_gpu():
    import torch
    nb = 257
    cfg = o.config_zz_batch(nb)
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    ub = cfg["u_batch"]
    cache = q.setup_grape_cache(cfg["A0"], cfg["x0"], ub.shape[-2:], batch=nb, dUkdp_order=3, store_costates=False)
    J1, g1 = q.evaluate(cache, cfg["A0"], cfg["A"], ub, cfg["x0"], cost[1], dUkdp_order=3)
    ngpu = torch.cuda.device_count()
    sh = sharding.InProcessSharded(cfg["A0"], cfg["A"], cfg["x0"], cost[1], ub.shape[-2:], [r % ngpu for r in range(3)], kind="batch",
                                   batch=nb, dUkdp_order=3)
    J, g = sh.evaluate(ub)
    assert np.abs(J - J1).max() <= 1e-12 and np.abs(g - g1).max() <= 1e-12 * max(1.0, np.abs(g1).max())
    sh.close(); cache.close()


def test_sharded_create_errors():
    cfg = o.config_zz()
    cost = q.setup_infidelity(cfg["T"], cfg["n"])
    with pytest.raises(q.QOCError):   # more ranks than slices
        sharding.InProcessSharded(cfg["A0"], cfg["A"], cfg["x0"], cost[1], (2, 3), [0, 0, 0, 0], kind="time")
    with pytest.raises(q.QOCError):   # unknown device
        sharding.InProcessSharded(cfg["A0"], cfg["A"], cfg["x0"], cost[1], cfg["u"].shape, [0, 99], kind="time")
