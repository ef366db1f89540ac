"""CPU-only checks of the host-side mirror: array marshalling conventions and the input generators."""
import numpy as np
import pytest

import qoc_b200
from qoc_b200 import configs, grape


def test_u_layout_is_julia_column_major():
    cache = qoc_b200.setup_grape_cache(np.zeros((3, 3), complex), np.zeros((3, 1), complex), (2, 5))
    u = np.arange(10.0).reshape(2, 5)
    uu = grape._u_arr(u, cache)
    flat = uu.ravel()
    for j in range(2):
        for k in range(5):
            assert flat[j + 2 * k] == u[j, k]
    with pytest.raises(qoc_b200.QOCError):
        grape._u_arr(np.zeros((2, 6)), cache)
    cb = qoc_b200.setup_grape_cache(np.zeros((3, 3), complex), np.zeros((3, 1), complex), (2, 5), batch=3)
    ub = np.arange(30.0).reshape(3, 2, 5)
    fb = grape._u_arr(ub, cb).ravel()
    assert fb[1 + 2 * (4 + 5 * 2)] == ub[2, 1, 4]


def test_c128_is_fortran_interleaved():
    a = np.array([[1 + 2j, 3 + 4j], [5 + 6j, 7 + 8j]])
    f = grape._c128(a)
    assert f.flags.f_contiguous
    raw = np.frombuffer(f.tobytes(order="A"), dtype=np.float64)
    assert list(raw[:4]) == [1, 2, 5, 6]  # column-major interleaved: re(a11), im(a11), re(a21), im(a21)


def test_builtin_cost_closures_match_reference_formulae():
    rng = np.random.default_rng(0)
    T = rng.standard_normal((9, 4)) + 1j * rng.standard_normal((9, 4))
    x = rng.standard_normal((9, 4)) + 1j * rng.standard_normal((9, 4))
    J, dJ = qoc_b200.setup_infidelity(T, 4)
    om = np.trace(T.conj().T @ x)
    assert J(x) == pytest.approx(1 - abs(om) ** 2 / 16)
    assert np.allclose(dJ(x), (-2 * om / 16) * T)
    L, dL = qoc_b200.setup_state_penalty([6, 7, 8], [0, 1, 2, 3], 0.22)
    assert L(x) == pytest.approx(0.22 * np.sum(abs(x[6:9, :]) ** 2))
    assert np.allclose(dL(x)[6:9], 0.44 * x[6:9]) and np.all(dL(x)[:6] == 0)


def test_config_generators():
    zz = configs.config_zz()
    assert zz["A0"].shape == (9, 9) and zz["u"].shape == (2, 100) and zz["x0"].shape == (9, 4)
    assert zz["qb"](["00", "01", "10", "11"]) == [0, 1, 3, 4] and zz["qb"](["20", "21", "22"]) == [6, 7, 8]
    B = zz["B"]
    assert B.shape == (100, 10) and abs(B.sum(1)[50] - 1) < 1e-12 and abs(B.sum(1)[0] - 4.577083e-5) < 1e-9
    nrm = [np.linalg.norm(zz["A0"] + zz["u"][0, k] * zz["A"][0] + zz["u"][1, k] * zz["A"][1], 1) for k in range(100)]
    assert 0.25 < min(nrm) and max(nrm) < 0.27     # [0.2521, 0.2697]
    bus = configs.config_bus(Nt=100)
    assert bus["A0"].shape == (27, 27) and bus["u"].shape == (1, 100) and np.all(bus["u"] > 0)
    cav = configs.config_cavity(12, Nt=550)
    assert cav["u"].shape == (2, 550) and abs(np.abs(cav["u"]).max() - 0.0506) < 1e-3
    # skew-Hermitian generators
    for c in (zz, bus, cav):
        for M in [c["A0"]] + c["A"]:
            assert np.abs(M + M.conj().T).max() < 1e-12
    s = configs.config_synthetic(16, 8)
    assert abs(np.linalg.norm(s["A0"], 1) - 2) < 1e-12 and abs(np.linalg.norm(s["A"][0], 1) - 1) < 1e-12
    assert np.allclose(s["T"].conj().T @ s["T"], np.eye(4))
    b = configs.config_zz_batch(3)
    assert b["u_batch"].shape == (3, 2, 100) and np.abs(b["u_batch"]).max() <= 2 * np.pi * 0.06 + 1e-12


def test_bench_l2_note_and_ncu_summary_tool(tmp_path):
    """bench.py states how the timed steps relate to the L2 for every workload; tools/ncu_summary.py condenses an
    `ncu --page raw --csv` export into the JSON kept under profiles/."""
    import importlib.util, json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert "exceeds the 126 MB L2" in bench.l2_note(27, 10000, 1, 1)          # C2: 242 MB of U_k / dU_k per step
    assert "242 MB" in bench.l2_note(27, 10000, 1, 1)
    assert "stays in" in bench.l2_note(24, 550, 1, 2)                          # C3: 17.7 MB
    raw = tmp_path / "raw.csv"
    cols = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
    raw.write_text(",".join('"%s"' % c for c in cols) + "\n" + '"","","ms","Mbyte","Kbyte","inst"\n'
                   + '"0","k1_kernel","1.25","0.5","1,500.0","2.5"\n')
    out = tmp_path / "out.json"
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), str(raw), str(out), "src", "wl"])
    k = json.load(open(out))["kernels"][0]
    assert k["name"] == "k1_kernel" and k["gpu__time_duration.sum"]["value"] == 1.25
    assert abs(k["traffic_bytes_per_launch"] - (0.5e6 + 1500.0e3)) < 1e-6
