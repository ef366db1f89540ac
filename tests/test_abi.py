"""CPU-only checks of the drop-in boundary: the C-ABI library builds/loads, exports every symbol include/qoc_b200.h
declares, and every compute entry point fails LOUDLY without a B200 (no CPU fallback, no oracle routing)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import qoc_b200
from qoc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qoc_b200.h")


def header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qoc_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_header_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/qoc_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms, "ctypes table and header drifted apart"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (qoc_[a-z0-9_]+)", out))
    assert set(syms) <= exported
    assert lib.qoc_version() >= 100


def test_library_contains_sm100a_dmma_kernels():
    # the hot path must be native sm_100a code using the FP64 tensor-core tile
    r = subprocess.run(["cuobjdump", "-sass", _lib.lib_path()], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    assert r.stdout.count("DMMA.8x8x4") > 100
    assert "LDGSTS" in r.stdout  # cp.async staging ring of K2/K3


def test_status_strings_match_reference_messages():
    lib = _lib.load()
    assert lib.qoc_status_string(_lib.ERR_STALE_CACHE) == b"Cache data from other control signal u"
    assert b"incompatiable dimensions" in lib.qoc_status_string(_lib.ERR_DIMENSION)
    assert lib.qoc_status_string(_lib.OK) == b"ok"


def test_problem_struct_layout():
    # must match the C struct (checked against the compiler's view through a tiny C program)
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu", sizeof(qoc_problem), offsetof(qoc_problem, pen_rows), offsetof(qoc_problem, mu), offsetof(qoc_problem, store_costates));}' % HEADER
    exe = "/tmp/qoc_layout_probe"
    subprocess.run(["gcc", "-x", "c", "-o", exe, "-"], input=src, text=True, check=True)
    vals = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    P = _lib.Problem
    assert vals == [C.sizeof(P), P.pen_rows.offset, P.mu.offset, P.store_costates.offset]


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="this test is about the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    d = 4
    A0 = -1j * np.eye(d)
    cache = qoc_b200.setup_grape_cache(A0, np.eye(d, 2, dtype=complex), (1, 5))
    with pytest.raises(qoc_b200.QOCError) as ei:
        qoc_b200.propagate(A0, [A0], np.zeros((1, 5)), np.eye(d, 2, dtype=complex), cache)
    assert ei.value.status == _lib.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    pr = _lib.Problem()
    pr.d, pr.m, pr.nc, pr.nt, pr.batch = 0, 1, 1, 1, 1
    z = np.zeros(8)
    p = z.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.qoc_create(C.byref(pr), p, p, p, None, C.byref(h)) == _lib.ERR_DIMENSION
    assert lib.qoc_create(None, p, p, p, None, C.byref(h)) == _lib.ERR_INVALID
    pr.d, pr.order = 2, 7
    assert lib.qoc_create(C.byref(pr), p, p, p, None, C.byref(h)) == _lib.ERR_INVALID
    assert lib.qoc_eval(None, p, p, p) == _lib.ERR_INVALID
    assert lib.qoc_destroy(None) == _lib.OK


def test_mirror_dimension_error_like_reference():
    # src/gradient_computations.jl:84-87
    with pytest.raises(qoc_b200.QOCError, match="incompatiable dimensions"):
        qoc_b200.setup_grape_cache(np.zeros((9, 9), complex), np.zeros((8, 4), complex), (2, 100))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "quantumoptimalcontrol.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "qoc_oracle" not in txt and "qoc_ref" not in txt, f


def test_c_driver_compiles_and_links_against_the_header():
    """tests/abi_driver.c (plain C99) builds against include/qoc_b200.h and links the in-tree library: the header is usable from
    C as shipped.  (It runs on the GPU box: tests/test_gpu_abi_driver.py.)"""
    _lib.load()
    libdir = os.path.dirname(_lib.lib_path())
    exe = "/tmp/qoc_abi_driver_linkcheck"
    r = subprocess.run(["gcc", "-O1", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "abi_driver.c"), "-o", exe, "-L", libdir, "-lqoc_b200", "-lm",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_sharded_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less build container")
    lib = _lib.load()
    pr = _lib.Problem()
    pr.d, pr.m, pr.nc, pr.nt, pr.batch, pr.order, pr.cost, pr.n, pr.device = 4, 1, 1, 8, 1, 0, 0, 1, 0
    z = np.zeros(64)
    dp = C.POINTER(C.c_double)
    h = C.c_void_p()
    rc = lib.qoc_create_sharded(C.byref(pr), z.ctypes.data_as(dp), z.ctypes.data_as(dp), z.ctypes.data_as(dp), z.ctypes.data_as(dp), 2, None, 1,
                                C.byref(h))
    assert rc == _lib.ERR_NO_DEVICE and h.value is None
    assert b"CUDA" in lib.qoc_sharded_last_error(None)
