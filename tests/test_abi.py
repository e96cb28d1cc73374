"""The C-ABI library loads on a box without a GPU, exports every symbol include/met2.h declares, validates arguments,
and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from multicomponent_t2_toolbox_b200 import _lib, batched

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "met2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(met2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    syms = declared_symbols()
    assert {"met2_epg_dictionary", "met2_gram_tables", "met2_fa_fit", "met2_t2_fit", "met2_last_error",
            "met2_version"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), "libmet2.so does not export %s" % s
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.met2_version() == 120


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.FaCfg) == 6 * 4 + 3 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.T2Cfg) == 6 * 4 + 6 * 8 + 2 * 4


def test_argument_validation_without_gpu():
    lib = _lib.load()
    bad = _lib.FaCfg(method=7, nTE=32, nT2=60, nA=91)
    assert lib.met2_fa_workspace_bytes(10, ctypes.byref(bad)) == -1
    assert b"unknown method" in lib.met2_last_error()
    big = _lib.FaCfg(method=0, nTE=32, nT2=500, nA=91)
    assert lib.met2_fa_workspace_bytes(10, ctypes.byref(big)) == -1
    ok = _lib.FaCfg(method=1, nTE=32, nT2=60, nA=273, nKnots=15)
    assert lib.met2_fa_workspace_bytes(1000, ctypes.byref(ok)) > 1000 * 15 * 8
    t2 = _lib.T2Cfg(method=2, nTE=32, nT2=60, nA=273, nLambda=50)
    assert lib.met2_t2_workspace_bytes(1000, ctypes.byref(t2)) > 4000
    rc = lib.met2_t2_fit(None, None, 5, ctypes.byref(t2), *([None] * 14))
    assert rc == -1 and b"NULL" in lib.met2_last_error()
    rc = lib.met2_epg_dictionary(None, 3, None, None, 60, 32, 10.0, 1000.0, None, None, None)
    assert rc == -1
    # GCV lambda-grid mode: the kernel stages at most MET2_MAX_LAMBDAS (64) grid values
    gcv = _lib.T2Cfg(method=4, nTE=32, nT2=60, nA=91, nLambda=65, flags=32)
    assert lib.met2_t2_workspace_bytes(10, ctypes.byref(gcv)) == -1 and b"GCV grid" in lib.met2_last_error()
    gcv.nLambda = 0
    assert lib.met2_t2_workspace_bytes(10, ctypes.byref(gcv)) == -1
    gcv.nLambda = 49
    assert lib.met2_t2_workspace_bytes(10, ctypes.byref(gcv)) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
    with pytest.raises(_lib.Met2Error):
        batched.Met2Plan(32, 10.0, 1000.0)


def _c_host_binary():
    from multicomponent_t2_toolbox_b200 import build
    if not os.path.exists(build.EXAMPLE_BIN):
        build.build_c_host_example()
    return build.EXAMPLE_BIN


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_c_host_program_links_and_fails_loudly_without_cuda():
    """examples/c_host: a C++ program on the C ABI alone (no Python, no torch).  Without a GPU it must load libmet2.so,
    report the ABI version and stop with an error — not compute anything on the CPU."""
    import subprocess
    r = subprocess.run([_c_host_binary()], capture_output=True, text=True, timeout=120)
    assert "met2 C-ABI version 120" in r.stdout
    assert r.returncode != 0 and ("error" in r.stderr.lower())


@pytest.mark.gpu
def test_c_host_program_runs_the_path_through_the_c_abi():
    """Dictionary -> FA brute force -> X2 fit + maps driven from C++ through include/met2.h only, with self-checks."""
    import subprocess
    r = subprocess.run([_c_host_binary()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK") and " 0/256 wrong" in r.stdout
