"""The CPU oracle (oracle/met2_oracle.py) held to the golden vectors produced by the unmodified reference
(oracle/make_golden.py).  Runs on any box; this is what pins the oracle (the reference itself ships no
tests or fixtures)."""
import numpy as np
import pytest

import met2_oracle as O

METHODS = ["NNLS", "T2SPARC", "X2", "L_curve", "GCV", "BayesReg"]
MATRICES = ["I", "L1", "L2", "InvT2"]


@pytest.fixture(scope="module")
def dics(golden_voxels):
    T2s = golden_voxels["T2s"]
    T1s = 1000.0 * np.ones(60)
    a273 = np.linspace(90.0, 180.0, 273)
    a15 = np.linspace(90.0, 180.0, 15)
    a91 = np.linspace(90.0, 180.0, 91)
    mk = lambda a: O.create_Dic_3D(60, T2s, T1s, 32, 10.0, a, 1000.0)
    return dict(a273=a273, a15=a15, a91=a91, d273=mk(a273), d15=mk(a15), d91=mk(a91))


def test_dictionary_matches_reference(golden_dictionary):
    g = golden_dictionary
    D = O.create_Dic_3D(60, g["T2s"], g["T1s"], int(g["nte"]), float(g["tau"]), g["alphas"], float(g["TR"]))
    assert np.abs(D - g["dic"]).max() <= 1e-13
    T2s100 = np.logspace(1, np.log10(2000.0), 100)
    D48 = O.create_Dic_3D(100, T2s100, 1000.0 * np.ones(100), 48, 8.0, np.array([90.0, 133.0, 180.0]), 2000.0)
    assert np.abs(D48 - g["dic48"]).max() <= 1e-13


def test_180_degree_column_is_monoexponential(golden_dictionary):
    g = golden_dictionary
    col = g["dic"][:, 0, 3]   # alpha = 180, T2 = 10 ms
    te = 10.0 * np.arange(1, 33)
    assert np.allclose(col, (1 - np.exp(-1000.0 / 1000.0)) * np.exp(-te / 10.0), rtol=1e-12, atol=0)


def test_fa_brute_force_row(golden_voxels, dics):
    g = golden_voxels
    FA, idx, KM, fs = O.fitting_slice_FA_brute_force(g["mask"], g["sig"], len(g["mask"]), dics["d91"], dics["a91"])
    assert np.array_equal(idx, g["fa_brute_idx"])
    assert np.array_equal(FA, g["fa_brute_deg"])
    assert np.allclose(KM, g["fa_brute_km"], rtol=1e-12, atol=0)
    assert np.allclose(fs, g["fa_brute_fsum"], rtol=1e-12, atol=1e-12)
    assert idx[3] == 0 and idx[5] == 0        # empty / masked voxels stay zero


def test_fa_spline_row(golden_voxels, dics):
    g = golden_voxels
    FA, idx, KM, fs = O.fitting_slice_FA_spline_method(dics["d15"], dics["d273"], g["sig"], g["mask"], dics["a15"],
                                                       len(g["mask"]), dics["a273"])
    assert np.array_equal(idx, g["fa_spline_idx"])
    assert np.allclose(KM, g["fa_spline_km"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("rm", MATRICES)
def test_t2_row_worker(golden_voxels, dics, method, rm):
    g = golden_voxels
    L = O._grids(method, rm, "spline", 40.0, 32, 10.0, 1000.0, npc=60)["L"]
    nx = len(g["mask"])
    f, s, reg = O.fitting_slice_T2(g["mask"], g["sig"], g["fa_spline_idx"], nx, dics["d273"], g["lambda_reg"], 60, 32,
                                   method, L)
    gf, gs, greg = g["t2_%s_%s_f" % (method, rm)], g["t2_%s_%s_s" % (method, rm)], g["t2_%s_%s_reg" % (method, rm)]
    assert np.array_equal(f > 0, gf > 0)
    assert np.allclose(f, gf, rtol=1e-12, atol=1e-12 * np.abs(gf).max())
    assert np.allclose(s, gs, rtol=1e-12, atol=1e-12 * np.abs(gs).max())
    assert np.allclose(reg, greg, rtol=1e-12, atol=0)
    assert not f[3].any() and not f[5].any() and not f[9].any()   # skipped voxels are all-zero


def test_t2sparc_cli_configuration(golden_voxels):
    g = golden_voxels
    gr = O._grids("T2SPARC", "InvT2", "brute-force", 40.0, 32, 10.0, 1000.0)
    assert gr["npc"] == 96
    D = O.create_Dic_3D(96, gr["T2s"], gr["T1s"], 32, 10.0, gr["alpha_values"], 1000.0)
    f, s, reg = O.fitting_slice_T2(g["mask"], g["sig"], g["fa_brute_idx"], len(g["mask"]), D, g["lambda_reg"], 96, 32,
                                   "T2SPARC", gr["L"])
    assert np.array_equal(f > 0, g["t2sparc96_f"] > 0)
    assert np.allclose(f, g["t2sparc96_f"], rtol=1e-12, atol=1e-12 * np.abs(f).max())
    assert np.all(reg[[0, 1, 2]] == 1.8)


def test_nnls_known_answer(golden_voxels, dics):
    g = golden_voxels
    D = np.ascontiguousarray(dics["d273"][:, :, int(g["nnls_D_index"])])
    M = g["sig"][0] / g["sig"][0, 0]
    x, rn = O.nnls(D, M)
    assert np.array_equal(x, g["nnls_x"]) and rn == float(g["nnls_rnorm"])
    # our own Lawson-Hanson restatement agrees with it
    x2, rn2, mode = O.lh_nnls(D, M)
    assert mode == 1 and np.array_equal(x2 > 0, x > 0)
    assert np.allclose(x2, x, rtol=1e-8, atol=1e-12) and abs(rn2 - rn) <= 1e-10 * rn


def test_config2_subset_slice(golden_config2, dics):
    """The oracle port against the reference outputs on voxels of BASELINE.json configs[1] (FA spline + X2-I)."""
    g = golden_config2
    sl = slice(0, 20480, 160)            # 128 voxels spread over the subset
    sig = g["sig"][sl]
    nx = sig.shape[0]
    ok = np.ones(nx)
    FA, idx, KM, _ = O.fitting_slice_FA_spline_method(dics["d15"], dics["d273"], sig, ok, dics["a15"], nx, dics["a273"])
    assert np.array_equal(idx.astype(np.int16), g["fa_idx"][sl])
    assert np.allclose(KM, g["km"][sl], rtol=1e-12, atol=0)
    lam = np.zeros(50)
    lam[1:] = np.logspace(-8, 1, 49)
    f, s, reg = O.fitting_slice_T2(ok, sig, idx, nx, dics["d273"], lam, 60, 32, "X2", np.eye(60))
    assert np.array_equal(f > 0, g["f"][sl] > 0)
    assert np.allclose(f, g["f"][sl], rtol=1e-10, atol=1e-12 * g["f"][sl].max())
    assert np.allclose(reg, g["reg"][sl], rtol=1e-10, atol=0)
