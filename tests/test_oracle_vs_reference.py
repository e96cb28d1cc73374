"""Direct comparison of the oracle with the unmodified reference (imported read-only through oracle/ref_shim.py).
Only runs where /root/reference exists (the build container); the committed golden vectors carry the same
information to other boxes."""
import numpy as np
import pytest

import met2_oracle as O
import ref_shim
from multicomponent_t2_toolbox_b200.phantom import make_phantom

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def setup():
    R = ref_shim.load_reference()
    ph = make_phantom((6, 2, 1), seed=77)
    sig = ph["data"].reshape(-1, 32)
    T2s = np.logspace(1, np.log10(2000.0), 60)
    T1s = 1000.0 * np.ones(60)
    a273, a15 = np.linspace(90, 180, 273), np.linspace(90, 180, 15)
    return R, sig, T2s, O.create_Dic_3D(60, T2s, T1s, 32, 10.0, a273, 1000.0), \
        O.create_Dic_3D(60, T2s, T1s, 32, 10.0, a15, 1000.0), a273, a15


def test_dictionary_bitwise(setup):
    R, sig, T2s, D, DLR, a273, a15 = setup
    ref = R["epg"].create_Dic_3D(60, T2s, 1000.0 * np.ones(60), 32, 10.0, a273[[0, 100, 272]], 1000.0)
    assert np.array_equal(ref, D[:, :, [0, 100, 272]])


def test_fa_row_workers_bitwise(setup):
    R, sig, T2s, D, DLR, a273, a15 = setup
    nx = sig.shape[0]
    mask = np.ones(nx)
    a = R["fa"].fitting_slice_FA_spline_method(DLR, D, sig, mask, a15, nx, a273)
    b = O.fitting_slice_FA_spline_method(DLR, D, sig, mask, a15, nx, a273)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    sub = np.ascontiguousarray(D[:, :, ::3])
    a = R["fa"].fitting_slice_FA_brute_force(mask, sig, nx, sub, a273[::3])
    b = O.fitting_slice_FA_brute_force(mask, sig, nx, sub, a273[::3])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("method,rm", [("X2", "I"), ("X2", "L2"), ("L_curve", "L1"), ("GCV", "I"), ("GCV", "L2"),
                                       ("BayesReg", "I"), ("BayesReg", "InvT2"), ("T2SPARC", "InvT2"), ("NNLS", "I")])
def test_t2_row_worker_bitwise(setup, method, rm):
    R, sig, T2s, D, DLR, a273, a15 = setup
    nx = sig.shape[0]
    mask = np.ones(nx)
    g = O._grids(method, rm, "spline", 40.0, 32, 10.0, 1000.0, npc=60)
    idx = np.full(nx, 200.0)
    a = R["motor"].fitting_slice_T2(mask, sig, idx, nx, D, g["lambda_reg"], 60, 32, method, g["L"], None)
    b = O.fitting_slice_T2(mask, sig, idx, nx, D, g["lambda_reg"], 60, 32, method, g["L"])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
