"""Restatements of third-party arithmetic inside the oracle (Lawson-Hanson NNLS, bounded Brent) against the SciPy
routines the reference calls, plus KKT properties of every NNLS output."""
import numpy as np
import pytest
from scipy.optimize import fminbound

import met2_oracle as O
from multicomponent_t2_toolbox_b200.phantom import make_phantom


@pytest.fixture(scope="module")
def problem():
    ph = make_phantom((6, 4, 1), seed=5)
    sig = ph["data"].reshape(-1, 32)
    T2s = np.logspace(1, np.log10(2000.0), 60)
    Dic = O.create_Dic_3D(60, T2s, 1000.0 * np.ones(60), 32, 10.0, np.array([105.0, 150.0, 180.0]), 1000.0)
    return sig, Dic


@pytest.mark.parametrize("lam", [0.0, 1e-8, 1e-4, 1e-2, 1.0, 3.8197])
def test_lh_restatement_matches_scipy(problem, lam):
    sig, Dic = problem
    for v in range(0, 24, 3):
        M = sig[v] / sig[v, 0]
        D = np.ascontiguousarray(Dic[:, :, v % 3])
        if lam == 0.0:
            A, b = D, M
        else:
            A = np.concatenate((D, np.sqrt(lam) * np.eye(60)))
            b = np.concatenate((M, np.zeros(60)))
        x0, r0 = O.nnls(A, b)
        c = O.FlopCounter()
        x1, r1, mode = O.lh_nnls(A, b, counter=c)
        assert mode == 1
        assert np.array_equal(x0 > 0, x1 > 0)
        assert np.abs(x0 - x1).max() <= 1e-8 * np.abs(x0).max()
        assert abs(r0 - r1) <= 1e-9 * r0
        assert c.flops > 0 and c.solves == 1
        # KKT: x >= 0, gradient <= tol on the zero set, ~0 on the support
        w = A.T @ (b - A @ x1)
        tol = 1e-9 * np.abs(A.T @ b).max()
        assert np.all(x1 >= 0) and np.all(w[x1 == 0] <= tol) and np.all(np.abs(w[x1 > 0]) <= tol)


def test_nnls_rejects_nonfinite(problem):
    sig, Dic = problem
    M = sig[0].copy()
    M[3] = np.nan
    with pytest.raises(ValueError):
        O.nnls(np.ascontiguousarray(Dic[:, :, 0]), M)


def test_brent_restatement_follows_scipy_exactly():
    rng = np.random.default_rng(0)
    for t in range(200):
        a, b, c = rng.uniform(0.1, 3), rng.uniform(-2, 12), rng.uniform(0, 2)
        if t % 2:
            f = lambda x: abs(a * np.sin(c * x) + 0.05 * (x - b) ** 2 - 0.3)
        else:
            f = lambda x: a * np.cos(c * x) + 0.01 * (x - b) ** 2
        x0, fv, ierr, nf = fminbound(f, 0.0, 10.0, xtol=1e-5, maxfun=300, full_output=1)
        tr = []
        x1, f1, n1 = O.brent_bounded(f, 0.0, 10.0, 1e-5, 300, trace=tr)
        assert x0 == x1 and nf == n1 and len(tr) == n1


def test_brent_handles_nan_and_inf_objectives():
    x, f, n = O.brent_bounded(lambda x: float("inf"), 1e-8, 2.0, 1e-5, 200)
    xs = fminbound(lambda x: float("inf"), 1e-8, 2.0, xtol=1e-5, maxfun=200)
    assert x == xs           # BayesReg-L2 degenerate case: lambda -> 1.99999599
    assert abs(x - 1.99999599) < 1e-7


def test_select_corner_degenerate_curve_defaults_to_last():
    assert O.select_corner(np.ones(50), np.ones(50)) == 49


def test_voxel_metrics_empty_spectrum():
    T2s = np.logspace(1, np.log10(2000.0), 60)
    m, t, c = T2s <= 40, (T2s > 40) & (T2s <= 200), T2s >= 200
    r = O.voxel_metrics(np.zeros(60), T2s, m, t, c)
    assert r[:3] == (0.0, 0.0, 0.0) and r[3] == 1.0 and r[4] == 1.0 and r[5] == 1e-16
