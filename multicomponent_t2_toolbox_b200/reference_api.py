"""Drop-in forms of the reference's Python call signatures, backed by the batched CUDA path.

Row workers and per-voxel solvers keep the reference's names, argument order and return values; internally each call
is one batched launch over all the voxels it was given (V = nx for a row worker, V = 1 for a per-voxel call).  The
thin modules epg/epg.py, flip_angle_algorithms/fa_estimation.py, intravoxel_algorithms/*.py and
motor/motor_recon_met2_real_data.py re-export these under the reference's module paths.
"""
import hashlib

import numpy as np
import torch

from . import batched, grids

_PLAN_CACHE = {}
_CACHE_MAX = 8


def _key(*arrays, extra=()):
    """Cache key over the FULL contents of every array (shape, dtype, SHA-1 of the contiguous bytes): a sampled
    fingerprint cannot tell I from L1 from L2 (all have 1.0 at both ends of the diagonal)."""
    parts = list(extra)
    for a in arrays:
        if a is None:
            parts.append(None)
            continue
        a = np.ascontiguousarray(a)
        parts.append((a.shape, a.dtype.str, hashlib.sha1(a.view(np.uint8).reshape(-1)).hexdigest()))
    return tuple(parts)


def _plan(Dic_3D, Laplac=None, reg_method="NNLS", FA_method="brute-force", Dic_3D_LR=None, alpha_values=None,
          alpha_values_spline=None, lambda_reg=None, T2s=None, myelin_T2=40.0):
    Dic_3D = np.asarray(Dic_3D, dtype=np.float64)
    if Dic_3D.ndim == 2:
        Dic_3D = Dic_3D[:, :, None]
    n = Dic_3D.shape[1]
    if Laplac is None:
        Laplac = np.eye(n)
    key = _key(Dic_3D, Laplac, Dic_3D_LR, alpha_values, alpha_values_spline, lambda_reg, T2s,
               extra=(reg_method, FA_method, myelin_T2))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= _CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        plan = batched.Met2Plan(Dic_3D.shape[0], 1.0, 1.0, reg_method=reg_method, reg_matrix="I", FA_method=FA_method,
                                myelin_T2=myelin_T2, Laplac=Laplac, Dic_3D=Dic_3D, Dic_3D_LR=Dic_3D_LR,
                                alpha_values=alpha_values, alpha_values_spline=alpha_values_spline,
                                lambda_reg=lambda_reg, T2s=T2s)
        _PLAN_CACHE[key] = plan
    return plan


# ---------------------------------------------------------------------------------------------------- epg/epg.py
def create_Dic_3D(Npc, T2s, T1s, nEchoes, tau, alpha_values, TR):
    """epg/epg.py:155 — Dic_3D[nEchoes, Npc, len(alpha_values)] (host copy of the GPU-built dictionary)."""
    dev = batched._require_cuda(None)
    d = batched.Dictionary(np.asarray(alpha_values, dtype=np.float64), np.asarray(T2s)[:Npc], np.asarray(T1s)[:Npc],
                           nEchoes, tau, TR, dev)
    return d.to_reference_layout()


def create_met2_design_matrix_epg(Npc, T2s, T1s, nEchoes, tau, flip_angle, TR):
    """epg/epg.py:47 — [nEchoes, Npc] design matrix for one refocusing angle (degrees)."""
    return create_Dic_3D(Npc, T2s, T1s, nEchoes, tau, np.array([flip_angle], dtype=np.float64), TR)[:, :, 0]


# ------------------------------------------------------------------ flip_angle_algorithms/fa_estimation.py
def _fa_rows(plan, data_1d, mask_1d, nx):
    data_1d = np.asarray(data_1d, dtype=np.float64)[:nx]
    keep = np.nonzero(np.asarray(mask_1d)[:nx] > 0.0)[0]
    FA = np.zeros(nx)
    FA_index = np.zeros(nx)
    KM = np.zeros(nx)
    Fsol = 0.0
    if keep.size:
        out = plan.fa_fit(data_1d[keep])
        FA[keep] = out["fa_deg"].cpu().numpy()
        FA_index[keep] = out["fa_index"].cpu().numpy()
        KM[keep] = out["km"].cpu().numpy()
        if np.any(out["status"].cpu().numpy() == 0):
            Fsol = out["fsol_sum"].cpu().numpy()
    return FA, FA_index, KM, Fsol


def fitting_slice_FA_brute_force(mask_1d, data_1d, nx, Dic_3D, alpha_values):
    """fa_estimation.py:92 — returns (FA[nx], FA_index[nx], KM[nx], sum of spectra)."""
    plan = _plan(Dic_3D, FA_method="brute-force", alpha_values=alpha_values)
    return _fa_rows(plan, data_1d, mask_1d, nx)


def fitting_slice_FA_spline_method(Dic_3D_LR, Dic_3D, data_1d, mask_1d, alpha_values_spline, nx, alpha_values):
    """fa_estimation.py:35."""
    plan = _plan(Dic_3D, FA_method="spline", Dic_3D_LR=Dic_3D_LR, alpha_values=alpha_values,
                 alpha_values_spline=alpha_values_spline)
    return _fa_rows(plan, data_1d, mask_1d, nx)


def compute_optimal_FA(M, Dic_3D, alpha_values):
    """fa_estimation.py:74 — (index, alpha, km, SSE, f) for one voxel."""
    plan = _plan(Dic_3D, FA_method="brute-force", alpha_values=alpha_values)
    M = np.ascontiguousarray(M, dtype=np.float64)
    fa = plan.fa_fit(M[None, :])
    index = int(fa["fa_index"][0])
    t2 = plan.t2_fit(M[None, :], fa["fa_index"], reg_method="NNLS", flags=2)
    f = t2["fsol"][0].cpu().numpy()
    Di = np.asarray(Dic_3D)[:, :, index]
    SSE = np.sum((np.dot(Di, f) - M) ** 2)
    return index, np.asarray(alpha_values)[index], float(fa["km"][0]), SSE, f


# ------------------------------------------------------------------ intravoxel_algorithms/algorithms.py
def _solve(D, M, L, method, lambda_reg=None, lam_fixed=None, factor=1.02):
    D = np.asarray(D, dtype=np.float64)
    M = np.asarray_chkfinite(np.asarray(M, dtype=np.float64))
    plan = _plan(D, Laplac=L, reg_method=method, lambda_reg=lambda_reg)
    single = (M.ndim == 1)
    Mb = M[None, :] if single else M
    zeros = np.zeros(len(Mb), dtype=np.int32)
    # flags = REG_IS_LAMBDA | NO_NORMALISE: the per-voxel functions fit M as given and return the selected lambda
    if lam_fixed is not None:   # nnls_tik with an arbitrary lambda = the fixed-lambda (T2SPARC) path
        out = plan.t2_fit(Mb, zeros, reg_method="T2SPARC", flags=3, lambda_fixed=float(lam_fixed))
    else:
        out = plan.t2_fit(Mb, zeros, reg_method=method, flags=3, factor=float(factor))
    f = out["fsol"].cpu().numpy()
    reg = out["reg"].cpu().numpy()
    return (f[0], reg[0]) if single else (f, reg)


def nnls(A, b):
    """algorithms.py:55 — (x, rnorm) with rnorm = |A x - b|_2 like Lawson-Hanson's.  A is [m, n] with m <= 64 (a
    per-echo dictionary), or the STACKED Tikhonov system the reference builds at algorithms.py:77,229,264,286 —
    [D; sqrt(lambda) L] with a zero tail of b — recognised by its banded square bottom block and solved through the
    Tikhonov path (D, b_top, L' = sqrt(lambda) L, lambda = 1)."""
    A = np.asarray_chkfinite(np.asarray(A, dtype=np.float64))
    b = np.asarray_chkfinite(np.asarray(b, dtype=np.float64))
    m, n = A.shape
    if m > 64:
        top = m - n
        Lp = A[top:] if top > 0 else None
        off = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) > 2
        if Lp is None or top > 64 or np.any(b[top:] != 0.0) or np.any(Lp[off] != 0.0):
            raise ValueError("nnls: A has %d rows; supported are m <= 64 or the stacked system [D; sqrt(lambda) L] with "
                             "a banded L and a zero tail of b" % m)
        x, _ = _solve(A[:top], b[:top], Lp, "T2SPARC", lam_fixed=1.0)
    else:
        x, _ = _solve(A, b, None, "NNLS")
    return x, float(np.sqrt(np.sum((np.dot(A, x) - b) ** 2)))


def nnls_tik(Dic_i, M, Laplac, reg_opt):
    """algorithms.py:262 — f = nnls([D; sqrt(reg) L], [M; 0])."""
    f, _ = _solve(Dic_i, M, Laplac, "T2SPARC", lam_fixed=reg_opt)
    return f


def nnls_x2(Dic_i, M, Laplac, factor):
    """algorithms.py:211 — (f, reg_opt, k_est)."""
    f, reg = _solve(Dic_i, M, Laplac, "X2", factor=factor)
    x0, _ = _solve(Dic_i, M, None, "NNLS")
    D = np.asarray(Dic_i, dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)
    SSE = np.sum((np.dot(D, x0) - M) ** 2)
    k_est = np.sum((np.dot(D, f) - M) ** 2) / SSE
    return f, reg, k_est


def nnls_lcurve_wrapper(D, y, Laplac_mod, lambda_reg):
    """algorithms.py:88 — the grid lambda at the L-curve corner."""
    _, reg = _solve(D, y, Laplac_mod, "L_curve", lambda_reg=np.asarray(lambda_reg, dtype=np.float64))
    return reg


def nnls_gcv(Dic_i, M, L):
    """algorithms.py:276 — (f, reg_opt)."""
    return _solve(Dic_i, M, L, "GCV")


def BayesReg_nnls(Dic_i, M, L):
    """bayesian_interpolation.py:84 — (f, reg_sol)."""
    return _solve(Dic_i, M, L, "BayesReg")


# ------------------------------------------------------------------ motor/motor_recon_met2_real_data.py
def create_Laplacian_matrix(Npc, order):
    """motor...:86."""
    return grids.create_Laplacian_matrix(Npc, order)


def fitting_slice_T2(mask_1d, data_1d, FA_index_1d, nx, Dic_3D, lambda_reg, T2dim, nEchoes, reg_method, Laplac,
                     dist_x_prior=None):
    """motor...:113 — (f[nx, T2dim], signal[nx, nEchoes], reg[nx])."""
    plan = _plan(Dic_3D, Laplac=Laplac, reg_method=reg_method, lambda_reg=lambda_reg)
    data_1d = np.asarray(data_1d, dtype=np.float64)[:nx]
    keep = np.nonzero(np.asarray(mask_1d)[:nx] > 0.0)[0]
    f = np.zeros((nx, T2dim))
    s = np.zeros((nx, nEchoes))
    reg = np.zeros(nx)
    if keep.size:
        out = plan.t2_fit(data_1d[keep], np.asarray(FA_index_1d)[:nx][keep].astype(np.int32))
        f[keep] = out["fsol"].cpu().numpy()
        s[keep] = out["est_signal"].cpu().numpy()
        reg[keep] = out["reg"].cpu().numpy()
    return f, s, reg
