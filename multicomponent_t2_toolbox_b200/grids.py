"""Host-side grids of the MET2 inverse problem (T2 grid, flip-angle grids, lambda grid, regularisation matrices).

Mirrors the hard-coded set-up of the reference orchestrator
(motor/motor_recon_met2_real_data.py:204-273 and :86-111).  Pure numpy, tiny, runs once per reconstruction.
"""
import math

import numpy as np

REG_MATRICES = ("I", "L1", "L2", "InvT2")
REG_METHODS = ("NNLS", "T2SPARC", "X2", "L_curve", "GCV", "BayesReg")
FA_METHODS = ("spline", "brute-force")


def t2_grid(npc, t2_min=10.0, t2_max=2000.0):
    """`T2s = logspace(log10(10), log10(2000), Npc)` (motor...:215-220)."""
    return np.logspace(math.log10(t2_min), math.log10(t2_max), num=npc, endpoint=True, base=10.0)


def default_npc(reg_method):
    """Npc = 96 for T2SPARC else 60 (motor...:207-213)."""
    return 96 if reg_method == "T2SPARC" else 60


def compartment_masks(T2s, myelin_T2, t2_tissue=200.0):
    """ind_m / ind_t / ind_csf (motor...:222-224).  Note ind_t and ind_csf overlap at T2 == 200 exactly."""
    ind_m = T2s <= myelin_T2
    ind_t = (T2s > myelin_T2) & (T2s <= t2_tissue)
    ind_csf = T2s >= t2_tissue
    return ind_m, ind_t, ind_csf


def fa_grids(FA_method, n_alphas=None):
    """Flip-angle grids (motor...:231-245): spline -> 273-grid + 15 knots; brute-force -> 91-grid."""
    if FA_method == "spline":
        alpha_values = np.linspace(90.0, 180.0, 91 * 3 if n_alphas is None else n_alphas)
        alpha_values_spline = np.linspace(90.0, 180.0, 15)
        return alpha_values, alpha_values_spline
    if FA_method == "brute-force":
        alpha_values = np.linspace(90.0, 180.0, 91 if n_alphas is None else n_alphas)
        return alpha_values, None
    raise ValueError("Error: Wrong FA_method option!")


def lambda_grid(num=50, lam_min=1e-8, lam_max=10.0):
    """L-curve grid: lambda_reg[0] = 0, lambda_reg[1:] = logspace(1e-8, 10, 49) (motor...:248-251)."""
    lam = np.zeros((num,))
    lam[1:] = np.logspace(math.log10(lam_min), math.log10(lam_max), num=num - 1, endpoint=True, base=10.0)
    return lam


def create_Laplacian_matrix(Npc, order):
    """Dense regularisation matrix of order 0 (I), 1 (lower bidiagonal) or 2 (tridiagonal, Neumann corners).

    Same matrices as motor/motor_recon_met2_real_data.py:86-111, built directly instead of through scipy.sparse.
    """
    if order == 0:
        return np.eye(Npc)
    if order == 1:
        L = np.eye(Npc)
        L[np.arange(1, Npc), np.arange(0, Npc - 1)] = -1.0
        return L
    if order == 2:
        L = 2.0 * np.eye(Npc)
        L[np.arange(1, Npc), np.arange(0, Npc - 1)] = -1.0
        L[np.arange(0, Npc - 1), np.arange(1, Npc)] = -1.0
        L[0, 0] = 1.0
        L[-1, -1] = 1.0
        return L
    raise ValueError("order must be 0, 1 or 2")


def reg_matrix(name, T2s):
    """`Laplac` for reg_matrix in {I, L1, L2, InvT2} (motor...:254-273)."""
    npc = len(T2s)
    if name == "I":
        return create_Laplacian_matrix(npc, 0)
    if name == "L1":
        return create_Laplacian_matrix(npc, 1)
    if name == "L2":
        return create_Laplacian_matrix(npc, 2)
    if name == "InvT2":
        T2s_mod = np.concatenate((np.array([T2s[0] - 1.0]), T2s[:-1]))
        deltaT2 = T2s - T2s_mod
        deltaT2[0] = deltaT2[1]
        return np.diag(1.0 / deltaT2)
    raise ValueError("Error: Wrong reg_matrix option!")
