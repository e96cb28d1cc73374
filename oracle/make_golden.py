"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz from OUTPUTS OF THE UNMODIFIED REFERENCE.

Runs only where /root/reference exists (the build container).  The reference modules are imported read-only through
oracle/ref_shim.py and driven with seeded phantom voxels; inputs and outputs are stored so that both the oracle port
(tests/test_oracle_golden.py, any box) and the CUDA path (tests/test_gpu_parity.py, GPU box) can be held to them.

    python oracle/make_golden.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

warnings.simplefilter("ignore")
OUT = os.path.join(ROOT, "tests", "golden")
METHODS = ["NNLS", "T2SPARC", "X2", "L_curve", "GCV", "BayesReg"]
MATRICES = ["I", "L1", "L2", "InvT2"]


def laplacian(R, name, T2s):
    n = len(T2s)
    if name == "InvT2":   # motor/motor_recon_met2_real_data.py:263-269
        T2s_mod = np.concatenate((np.array([T2s[0] - 1.0]), T2s[:-1]))
        d = T2s - T2s_mod
        d[0] = d[1]
        return np.diag(1.0 / d)
    return R["motor"].create_Laplacian_matrix(n, {"I": 0, "L1": 1, "L2": 2}[name])


def main():
    R = ref_shim.load_reference()
    os.makedirs(OUT, exist_ok=True)
    nte, tau, TR = 32, 10.0, 1000.0
    T2s = np.logspace(np.log10(10.0), np.log10(2000.0), 60)
    T1s = 1000.0 * np.ones(60)
    a91 = np.linspace(90.0, 180.0, 91)
    a273 = np.linspace(90.0, 180.0, 273)
    a15 = np.linspace(90.0, 180.0, 15)
    # ---- dictionary (epg/epg.py:155)
    sel = np.array([0, 45, 136, 272])
    dic_sel = R["epg"].create_Dic_3D(60, T2s, T1s, nte, tau, a273[sel], TR)
    dic48 = R["epg"].create_Dic_3D(100, np.logspace(1, np.log10(2000.0), 100), 1000.0 * np.ones(100), 48, 8.0,
                                   np.array([90.0, 133.0, 180.0]), 2000.0)
    np.savez_compressed(os.path.join(OUT, "dictionary.npz"), T2s=T2s, T1s=T1s, alphas=a273[sel], sel=sel, dic=dic_sel,
                        nte=nte, tau=tau, TR=TR, dic48=dic48)
    # ---- voxels
    ph = make_phantom((8, 6, 1), n_echoes=nte, tau=tau, TR=TR, seed=20261018)
    sig = ph["data"].reshape(-1, nte).copy()
    sig[3] = 0.0          # empty voxel: skipped by every stage
    sig[9, 0] = 0.0       # M[0] == 0: FA stage fits it, T2 stage skips it (motor...:127-131)
    nx = sig.shape[0]
    mask = np.ones(nx)
    mask[5] = 0.0         # masked-out voxel
    Dic91 = R["epg"].create_Dic_3D(60, T2s, T1s, nte, tau, a91, TR)
    Dic273 = R["epg"].create_Dic_3D(60, T2s, T1s, nte, tau, a273, TR)
    Dic15 = R["epg"].create_Dic_3D(60, T2s, T1s, nte, tau, a15, TR)
    FAb, FAb_idx, KMb, Fsb = R["fa"].fitting_slice_FA_brute_force(mask, sig, nx, Dic91, a91)
    FAs, FAs_idx, KMs, Fss = R["fa"].fitting_slice_FA_spline_method(Dic15, Dic273, sig, mask, a15, nx, a273)
    gold = dict(sig=sig, mask=mask, T2s=T2s, fa_brute_deg=FAb, fa_brute_idx=FAb_idx, fa_brute_km=KMb, fa_brute_fsum=Fsb,
                fa_spline_deg=FAs, fa_spline_idx=FAs_idx, fa_spline_km=KMs, fa_spline_fsum=Fss)
    lam = np.zeros(50)
    lam[1:] = np.logspace(-8, 1, 49)
    gold["lambda_reg"] = lam
    for method in METHODS:
        for rm in MATRICES:
            L = laplacian(R, rm, T2s)
            f, s, reg = R["motor"].fitting_slice_T2(mask, sig, FAs_idx, nx, Dic273, lam, 60, nte, method, L, None)
            gold["t2_%s_%s_f" % (method, rm)] = f
            gold["t2_%s_%s_s" % (method, rm)] = s
            gold["t2_%s_%s_reg" % (method, rm)] = reg
            print(method, rm, "done", flush=True)
    # T2SPARC as the CLI runs it: 96 T2 bins, InvT2 (run_real_data_script.py:91-93, motor...:207-210)
    T2s96 = np.logspace(np.log10(10.0), np.log10(2000.0), 96)
    Dic96 = R["epg"].create_Dic_3D(96, T2s96, 1000.0 * np.ones(96), nte, tau, a91, TR)
    L96 = laplacian(R, "InvT2", T2s96)
    f, s, reg = R["motor"].fitting_slice_T2(mask, sig, FAb_idx, nx, Dic96, lam, 96, nte, "T2SPARC", L96, None)
    gold["t2sparc96_f"], gold["t2sparc96_s"], gold["t2sparc96_reg"] = f, s, reg
    # per-voxel API samples (algorithms.py): residual norms of nnls, lcurve curves are covered through the row worker
    D = np.ascontiguousarray(Dic273[:, :, 136])
    M = sig[0] / sig[0, 0]
    x, rn = R["algorithms"].nnls(D, M)
    gold["nnls_x"], gold["nnls_rnorm"], gold["nnls_D_index"] = x, rn, 136
    np.savez_compressed(os.path.join(OUT, "voxels.npz"), **gold)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
