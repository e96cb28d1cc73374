"""TEST INFRASTRUCTURE ONLY — writes tests/golden/methods_subset.npz from OUTPUTS OF THE UNMODIFIED REFERENCE.

Parity at scale for the regularisation methods other than the headline one: 2 048 voxels of the config-2 phantom
(every 10th voxel of tests/golden/config2_subset.npz, so the signals are not stored twice) are fitted by the reference's
own row workers (fa_estimation.py:35,92; motor/motor_recon_met2_real_data.py:113) for

    NNLS + brute-force FA (91 angles)      L_curve-I      BayesReg-I      X2-L2      T2SPARC-InvT2 (96 bins)
    GCV-L2 + brute-force FA (BASELINE.json configs[2] as the reference runs it; statistical parity only, SURVEY a-8)

and for each the FA index, reg_param and the spectrum (support bit mask + non-zero coefficients) are stored.

    python oracle/make_golden_methods.py
"""
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from make_golden import laplacian  # noqa: E402

warnings.simplefilter("ignore")
OUT = os.path.join(ROOT, "tests", "golden", "methods_subset.npz")
STRIDE, CHUNK = 10, 32
CASES = [("NNLS", "I", "brute-force", 60), ("L_curve", "I", "spline", 60), ("BayesReg", "I", "spline", 60),
         ("X2", "L2", "spline", 60), ("T2SPARC", "InvT2", "spline", 96), ("GCV", "L2", "brute-force", 60)]
_G = {}


def _fa(args):
    lo, fam, npc = args
    R = _G["R"]
    sig = _G["sig"][lo:lo + CHUNK]
    nx = sig.shape[0]
    ok = np.ones(nx)
    d = _G["dic"][npc]
    if fam == "spline":
        FA, idx, KM, _ = R["fa"].fitting_slice_FA_spline_method(d["lr"], d["hr273"], sig, ok, _G["a15"], nx, _G["a273"])
    else:
        FA, idx, KM, _ = R["fa"].fitting_slice_FA_brute_force(ok, sig, nx, d["hr91"], _G["a91"])
    return lo, idx


def _t2(args):
    lo, method, rm, fam, npc = args
    R = _G["R"]
    sig = _G["sig"][lo:lo + CHUNK]
    nx = sig.shape[0]
    idx = _G["fa"][(fam, npc)][lo:lo + nx]
    d = _G["dic"][npc]
    Dic = d["hr273"] if fam == "spline" else d["hr91"]
    L = laplacian(R, rm, d["T2s"])
    f, s, reg = R["motor"].fitting_slice_T2(np.ones(nx), sig, idx, nx, Dic, _G["lam"], npc, 32, method, L, None)
    return lo, f, reg


def main():
    t0 = time.time()
    R = ref_shim.load_reference()
    g2 = np.load(os.path.join(ROOT, "tests", "golden", "config2_subset.npz"))
    sig = np.ascontiguousarray(g2["sig"][::STRIDE])
    S = sig.shape[0]
    a273, a91, a15 = np.linspace(90, 180, 273), np.linspace(90, 180, 91), np.linspace(90, 180, 15)
    lam = np.zeros(50)
    lam[1:] = np.logspace(-8, 1, 49)
    dic = {}
    for npc in (60, 96):
        T2s = np.logspace(np.log10(10.0), np.log10(2000.0), npc)
        T1s = 1000.0 * np.ones(npc)
        mk = lambda a: R["epg"].create_Dic_3D(npc, T2s, T1s, 32, 10.0, a, 1000.0)
        dic[npc] = dict(T2s=T2s, hr273=mk(a273), lr=mk(a15), hr91=mk(a91) if npc == 60 else None)
    _G.update(R=R, sig=sig, a273=a273, a91=a91, a15=a15, lam=lam, dic=dic, fa={})
    print("dictionaries %.0f s" % (time.time() - t0), flush=True)
    out = dict(stride=STRIDE, n=S)
    ctx = mp.get_context("fork")
    for fam, npc in sorted({(c[2], c[3]) for c in CASES}):
        idx_all = np.zeros(S)
        with ctx.Pool(os.cpu_count()) as pool:
            for lo, idx in pool.imap_unordered(_fa, [(lo, fam, npc) for lo in range(0, S, CHUNK)]):
                idx_all[lo:lo + len(idx)] = idx
        _G["fa"][(fam, npc)] = idx_all
        out["fa_%s_%d" % (fam, npc)] = idx_all.astype(np.int16)
        print("FA", fam, npc, "%.0f s" % (time.time() - t0), flush=True)
    for method, rm, fam, npc in CASES:
        f_all = np.zeros((S, npc))
        reg_all = np.zeros(S)
        with ctx.Pool(os.cpu_count()) as pool:
            for lo, f, reg in pool.imap_unordered(_t2, [(lo, method, rm, fam, npc) for lo in range(0, S, CHUNK)]):
                f_all[lo:lo + len(reg)], reg_all[lo:lo + len(reg)] = f, reg
        key = "%s_%s" % (method, rm)
        sup = f_all > 0
        out[key + "_reg"] = reg_all
        out[key + "_support"] = np.packbits(sup, axis=1)
        out[key + "_fnz"] = f_all[sup]
        print(key, "%.0f s, mean support %.1f" % (time.time() - t0, sup.sum(1).mean()), flush=True)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%.1f MB)" % (OUT, os.path.getsize(OUT) / 1e6))


if __name__ == "__main__":
    main()
