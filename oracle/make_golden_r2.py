"""TEST INFRASTRUCTURE ONLY — writes tests/golden/plain_nnls_wide.npz and tests/golden/config4_subset.npz from OUTPUTS OF
THE UNMODIFIED REFERENCE (imported read-only through oracle/ref_shim.py; runs only where /root/reference exists).

plain_nnls_wide.npz — plain Lawson-Hanson NNLS (intravoxel_algorithms/algorithms.py:55-82) on the WIDE T2 grids, where
    the long-T2 columns are nearly collinear ("Plain NNLS on the 96-bin grid"): the 20 480 config-2
    voxels of tests/golden/config2_subset.npz (signals are not stored twice) against the 60-, 96- (T2SPARC) and 100-bin
    dictionaries at the fixture's flip-angle index, through the reference's row worker fitting_slice_T2(..., 'NNLS')
    (motor/motor_recon_met2_real_data.py:113), and the reference's spline FA search on the 96-bin dictionaries
    (flip_angle_algorithms/fa_estimation.py:35) — the FA stage of every T2SPARC reconstruction.
        nnls60_support / nnls60_fnz, nnls96_*, nnls100_*, fa96_idx [S] int16, fa96_km [S]

config4_subset.npz — BASELINE.json configs[3] sizes: nTE = 48 (tau 8 ms), 100 T2 bins, brute-force FA over 91 angles,
    BayesReg + InvT2 (intravoxel_algorithms/bayesian_interpolation.py:84-126) on 2 048 seeded phantom voxels:
        sig [S, 48], fa_idx, km, nnls_* (plain NNLS at that index), bayes_reg / bayes_support / bayes_fnz,
    and the SAME BayesReg fit of the signals perturbed by a 1e-13 relative factor (bayesp_*): the reference's own
    reproducibility at rounding level, which bounds what any re-implementation can be held to (flat evidence curve,
    Brent xtol = 1e-5 absolute).

    python oracle/make_golden_r2.py
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
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

warnings.simplefilter("ignore")
GOLD = os.path.join(ROOT, "tests", "golden")
CHUNK = 64
_G = {}


def _nnls_wide(args):
    lo, npc = args
    R = _G["R"]
    sig = _G["sig"][lo:lo + CHUNK]
    nx = sig.shape[0]
    f, s, reg = R["motor"].fitting_slice_T2(np.ones(nx), sig, _G["fa"][lo:lo + nx], nx, _G["dic"][npc], _G["lam"], npc,
                                            sig.shape[1], "NNLS", np.eye(npc), None)
    return lo, f


def _fa96(lo):
    R = _G["R"]
    sig = _G["sig"][lo:lo + CHUNK]
    nx = sig.shape[0]
    FA, idx, KM, _ = R["fa"].fitting_slice_FA_spline_method(_G["dic96lr"], _G["dic"][96], sig, np.ones(nx), _G["a15"], nx,
                                                            _G["a273"])
    return lo, idx, KM


def _c4_fa(lo):
    R = _G["R"]
    sig = _G["sig4"][lo:lo + CHUNK]
    nx = sig.shape[0]
    FA, idx, KM, _ = R["fa"].fitting_slice_FA_brute_force(np.ones(nx), sig, nx, _G["dic4"], _G["a91"])
    return lo, idx, KM


def _c4_t2(args):
    lo, method, which = args
    R = _G["R"]
    sig = _G[which][lo:lo + CHUNK]
    nx = sig.shape[0]
    L = _G["L4"] if method == "BayesReg" else np.eye(100)
    f, s, reg = R["motor"].fitting_slice_T2(np.ones(nx), sig, _G["fa4"][lo:lo + nx], nx, _G["dic4"], _G["lam"], 100, 48,
                                            method, L, None)
    return lo, f, reg


def _pack(out, key, f):
    sup = f > 0
    out[key + "_support"] = np.packbits(sup, axis=1)
    out[key + "_fnz"] = f[sup]


def main():
    t0 = time.time()
    R = ref_shim.load_reference()
    ctx = mp.get_context("fork")
    lam = np.zeros(50)
    lam[1:] = np.logspace(-8, 1, 49)
    a273, a91, a15 = np.linspace(90, 180, 273), np.linspace(90, 180, 91), np.linspace(90, 180, 15)
    _G.update(R=R, lam=lam, a273=a273, a91=a91, a15=a15)

    # ------------------------------------------------------------------ plain NNLS on the 96 / 100-bin grids
    g2 = np.load(os.path.join(GOLD, "config2_subset.npz"))
    sig = np.ascontiguousarray(g2["sig"])
    S = sig.shape[0]
    dic = {}
    for npc in (60, 96, 100):
        T2s = np.logspace(np.log10(10.0), np.log10(2000.0), npc)
        dic[npc] = R["epg"].create_Dic_3D(npc, T2s, 1000.0 * np.ones(npc), 32, 10.0, a273, 1000.0)
    T2s96 = np.logspace(np.log10(10.0), np.log10(2000.0), 96)
    _G.update(sig=sig, fa=g2["fa_idx"].astype(np.float64), dic=dic,
              dic96lr=R["epg"].create_Dic_3D(96, T2s96, 1000.0 * np.ones(96), 32, 10.0, a15, 1000.0))
    print("dictionaries %.0f s" % (time.time() - t0), flush=True)
    out = dict(n=S)
    for npc in (60, 96, 100):
        f_all = np.zeros((S, npc))
        with ctx.Pool(os.cpu_count()) as pool:
            for lo, f in pool.imap_unordered(_nnls_wide, [(lo, npc) for lo in range(0, S, CHUNK)]):
                f_all[lo:lo + len(f)] = f
        _pack(out, "nnls%d" % npc, f_all)
        print("NNLS %d bins %.0f s, mean support %.2f" % (npc, time.time() - t0, (f_all > 0).sum(1).mean()), flush=True)
    idx_all, km_all = np.zeros(S), np.zeros(S)
    with ctx.Pool(os.cpu_count()) as pool:
        for lo, idx, KM in pool.imap_unordered(_fa96, range(0, S, CHUNK)):
            idx_all[lo:lo + len(idx)], km_all[lo:lo + len(idx)] = idx, KM
    out["fa96_idx"] = idx_all.astype(np.int16)
    out["fa96_km"] = km_all
    print("FA spline 96 bins %.0f s" % (time.time() - t0), flush=True)
    p = os.path.join(GOLD, "plain_nnls_wide.npz")
    np.savez_compressed(p, **out)
    print("wrote %s (%.1f MB)" % (p, os.path.getsize(p) / 1e6), flush=True)

    # ------------------------------------------------------------------ config-4 sizes
    S4 = 2048
    ph = make_phantom((64, 64, 2), n_echoes=48, tau=8.0, TR=1000.0, seed=4, fa_mode="b1")
    allsig = ph["data"].reshape(-1, 48)
    pick = np.sort(np.random.default_rng(4).choice(allsig.shape[0], S4, replace=False))
    sig4 = np.ascontiguousarray(allsig[pick])
    rel = 1.0 + 1e-13 * np.random.default_rng(44).standard_normal(sig4.shape)
    T2s100 = np.logspace(np.log10(10.0), np.log10(2000.0), 100)
    _G.update(sig4=sig4, sig4p=sig4 * rel, L4=laplacian(R, "InvT2", T2s100),
              dic4=R["epg"].create_Dic_3D(100, T2s100, 1000.0 * np.ones(100), 48, 8.0, a91, 1000.0))
    fa4, km4 = np.zeros(S4), np.zeros(S4)
    with ctx.Pool(os.cpu_count()) as pool:
        for lo, idx, KM in pool.imap_unordered(_c4_fa, range(0, S4, CHUNK)):
            fa4[lo:lo + len(idx)], km4[lo:lo + len(idx)] = idx, KM
    _G["fa4"] = fa4
    print("config 4 FA %.0f s" % (time.time() - t0), flush=True)
    out4 = dict(sig=sig4, rel_perturbation=rel, fa_idx=fa4.astype(np.int16), km=km4)
    for key, method, which in (("nnls", "NNLS", "sig4"), ("bayes", "BayesReg", "sig4"), ("bayesp", "BayesReg", "sig4p")):
        f_all, reg_all = np.zeros((S4, 100)), np.zeros(S4)
        with ctx.Pool(os.cpu_count()) as pool:
            for lo, f, reg in pool.imap_unordered(_c4_t2, [(lo, method, which) for lo in range(0, S4, CHUNK)]):
                f_all[lo:lo + len(reg)], reg_all[lo:lo + len(reg)] = f, reg
        _pack(out4, key, f_all)
        out4[key + "_reg"] = reg_all
        print("config 4 %s %.0f s, mean support %.2f" % (key, time.time() - t0, (f_all > 0).sum(1).mean()), flush=True)
    p = os.path.join(GOLD, "config4_subset.npz")
    np.savez_compressed(p, **out4)
    print("wrote %s (%.1f MB)" % (p, os.path.getsize(p) / 1e6), flush=True)


if __name__ == "__main__":
    main()
