"""TEST INFRASTRUCTURE ONLY — writes tests/golden/methods_invt2_subset.npz from OUTPUTS OF THE UNMODIFIED REFERENCE.

The regularisation matrix InvT2 (motor/motor_recon_met2_real_data.py:263-269) at scale: the 2 048 voxels of
tests/golden/methods_subset.npz (every 10th voxel of config2_subset.npz, 60 bins, spline FA indices taken from that file)
fitted by the reference's own row worker (motor...:113) with

    X2-InvT2      L_curve-InvT2      BayesReg-InvT2      BayesReg-InvT2 on the signals * (1 + 1e-13 N(0,1))

The last one measures the reference's own reproducibility (flat evidence, absolute Brent xtol): the bound the GPU test
holds the BayesReg kernels to, as tests/golden/config4_subset.npz does at config-4 sizes.  InvT2 is where the
echo-space formulation is worst conditioned (columns scaled by T2: up to 2000), hence a fixture of its own.

    python oracle/make_golden_methods_invt2.py          (~15 min on 8 cores)
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
import make_golden_methods as M  # noqa: E402

warnings.simplefilter("ignore")
OUT = os.path.join(ROOT, "tests", "golden", "methods_invt2_subset.npz")
CASES = [("X2", "InvT2"), ("L_curve", "InvT2"), ("BayesReg", "InvT2")]


def main():
    t0 = time.time()
    R = ref_shim.load_reference()
    g2 = np.load(os.path.join(ROOT, "tests", "golden", "config2_subset.npz"))
    gm = np.load(os.path.join(ROOT, "tests", "golden", "methods_subset.npz"))
    stride = int(gm["stride"])
    sig = np.ascontiguousarray(g2["sig"][::stride])
    S = sig.shape[0]
    a273 = np.linspace(90, 180, 273)
    lam = np.zeros(50)
    lam[1:] = np.logspace(-8, 1, 49)
    T2s = np.logspace(np.log10(10.0), np.log10(2000.0), 60)
    dic = {60: dict(T2s=T2s, hr273=R["epg"].create_Dic_3D(60, T2s, 1000.0 * np.ones(60), 32, 10.0, a273, 1000.0))}
    M._G.update(R=R, sig=sig, a273=a273, lam=lam, dic=dic, fa={("spline", 60): gm["fa_spline_60"].astype(np.float64)})
    out = dict(stride=stride, n=S)
    ctx = mp.get_context("fork")
    rng = np.random.default_rng(20260)
    runs = [(m, rm, "") for m, rm in CASES] + [("BayesReg", "InvT2", "p")]
    for method, rm, tag in runs:
        if tag == "p":
            M._G["sig"] = sig * (1.0 + 1e-13 * rng.standard_normal(sig.shape))
        f_all = np.zeros((S, 60))
        reg_all = np.zeros(S)
        with ctx.Pool(os.cpu_count()) as pool:
            for lo, f, reg in pool.imap_unordered(M._t2, [(lo, method, rm, "spline", 60) for lo in range(0, S, M.CHUNK)]):
                f_all[lo:lo + len(reg)], reg_all[lo:lo + len(reg)] = f, reg
        key = "%s_%s%s" % (method, rm, tag)
        sup = f_all > 0
        out[key + "_reg"] = reg_all
        out[key + "_support"] = np.packbits(sup, axis=1)
        out[key + "_fnz"] = f_all[sup]
        print(key, "%.0f s, mean support %.1f" % (time.time() - t0, sup.sum(1).mean()), flush=True)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%.1f MB)" % (OUT, os.path.getsize(OUT) / 1e6))


if __name__ == "__main__":
    main()
