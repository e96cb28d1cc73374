"""Seeded re-run of the reference's Monte-Carlo evaluation on the GPU path: the recipe of
scripts_synthetic_data_evaluation/Paper_Comparison/evaluate_all_methods_two_lobes_SNR*.py — two Gaussian lobes on a
1 000-point T2 grid pushed through the EPG model at a continuous flip angle, Rician noise, 60-bin dictionary at TR = 3000
with 91 flip angles, brute-force FA, L-curve grid logspace(1e-8, 100, 49) — for the ten methods of the paper tables
(NNLS, X2 / L-curve / GCV x I / L1 / L2), N voxels per SNR band.  The only numbers the reference itself holds for this
path are those tables (Results/SNRs_*/All_methods_10000iters/table_{errors,regularization}.txt): MWF mean absolute error
and mean lambda per method.  They were produced with unseeded draws, so the comparison is statistical.

    python tools/montecarlo.py [N=10000] [seed=0]  ->  gpurun_out/montecarlo.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (MWF MAE, mean lambda, std lambda) per method from the reference's tables, lines 3-12
REFERENCE = {
    "50_150": {"NNLS": (0.0679834, 0.0, 0.0), "X2-I": (0.0548569, 0.00257832, 0.00448045),
               "X2-L1": (0.0557831, 0.0224889, 0.0656093), "X2-L2": (0.0556009, 0.179334, 0.642022),
               "Lcurve-I": (0.0543839, 0.00543128, 0.00587748), "Lcurve-L1": (0.0568595, 0.127913, 0.132735),
               "Lcurve-L2": (0.0558122, 0.264592, 0.267592), "GCV-I": (0.0581128, 0.000774211, 0.00273344),
               "GCV-L1": (0.0587813, 0.107992, 0.226097), "GCV-L2": (0.0598534, 0.865506, 1.3048)},
    "150_300": {"NNLS": (0.0517479, 0.0, 0.0), "X2-I": (0.0432803, 0.000462573, 0.000729411),
                "X2-L1": (0.0445168, 0.00238808, 0.00576503), "X2-L2": (0.0444649, 0.0117339, 0.039724),
                "Lcurve-I": (0.0497683, 0.000866545, 0.000762231), "Lcurve-L1": (0.0549936, 0.0193198, 0.0202572),
                "Lcurve-L2": (0.0549595, 0.0666901, 0.0742864), "GCV-I": (0.0433162, 9.71771e-05, 0.000474276),
                "GCV-L1": (0.0464862, 0.020297, 0.0584989), "GCV-L2": (0.0525414, 0.287106, 0.629287)},
}
BANDS = {"50_150": (50.0, 150.0), "150_300": (150.0, 300.0)}
METHODS = [("NNLS", "NNLS", "I"), ("X2-I", "X2", "I"), ("X2-L1", "X2", "L1"), ("X2-L2", "X2", "L2"),
           ("Lcurve-I", "L_curve", "I"), ("Lcurve-L1", "L_curve", "L1"), ("Lcurve-L2", "L_curve", "L2"),
           ("GCV-I", "GCV", "I"), ("GCV-L1", "GCV", "L1"), ("GCV-L2", "GCV", "L2")]


def simulate(N, snr_lo, snr_hi, rng, T2s):
    """Signals [N, 32] and the true myelin fraction of the 60-bin version of every voxel's distribution
    (evaluate_all_methods_two_lobes_SNR50_150.py:156-176, 373-431)."""
    from scipy.stats import norm

    from multicomponent_t2_toolbox_b200.phantom import _epg_signal_batch_gpu
    mwf = rng.uniform(0.05, 0.25, N)
    t2m = rng.uniform(15.0, 35.0, N)
    t2ie = rng.uniform(60.0, 90.0, N)
    fa = rng.uniform(90.0, 180.0, N)
    snr = rng.uniform(snr_lo, snr_hi, N)
    sig_m = rng.uniform(1.0, 3.0, N)
    sig_ie = rng.uniform(6.0, 12.0, N)
    grid, dgrid = np.linspace(1.0, 300.0, 1000, retstep=True)
    TR, T1, Km, nte, tau = 3000.0, 1000.0, 1000.0, 32, 10.0
    # bin edges of the low-resolution pdf (:408-426)
    mid = T2s[:-1] + np.diff(T2s) / 2.0
    bins = np.searchsorted(mid, grid, side="right")        # grid point -> T2 bin
    ind_m = T2s <= 40.0
    signals = np.empty((N, nte))
    true_fm = np.empty(N)
    chunk = 500
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        dist = (mwf[s:e, None] * norm.pdf(grid[None, :], t2m[s:e, None], sig_m[s:e, None])
                + (1.0 - mwf[s:e, None]) * norm.pdf(grid[None, :], t2ie[s:e, None], sig_ie[s:e, None]))
        dist /= dist.sum(axis=1, keepdims=True)
        m = e - s
        curves = _epg_signal_batch_gpu(nte, tau, np.full(m * 1000, T1), np.tile(grid, m), np.repeat(fa[s:e], 1000))
        curves = curves.reshape(m, 1000, nte) * (1.0 - np.exp(-TR / T1))
        clean = Km * np.einsum("vk,vke->ve", dist, curves)
        sd = clean[:, :1] / snr[s:e, None]
        n1 = rng.normal(0.0, 1.0, clean.shape) * sd
        n2 = rng.normal(0.0, 1.0, clean.shape) * sd
        signals[s:e] = np.sqrt((clean + n1) ** 2 + n2 ** 2)
        d2 = np.zeros((m, len(T2s)))
        np.add.at(d2, (np.repeat(np.arange(m), 1000), np.tile(bins, m)), (dist * dgrid).ravel())
        d2 /= d2.sum(axis=1, keepdims=True)
        true_fm[s:e] = d2[:, ind_m].sum(axis=1)
    return signals, true_fm


def run(N=10000, seed=0, device=None):
    import torch

    from multicomponent_t2_toolbox_b200 import batched
    lam = np.zeros(50)
    lam[1:] = np.logspace(-8, 2, 49)                        # the paper scripts' grid (the CLI path stops at 10)
    out = {"N": int(N), "seed": int(seed), "bands": {}}
    for bi, (band, (lo, hi)) in enumerate(BANDS.items()):
        rng = np.random.default_rng(seed + 1000 * bi)
        plans = {}
        for name, method, rm in METHODS:
            plans[name] = batched.Met2Plan(32, 10.0, 3000.0, reg_method=method, reg_matrix=rm, FA_method="brute-force",
                                           lambda_reg=lam, device=device)
        T2s = plans["NNLS"].T2s
        t0 = time.time()
        sig, true_fm = simulate(N, lo, hi, rng, T2s)
        t_sim = time.time() - t0
        d_sig = torch.as_tensor(sig).to(plans["NNLS"].dev)
        fa = plans["NNLS"].fa_fit(d_sig)
        res = {}
        t0 = time.time()
        for name, method, rm in METHODS:
            t2 = plans[name].t2_fit(d_sig, fa["fa_index"], flags=batched.REG_IS_LAMBDA)
            mwf = t2["maps"][:, 0].cpu().numpy()
            reg = t2["reg"].cpu().numpy()
            ref = REFERENCE[band][name]
            res[name] = {"MAE": float(np.mean(np.abs(mwf - true_fm))), "mean_lambda": float(reg.mean()),
                         "std_lambda": float(reg.std()), "status_nonzero": int((t2["status"] != 0).sum()),
                         "reference_MAE": ref[0], "reference_mean_lambda": ref[1], "reference_std_lambda": ref[2]}
        torch.cuda.synchronize()
        out["bands"][band] = {"methods": res, "simulate_s": t_sim, "fit_s": time.time() - t0,
                              "fa_mean_abs_err_deg": None}
    return out


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rec = run(N, seed)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "montecarlo.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    for band, b in rec["bands"].items():
        print("SNR", band, "simulate %.1f s, ten fits %.2f s" % (b["simulate_s"], b["fit_s"]))
        for name, r in b["methods"].items():
            print("  %-10s MAE %.4f (reference %.4f, %+5.1f %%)   mean lambda %.5g (reference %.5g)" % (
                name, r["MAE"], r["reference_MAE"], 100 * (r["MAE"] / r["reference_MAE"] - 1), r["mean_lambda"],
                r["reference_mean_lambda"]))
