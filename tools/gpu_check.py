"""Development check (run under gpurun): CUDA path vs the CPU oracle on a small phantom, with verbose diagnostics."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

import met2_oracle as O  # noqa: E402
from multicomponent_t2_toolbox_b200 import batched  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

NV = int(os.environ.get("NV", "96"))
report = {}


def spectrum_stats(f_gpu, f_ref):
    sup_bad = int(np.sum(np.any((f_gpu > 0) != (f_ref > 0), axis=1)))
    scale = np.abs(f_ref).max(axis=1) + 1e-300
    rel = np.abs(f_gpu - f_ref).max(axis=1) / scale
    return sup_bad, float(rel.max())


def main():
    ph = make_phantom((16, 16, 4), seed=1)
    sig = ph["data"].reshape(-1, 32)[:NV].copy()
    sig[5] = 0.0           # skipped voxel
    sig[7, 0] = 0.0        # M[0] == 0 -> FA fitted, T2 skipped
    t0 = time.time()
    # ---------------- dictionary
    plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="NNLS", reg_matrix="I", FA_method="brute-force")
    torch.cuda.synchronize()
    Dic = plan.dict_hr.to_reference_layout()
    sel = [0, 17, 45, 90]
    Dref = O.create_Dic_3D(60, plan.T2s, plan.T1s, 32, 10.0, plan.alpha_values[sel], 1000.0)
    report["dic_max_abs"] = float(np.abs(Dic[:, :, sel] - Dref).max())
    G = plan.dict_hr.G.cpu().numpy()
    report["gram_max_rel"] = float(np.abs(G[17] - Dic[:, :, 17].T @ Dic[:, :, 17]).max() / np.abs(G[17]).max())
    print("dictionary", report, flush=True)
    # ---------------- FA brute force
    fa = plan.fa_fit(sig)
    torch.cuda.synchronize()
    idx_g = fa["fa_index"].cpu().numpy()
    km_g = fa["km"].cpu().numpy()
    idx_r = np.zeros(NV, int); km_r = np.zeros(NV); fsum = 0
    for v in range(NV):
        if sig[v].sum() > 0:
            i, a, km, sse, f = O.compute_optimal_FA(sig[v], Dic, plan.alpha_values)
            idx_r[v], km_r[v] = i, km
            fsum = fsum + f
    report["fa_brute_idx_mismatch"] = int(np.sum(idx_g != idx_r))
    report["fa_brute_km_max_rel"] = float(np.max(np.abs(km_g - km_r) / (np.abs(km_r) + 1e-300)))
    report["fa_brute_fsum_max_rel"] = float(np.abs(fa["fsol_sum"].cpu().numpy() - fsum).max() / np.abs(fsum).max())
    report["fa_status"] = fa["status"].cpu().numpy()[:10].tolist()
    print("fa brute", {k: v for k, v in report.items() if k.startswith("fa_")}, flush=True)
    # ---------------- T2 plain NNLS on those indices
    t2 = plan.t2_fit(sig, fa["fa_index"])
    torch.cuda.synchronize()
    f_g = t2["fsol"].cpu().numpy()
    f_r = np.zeros_like(f_g); s_r = np.zeros((NV, 32)); reg_r = np.zeros(NV)
    lam = plan.lambda_reg
    ok = np.ones(NV, dtype=np.int64)
    f_r, s_r, reg_r = O.fitting_slice_T2(ok, sig, idx_r.astype(float), NV, Dic, lam, 60, 32, "NNLS", plan.Laplac)
    sb, rel = spectrum_stats(f_g, f_r)
    report["t2_nnls"] = dict(support_mismatch=sb, max_rel=rel,
                             est_max_rel=float(np.abs(t2["est_signal"].cpu().numpy() - s_r).max() / np.abs(s_r).max()),
                             status=t2["status"].cpu().numpy()[:10].tolist())
    maps_g = t2["maps"].cpu().numpy()
    maps_r = np.array([O.voxel_metrics(f_r[v], plan.T2s, plan.ind_m, plan.ind_t, plan.ind_csf) for v in range(NV)])
    report["t2_nnls"]["maps_max_abs"] = np.abs(maps_g - maps_r).max(axis=0).tolist()
    print("t2 nnls", report["t2_nnls"], flush=True)
    # ---------------- spline FA
    plan_s = batched.Met2Plan(32, 10.0, 1000.0, reg_method="X2", reg_matrix="I", FA_method="spline")
    Dic3 = plan_s.dict_hr.to_reference_layout()
    DicLR = plan_s.dict_lr.to_reference_layout()
    fa_s = plan_s.fa_fit(sig)
    torch.cuda.synchronize()
    idx_gs = fa_s["fa_index"].cpu().numpy()
    idx_rs = np.zeros(NV, int); xs_r = np.zeros(NV)
    for v in range(NV):
        if sig[v].sum() > 0:
            i, a, km, f, resid, xmin = O.spline_optimal_FA(sig[v], DicLR, Dic3, plan_s.alpha_spline, plan_s.alpha_values, True)
            idx_rs[v] = i; xs_r[v] = xmin
    report["fa_spline_idx_mismatch"] = int(np.sum(idx_gs != idx_rs))
    report["fa_spline_idx_pairs"] = [(int(a), int(b)) for a, b in zip(idx_gs, idx_rs) if a != b][:10]
    print("fa spline mismatch", report["fa_spline_idx_mismatch"], report["fa_spline_idx_pairs"], flush=True)
    # ---------------- regularised methods
    for method, rm in [("X2", "I"), ("X2", "L2"), ("X2", "L1"), ("X2", "InvT2"), ("L_curve", "I"), ("L_curve", "L2"),
                       ("T2SPARC", "InvT2")]:
        pl = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline")
        Dm = pl.dict_hr.to_reference_layout()
        tg0 = time.time()
        t2 = pl.t2_fit(sig, idx_rs.astype(np.int32))
        torch.cuda.synchronize()
        tg = time.time() - tg0
        f_g = t2["fsol"].cpu().numpy(); reg_g = t2["reg"].cpu().numpy()
        tr0 = time.time()
        f_r, s_r, reg_r = O.fitting_slice_T2(ok, sig, idx_rs.astype(float), NV, Dm, pl.lambda_reg, pl.npc, 32, method, pl.Laplac)
        tr = time.time() - tr0
        sb, rel = spectrum_stats(f_g, f_r)
        maps_g = t2["maps"].cpu().numpy()
        maps_r = np.array([O.voxel_metrics(f_r[v], pl.T2s, pl.ind_m, pl.ind_t, pl.ind_csf) for v in range(NV)])
        regdiff = np.abs(reg_g - reg_r) / (np.abs(reg_r) + 1e-300)
        report["t2_%s_%s" % (method, rm)] = dict(support_mismatch=sb, max_rel=rel, reg_max_rel=float(regdiff.max()),
                                                 reg_n_diff=int(np.sum(regdiff > 1e-6)), mwf_max_abs=float(np.abs(maps_g[:, 0] - maps_r[:, 0]).max()),
                                                 gpu_s=tg, cpu_s=tr, status_or=int(np.bitwise_or.reduce(t2["status"].cpu().numpy())))
        print(method, rm, report["t2_%s_%s" % (method, rm)], flush=True)
    report["total_s"] = time.time() - t0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as fh:
        json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
