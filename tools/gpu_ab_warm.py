"""A/B on the GPU: cold-started (reference path) vs warm-started lambda search over a full volume."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multicomponent_t2_toolbox_b200 import batched
from multicomponent_t2_toolbox_b200.phantom import make_phantom
shape = tuple(int(x) for x in os.environ.get("SHAPE", "96,96,60").split(","))
method = os.environ.get("METHOD", "X2"); rm = os.environ.get("RM", "I")
ph = make_phantom(shape, seed=2, fa_mode="b1", backend="gpu")
sig = torch.as_tensor(ph["data"].reshape(-1, 32)).cuda()
plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline")
fa = plan.fa_fit(sig)
res = {}
for name, flags in (("cold", 4), ("warm", 0)):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        out = plan.t2_fit(sig, fa["fa_index"], flags=flags)
        torch.cuda.synchronize(); dt = time.time() - t0
    res[name] = {k: v.clone() for k, v in out.items()}
    res[name + "_ms"] = dt * 1e3
c, w = res["cold"], res["warm"]
sup = ((c["fsol"] > 0) != (w["fsol"] > 0)).any(dim=1)
scale = c["fsol"].abs().max(dim=1).values.clamp_min(1e-300)
rel = (c["fsol"] - w["fsol"]).abs().max(dim=1).values / scale
regrel = (c["reg"] - w["reg"]).abs() / c["reg"].abs().clamp_min(1e-300)
rep = dict(method=method, rm=rm, V=int(sig.shape[0]), cold_ms=res["cold_ms"], warm_ms=res["warm_ms"],
           support_mismatch_voxels=int(sup.sum()), rel_spectrum_max=float(rel.max()),
           rel_spectrum_gt_1e6=int((rel > 1e-6).sum()), reg_rel_max=float(regrel.max()), reg_rel_gt_1e6=int((regrel > 1e-6).sum()),
           mwf_abs_max=float((c["maps"][:, 0] - w["maps"][:, 0]).abs().max()),
           status_cold=int((c["status"] != 0).sum()), status_warm=int((w["status"] != 0).sum()))
print(json.dumps(rep))
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "ab_warm_%s_%s.json" % (method, rm)), "w"), indent=1)

# which path does the CPU oracle side with on the discrepant voxels?
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import met2_oracle as O
bad = torch.nonzero((rel > 1e-6) | sup).flatten().cpu().numpy()
rng = np.random.default_rng(1)
extra = rng.choice(sig.shape[0], 40, replace=False)
Dic = plan.dict_hr.to_reference_layout()
sig_h = ph["data"].reshape(-1, 32)
idx_all = fa["fa_index"].cpu().numpy()
for label, vox in (("discrepant", bad), ("random", extra)):
    if len(vox) == 0:
        continue
    f_ref, s_ref, reg_ref = O.fitting_slice_T2(np.ones(len(vox)), sig_h[vox], idx_all[vox].astype(float), len(vox), Dic,
                                               plan.lambda_reg, plan.npc, 32, method, plan.Laplac)
    for name in ("cold", "warm"):
        f = res[name]["fsol"][torch.as_tensor(vox).cuda()].cpu().numpy()
        r = np.abs(f - f_ref).max(1) / np.abs(f_ref).max(1)
        rg = res[name]["reg"][torch.as_tensor(vox).cuda()].cpu().numpy()
        print(label, name, "n=%d" % len(vox), "n(rel>1e-6)=%d" % (r > 1e-6).sum(), "max rel %.2e" % r.max(),
              "support mismatches %d" % np.any((f > 0) != (f_ref > 0), 1).sum(), "max reg rel %.2e" % np.max(np.abs(rg - reg_ref) / np.abs(reg_ref)))
