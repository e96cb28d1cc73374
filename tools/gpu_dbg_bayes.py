import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, met2_oracle as O
from multicomponent_t2_toolbox_b200 import batched
from multicomponent_t2_toolbox_b200.phantom import make_phantom
method, rm = sys.argv[1], sys.argv[2]
ph = make_phantom((16, 16, 4), seed=1); sig = ph["data"].reshape(-1, 32)[:160]
plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline", npc=60)
fa, t2 = plan.fit(sig)
Dic = plan.dict_hr.to_reference_layout()
idx = fa["fa_index"].cpu().numpy().astype(float)
V = len(sig); ok = np.ones(V)
f_ref, s_ref, reg_ref = O.fitting_slice_T2(ok, sig, idx, V, Dic, plan.lambda_reg, 60, 32, method, plan.Laplac)
f = t2["fsol"].cpu().numpy(); reg = t2["reg"].cpu().numpy()
rel = np.abs(f - f_ref).max(1) / np.abs(f_ref).max(1)
rr = np.abs(reg - reg_ref) / np.abs(reg_ref)
maps = t2["maps"].cpu().numpy()
mref = np.array([O.voxel_metrics(f_ref[v], plan.T2s, plan.ind_m, plan.ind_t, plan.ind_csf) for v in range(V)])
print("support mismatch", int(np.any((f > 0) != (f_ref > 0), 1).sum()))
print("rel spectrum: max %.3e, n>1e-6: %d, n>1e-4: %d" % (rel.max(), (rel > 1e-6).sum(), (rel > 1e-4).sum()))
print("reg rel: max %.3e n>1e-6 %d" % (rr.max(), (rr > 1e-6).sum()))
print("maps abs max", np.abs(maps - mref).max(0))
w = np.argsort(-rr)[:5]
for v in w: print(v, reg[v], reg_ref[v], rel[v], t2["status"][v].item())
