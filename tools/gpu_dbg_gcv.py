import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, met2_oracle as O
from multicomponent_t2_toolbox_b200 import batched
from multicomponent_t2_toolbox_b200.phantom import make_phantom
method, rm = "GCV", sys.argv[1]
NV = int(sys.argv[2]) if len(sys.argv) > 2 else 160
ph = make_phantom((16, 16, 4), seed=1); sig = ph["data"].reshape(-1, 32)[:NV]
plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline", npc=60)
import time
fa = plan.fa_fit(sig); torch.cuda.synchronize(); t0 = time.time()
t2 = plan.t2_fit(sig, fa["fa_index"]); torch.cuda.synchronize(); print("gpu t2 s", time.time() - t0)
Dic = plan.dict_hr.to_reference_layout()
idx = fa["fa_index"].cpu().numpy().astype(float)
V = len(sig); ok = np.ones(V)
f_ref, s_ref, reg_ref = O.fitting_slice_T2(ok, sig, idx, V, Dic, plan.lambda_reg, 60, 32, method, plan.Laplac)
rng = np.random.default_rng(0)
sig2 = sig * (1 + 1e-13 * rng.standard_normal(sig.shape))
f_p, s_p, reg_p = O.fitting_slice_T2(ok, sig2, idx, V, Dic, plan.lambda_reg, 60, 32, method, plan.Laplac)
f = t2["fsol"].cpu().numpy(); reg = t2["reg"].cpu().numpy()
def stats(fa_, ra_, name):
    rel = np.abs(fa_ - f_ref).max(1) / np.abs(f_ref).max(1)
    rr = np.abs(ra_ - reg_ref) / np.abs(reg_ref)
    mw = lambda ff: np.array([O.voxel_metrics(ff[v], plan.T2s, plan.ind_m, plan.ind_t, plan.ind_csf)[0] for v in range(V)])
    dm = np.abs(mw(fa_) - mw(f_ref))
    print(name, "support mismatch %d/%d" % (np.any((fa_ > 0) != (f_ref > 0), 1).sum(), V),
          "frac(rel spectrum<1e-6)=%.3f" % (rel < 1e-6).mean(), "frac(reg rel<1e-6)=%.3f" % (rr < 1e-6).mean(),
          "max rel %.2e" % rel.max(), "frac(|dMWF|<1e-4)=%.3f max %.2e" % ((dm < 1e-4).mean(), dm.max()))
stats(f, reg, "GPU vs oracle      ")
stats(f_p, reg_p, "oracle(1e-13 pert) ")
print("status or", int(np.bitwise_or.reduce(t2["status"].cpu().numpy())))
print("reg gpu", reg[:6]); print("reg ref", reg_ref[:6])
