import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, met2_oracle as O
from multicomponent_t2_toolbox_b200 import batched
from multicomponent_t2_toolbox_b200.phantom import make_phantom
rm = sys.argv[1]; NV = 96
ph = make_phantom((16, 16, 4), seed=1); sig = ph["data"].reshape(-1, 32)[:NV]
plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="GCV", reg_matrix=rm, FA_method="spline", npc=60)
fa = plan.fa_fit(sig)
Dic = plan.dict_hr.to_reference_layout(); idx = fa["fa_index"].cpu().numpy()
L = plan.Laplac
for lam in [1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 0.1, 1.0, 3.8197]:
    t2 = plan.t2_fit(sig, fa["fa_index"], flags=8, lambda_fixed=lam)
    og = t2["reg"].cpu().numpy()
    oref = np.zeros(NV); kk = np.zeros(NV, int)
    for v in range(NV):
        D = np.ascontiguousarray(Dic[:, :, idx[v]]); M = sig[v] / sig[v, 0]
        Maug = np.concatenate((M, np.zeros(60)))
        oref[v] = O.obj_nnls_gcv(lam, D, L, Maug, 32, np.eye(32))
        f, _ = O.nnls(np.concatenate((D, np.sqrt(lam) * L)), Maug); kk[v] = (f > 0).sum()
    d = np.abs(og - oref)
    print("lam=%g k=%d..%d  frac(|dobj|<1e-8)=%.3f  median %.1e max %.2e   sample gpu %.10f ref %.10f" % (lam, kk.min(), kk.max(), (d < 1e-8).mean(), np.median(d), d.max(), og[0], oref[0]))
