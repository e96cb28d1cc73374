"""Which kernel family does the reference side with?  L-curve (algorithms.py:88-113) over the whole config-2 volume in the
reduced echo space and in the Gram domain; on the voxels where the chosen corner differs (and on a random sample of the
others) the oracle is run on the CPU and the three lambdas are compared.

    gpurun --timeout 900 -- 'timeout 800 python tools/gpu_lcurve_arbiter.py > gpurun_out/lcurve_arbiter.log 2>&1'
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import met2_oracle as O  # noqa: E402
from multicomponent_t2_toolbox_b200 import batched  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

ph = make_phantom((96, 96, 60), seed=2, fa_mode="b1", backend="gpu")
sig_h = ph["data"].reshape(-1, 32)
sig = torch.as_tensor(sig_h).cuda()
rep = {}
nmax = int(os.environ.get("NMAX", "120"))
for rm in os.environ.get("RMS", "InvT2,I").split(","):
    out = {}
    for name, kw in (("echo", {}), ("gram", dict(echo_space=False))):
        plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="L_curve", reg_matrix=rm, FA_method="spline", **kw)
        if name == "echo":
            fa = plan.fa_fit(sig)
        t2 = plan.t2_fit(sig, fa["fa_index"])
        out[name] = t2["reg"].cpu().numpy()
    idx = fa["fa_index"].cpu().numpy()
    Dic = plan.dict_hr.to_reference_layout()
    differ = np.nonzero(out["echo"] != out["gram"])[0]
    rng = np.random.default_rng(0)
    same = rng.choice(np.nonzero(out["echo"] == out["gram"])[0], 200, replace=False)
    r = dict(voxels=int(len(idx)), corner_differs=int(len(differ)))
    for label, pick in (("differing", differ[:nmax]), ("agreeing_sample", same)):
        c = dict(n=int(len(pick)), echo_is_reference=0, gram_is_reference=0, neither=0, both=0)
        lows = []
        for v in pick:
            D = np.ascontiguousarray(Dic[:, :, idx[v]])
            _, _, reg_ref = O.t2_fit_voxel(sig_h[v], D, "L_curve", plan.Laplac, plan.lambda_reg)
            e, g = out["echo"][v] == reg_ref, out["gram"][v] == reg_ref
            c["both" if (e and g) else "echo_is_reference" if e else "gram_is_reference" if g else "neither"] += 1
            if label == "differing":
                lows.append((float(out["echo"][v]), float(out["gram"][v]), float(reg_ref)))
        r[label] = c
        if lows:
            r["examples_echo_gram_reference"] = lows[:12]
    rep["L_curve_" + rm] = r
    print(json.dumps({"L_curve_" + rm: r}), flush=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "lcurve_arbiter.json"), "w"), indent=1)
