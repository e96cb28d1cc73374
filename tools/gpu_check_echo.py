"""Round-2 entry point for the experimental echo-space X2 kernel (csrc/met2_t2_echo.cu, MET2_T2_FLAG_ECHO_SPACE = 64):
parity against the golden voxels of the unmodified reference, then an A/B against the default kernel over the full
config-2 volume (supports, spectra, k_est, time).  Run it under `timeout` on the GPU box:

    gpurun --timeout 600 -- 'timeout 300 python tools/gpu_check_echo.py > gpurun_out/echo.log 2>&1'

RM=InvT2 switches the regularisation matrix; SHAPE=48,48,30 shrinks the volume; METHOD=T2SPARC checks the fixed-lambda
echo-space kernel (t2_echo_tik_kernel; 96 bins) against the default kernel instead (no golden step)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multicomponent_t2_toolbox_b200 import batched  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

ECHO = 64
rm = os.environ.get("RM", "I")
method = os.environ.get("METHOD", "X2")
shape = tuple(int(x) for x in os.environ.get("SHAPE", "96,96,60").split(","))
plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline", echo_space=False)
rep = dict(rm=rm, method=method)

# ---- 1. golden voxels (20 480, fitted by the unmodified reference with X2-I)
if rm == "I" and method == "X2":
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "config2_subset.npz")))
    sup = np.unpackbits(g["support"], axis=1)[:, :60].astype(bool)
    f_ref = np.zeros(sup.shape)
    f_ref[sup] = g["f_nz"]
    sig = torch.as_tensor(g["sig"]).cuda()
    idx = torch.as_tensor(g["fa_idx"].astype(np.int32)).cuda()
    out = plan.t2_fit(sig, idx, flags=ECHO)
    torch.cuda.synchronize()
    f = out["fsol"].cpu().numpy()
    rep["golden"] = dict(voxels=int(len(f)), status_nonzero=int((out["status"] != 0).sum()),
                         support_disagreements=int(np.any((f > 0) != (f_ref > 0), axis=1).sum()),
                         spectrum_rel_max=float(np.max(np.abs(f - f_ref).max(1) / np.abs(f_ref).max(1))),
                         k_est_rel_max=float(np.max(np.abs(out["reg"].cpu().numpy() - g["reg"]) / np.abs(g["reg"]))))
    print(json.dumps(rep), flush=True)

# ---- 2. A/B over a volume
ph = make_phantom(shape, seed=2, fa_mode="b1", backend="gpu")
sig = torch.as_tensor(ph["data"].reshape(-1, 32)).cuda()
fa = plan.fa_fit(sig)
res = {}
for name, flags in (("default", 0), ("echo", ECHO)):
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = plan.t2_fit(sig, fa["fa_index"], flags=flags)
        e1.record()
        torch.cuda.synchronize()
    res[name] = {k: v.clone() for k, v in out.items()}
    rep[name + "_ms"] = e0.elapsed_time(e1)
d, e = res["default"], res["echo"]
scale = d["fsol"].abs().max(dim=1).values.clamp_min(1e-300)
rel = (d["fsol"] - e["fsol"]).abs().max(dim=1).values / scale
rep["ab"] = dict(voxels=int(sig.shape[0]), support_mismatch_voxels=int(((d["fsol"] > 0) != (e["fsol"] > 0)).any(dim=1).sum()),
                 spectrum_rel_max=float(rel.max()), spectrum_rel_gt_1e6=int((rel > 1e-6).sum()),
                 mwf_abs_max=float((d["maps"][:, 0] - e["maps"][:, 0]).abs().max()),
                 status_default=int((d["status"] != 0).sum()), status_echo=int((e["status"] != 0).sum()))
print(json.dumps(rep))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "echo_check_%s_%s.json" % (method, rm)), "w"), indent=1)
