"""Development timing (run under gpurun): full-size config 2 stage timings with CUDA events."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multicomponent_t2_toolbox_b200 import batched
from multicomponent_t2_toolbox_b200.phantom import make_phantom

shape = tuple(int(x) for x in os.environ.get("SHAPE", "96,96,60").split(","))
method = os.environ.get("METHOD", "X2"); rm = os.environ.get("RM", "I"); fam = os.environ.get("FA", "spline")
reps = int(os.environ.get("REPS", "3"))
t0 = time.time()
ph = make_phantom(shape, seed=2, fa_mode="b1", backend="gpu")
print("phantom s", time.time() - t0, flush=True)
sig = torch.as_tensor(ph["data"].reshape(-1, 32)).cuda()
V = sig.shape[0]
plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method=fam, t2_flags=int(os.environ.get("T2FLAGS", "0")))
torch.cuda.synchronize()
res = {"V": V, "method": method, "rm": rm, "fa": fam}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for r in range(reps):
    ev[0].record()
    fa = plan.fa_fit(sig)
    ev[1].record()
    t2 = plan.t2_fit(sig, fa["fa_index"])
    ev[2].record()
    torch.cuda.synchronize()
    res["rep%d" % r] = dict(fa_ms=ev[0].elapsed_time(ev[1]), t2_ms=ev[1].elapsed_time(ev[2]))
    print(res["rep%d" % r], flush=True)
tot = res["rep%d" % (reps - 1)]
res["vox_per_s"] = V / ((tot["fa_ms"] + tot["t2_ms"]) * 1e-3)
st = t2["status"].cpu().numpy()
res["status_counts"] = {str(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}
res["fa_hist_nonzero_bins"] = int((torch.bincount(fa["fa_index"].long()) > 0).sum())
res["mwf_mean"] = float(t2["maps"][:, 0].mean())
res["mwf_true_mean"] = float(ph["truth"]["mwf"].mean())
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gpu_time_%s_%s_%s.json" % (method, rm, fam)), "w"), indent=1)
