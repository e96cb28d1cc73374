"""Development timing (run under gpurun): the kernels either side of the voxel fit on the config-2 volume —
NESMA denoiser, FA-stage Gaussian smoothing, mean-spectrum diagnostics, ROI estimator — with the oracle's NumPy
restatement of NESMA timed on a sub-volume beside it."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from multicomponent_t2_toolbox_b200 import batched, pipeline
from multicomponent_t2_toolbox_b200.phantom import make_phantom

def ev_time(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return r, best

shape = tuple(int(x) for x in os.environ.get("SHAPE", "96,96,60").split(","))
ph = make_phantom(shape, seed=2, fa_mode="b1", backend="gpu", snr_range=(300.0, 600.0))
data = torch.as_tensor(ph["data"]).cuda()
mask = torch.ones(shape, dtype=torch.int32, device="cuda")
V = int(np.prod(shape))
res = {"shape": shape, "V": V}
den, ms = ev_time(lambda: batched.nesma_filter(data, mask))
res["nesma_ms"] = ms
res["nesma_voxels_per_s"] = V / (ms * 1e-3)
# algorithmic traffic: every voxel reads its (up to) 12^3 window x nTE doubles once for the distance and the accepted
# ones again for the mean; unique HBM bytes are just volume in + echo-major copy + volume out
res["nesma_window_bytes_algorithmic"] = float(V) * 1728 * 32 * 8
res["nesma_window_TBps"] = res["nesma_window_bytes_algorithmic"] / (ms * 1e-3) / 1e12
res["nesma_changed_fraction"] = float(((den - data).abs().amax(dim=3) > 0).float().mean())
sm, ms = ev_time(lambda: batched.gaussian_smooth(data, 2.0))
res["gaussian_smooth_ms"] = ms
plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="X2", reg_matrix="I", FA_method="spline")
sig = data.reshape(-1, 32)
fa = plan.fa_fit(sig)
inm = torch.ones(V, dtype=torch.int32, device="cuda")
_, ms = ev_time(lambda: pipeline.mean_spectrum_diagnostics(plan, sig, fa["fa_index"], inm, fa["fsol_sum"]))
res["mean_spectrum_diagnostics_ms"] = ms
_, ms = ev_time(lambda: batched.segment_means(sig, fa["fa_index"], torch.zeros(V, dtype=torch.int32, device="cuda"), 1, plan.dict_hr))
res["segment_means_1seg_ms"] = ms
lab = (torch.arange(V, device="cuda") // 4096 % 97 + 1)
vals = np.arange(1, 98)
_, ms = ev_time(lambda: pipeline.roi_estimates(plan, sig, fa["fa_index"], lab, vals))
res["roi_estimates_97rois_ms"] = ms
# CPU: the oracle's loop-for-loop NumPy restatement of the reference NESMA on a 24x24x12 corner (single core)
import met2_oracle as O
sub = ph["data"][:24, :24, :12].copy()
t0 = time.time(); O.nesma_filter(sub, np.ones(sub.shape[:3], dtype=np.int64)); dt = time.time() - t0
res["nesma_numpy_voxels_per_s_1core"] = sub.shape[0] * sub.shape[1] * sub.shape[2] / dt
res["nesma_numpy_sample"] = "24x24x12 corner, %.1f s" % dt
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "aux_time.json"), "w"), indent=1)
