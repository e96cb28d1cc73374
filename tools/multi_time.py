"""pipeline.MultiGpuFit on ONE config-2 volume, standalone (one process, all visible GPUs): wall time per call for 1..N
devices with a per-device phase trace (MET2_MULTI_TRACE=1).  -> gpurun_out/multi_time.json
env: SHAPE (96,96,60), TILE (replication factor per axis: 2 -> the 4.42 M-voxel volume of config 5), METHOD / RM / NPC,
REPS (timed calls, default 5), OUT (file name under gpurun_out/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multicomponent_t2_toolbox_b200 import pipeline  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

shape = tuple(int(x) for x in os.environ.get("SHAPE", "96,96,60").split(","))
ph = make_phantom(shape, seed=2, fa_mode="b1", backend="gpu")
tile = int(os.environ.get("TILE", "1"))
data = np.tile(ph["data"], (tile, tile, tile, 1)) if tile > 1 else ph["data"]
method, rm = os.environ.get("METHOD", "X2"), os.environ.get("RM", "I")
npc = int(os.environ["NPC"]) if "NPC" in os.environ else None
reps = int(os.environ.get("REPS", "5"))
vol = torch.as_tensor(data.reshape(-1, 32)).pin_memory()
V = vol.shape[0]
n_all = torch.cuda.device_count()
rec = {"voxels": int(V), "devices_visible": n_all, "method": method, "reg_matrix": rm, "runs": []}
bufs = None
for n in sorted({1, 2, 4, n_all} & set(range(1, n_all + 1))):
    multi = pipeline.MultiGpuFit.create(n, 32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline", npc=npc)
    if bufs is None:
        bufs = pipeline.host_buffers(multi.plans[0], V)
    for _ in range(2 if tile > 1 else 3):
        multi.fit(vol, out=bufs)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = multi.fit(vol, out=bufs)
        ts.append(1e3 * (time.perf_counter() - t0))
    rec["runs"].append({"gpus": n, "ms_median": float(np.median(ts)), "ms_all": ts, "trace": getattr(multi, "last_trace", None),
                        "mwf_mean": float(r["maps"][:, 0].mean())})
    print(json.dumps(rec["runs"][-1]), flush=True)
t1 = rec["runs"][0]["ms_median"]
for r in rec["runs"]:
    r["strong_efficiency"] = t1 / (r["gpus"] * r["ms_median"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", os.environ.get("OUT", "multi_time.json")), "w"), indent=1)
print(json.dumps({r["gpus"]: (round(r["ms_median"], 1), round(r["strong_efficiency"], 3)) for r in rec["runs"]}))
