"""A/B of the two ranks of the reduced echo space (16: the plan's choice for the 32-echo protocol, 24: exact to the
dictionary's rounding) and the Gram-domain kernels over the whole config-2 volume (552 960 voxels): T2-stage time,
voxels whose active set differs, largest spectrum / MWF difference.  X2-I and T2SPARC-InvT2 (96 bins).

    gpurun --timeout 600 -- 'timeout 400 python tools/gpu_ab_echo_rank.py > gpurun_out/ab_echo_rank.log 2>&1'
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multicomponent_t2_toolbox_b200 import batched  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

shape = tuple(int(x) for x in os.environ.get("SHAPE", "96,96,60").split(","))
ph = make_phantom(shape, seed=2, fa_mode="b1", backend="gpu")
sig = torch.as_tensor(ph["data"].reshape(-1, 32)).cuda()
rep = {}
for method, rm, npc in (("X2", "I", None), ("X2", "InvT2", None), ("T2SPARC", "InvT2", 96)):
    res = {}
    for name, kw in (("rank16", dict(echo_ranks=(16,))), ("rank24", dict(echo_ranks=(24,))), ("gram", dict(echo_space=False))):
        plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline", npc=npc, **kw)
        fa = plan.fa_fit(sig)
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = plan.t2_fit(sig, fa["fa_index"])
            e1.record()
            torch.cuda.synchronize()
        res[name] = dict(ms=e0.elapsed_time(e1), out={k: v.clone() for k, v in out.items()},
                         tail={R: t[2] for R, t in plan.dict_hr.__dict__.get("_echo_tables", {}).items()})
    r = dict(voxels=int(sig.shape[0]))
    base = res["rank24"]["out"]
    scale = base["fsol"].abs().max(dim=1).values.clamp_min(1e-300)
    for name in ("rank16", "rank24", "gram"):
        o = res[name]["out"]
        rel = (o["fsol"] - base["fsol"]).abs().max(dim=1).values / scale
        bad = ((o["fsol"] > 0) != (base["fsol"] > 0)).any(dim=1)
        r[name] = dict(t2_ms=res[name]["ms"], tail=res[name]["tail"], status_nonzero=int((o["status"] != 0).sum()),
                       active_set_differs_from_rank24=int(bad.sum()),
                       spectrum_rel_max_agreeing_vs_rank24=float(rel[~bad].max()),
                       spectrum_rel_gt_1e6_vs_rank24=int((rel > 1e-6).sum()),
                       mwf_abs_max_agreeing_vs_rank24=float((o["maps"][:, 0] - base["maps"][:, 0]).abs()[~bad].max()))
    rep["%s_%s" % (method, rm)] = r
    print(json.dumps({"%s_%s" % (method, rm): r}), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "ab_echo_rank.json"), "w"), indent=1)
