"""One T2-stage launch of a given method on a slab of the config-2 phantom, for ncu captures.
env: METHOD, RM, FA, SHAPE (default 96,96,6 = 55 296 voxels), T2FLAGS (MET2_T2_FLAG_* bits), NTE, TAU, NPC."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multicomponent_t2_toolbox_b200 import batched
from multicomponent_t2_toolbox_b200.phantom import make_phantom
shape = tuple(int(x) for x in os.environ.get("SHAPE", "96,96,6").split(","))
method = os.environ.get("METHOD", "X2"); rm = os.environ.get("RM", "I"); fam = os.environ.get("FA", "spline")
nte = int(os.environ.get("NTE", "32")); tau = float(os.environ.get("TAU", "10.0"))
npc = int(os.environ["NPC"]) if "NPC" in os.environ else None
ph = make_phantom(shape, n_echoes=nte, tau=tau, seed=2, fa_mode="b1", backend="gpu")
sig = torch.as_tensor(ph["data"].reshape(-1, nte)).cuda()
plan = batched.Met2Plan(nte, tau, 1000.0, reg_method=method, reg_matrix=rm, FA_method=fam, npc=npc,
                        t2_flags=int(os.environ.get("T2FLAGS", "0")))
fa = plan.fa_fit(sig)
t2 = plan.t2_fit(sig, fa["fa_index"])
torch.cuda.synchronize()
print("ok", method, rm, sig.shape[0], float(t2["maps"][:, 0].mean()))
