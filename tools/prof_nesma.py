"""One NESMA launch on the config-2 volume (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multicomponent_t2_toolbox_b200 import batched
from multicomponent_t2_toolbox_b200.phantom import make_phantom
ph = make_phantom((96, 96, 60), seed=2, fa_mode="b1", backend="gpu", snr_range=(300.0, 600.0))
data = torch.as_tensor(ph["data"]).cuda()
mask = torch.ones((96, 96, 60), dtype=torch.int32, device="cuda")
out = batched.nesma_filter(data, mask)
torch.cuda.synchronize()
print("ok", float(out.mean()))
