"""One process per GPU (torchrun): every rank fits its chunks of ONE volume, the slabs are all-gathered over NCCL
(pipeline.gather_slabs) and rank 0 compares the gathered arrays byte for byte with its own single-GPU fit of the whole
volume.  Also: pipeline.MultiGpuFit (one process, all GPUs) against the same fit.

    gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29544 tools/gpu_multi_check.py > gpurun_out/multi_check.log 2>&1'"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from multicomponent_t2_toolbox_b200 import batched, pipeline  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
host_group = dist.new_group(backend="gloo")     # host-side wait: the other ranks' GPUs stay idle while rank 0 drives them
shape = tuple(int(x) for x in os.environ.get("SHAPE", "48,48,12").split(","))
ph = make_phantom(shape, seed=2, fa_mode="b1", backend="gpu")
sig = ph["data"].reshape(-1, 32)
V = sig.shape[0]
plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method="X2", reg_matrix="I", FA_method="spline", device=dev)
ranges = pipeline.chunk_deal(V, world)[rank]
index = np.concatenate([np.arange(lo, hi) for lo, hi in ranges])
mine = pipeline.fit_voxels(plan, sig[index])
keys = ("fa_index", "fa_deg", "km", "fsol", "est_signal", "reg", "maps", "status")
full = pipeline.gather_slabs({k: mine[k] for k in keys}, index, V)
rep = dict(world=world, voxels=int(V))
if rank == 0:
    ref = pipeline.fit_voxels(plan, sig)
    rep["nccl_gather_identical"] = {k: bool(np.array_equal(full[k], ref[k])) for k in keys}
    n = torch.cuda.device_count()
    multi = pipeline.MultiGpuFit.create(n, 32, 10.0, 1000.0, reg_method="X2", reg_matrix="I", FA_method="spline")
    out = multi.fit(torch.as_tensor(sig).pin_memory())
    rep["multi_gpu_fit_devices"] = len(multi.plans)
    rep["multi_gpu_fit_identical"] = {k: bool(np.array_equal(out[k], ref[k])) for k in keys}
    ok = all(rep["nccl_gather_identical"].values()) and all(rep["multi_gpu_fit_identical"].values())
    rep["ok"] = bool(ok)
    print(json.dumps(rep), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "multi_check_%dgpu.json" % world), "w"), indent=1)
dist.barrier(group=host_group)
dist.destroy_process_group()
