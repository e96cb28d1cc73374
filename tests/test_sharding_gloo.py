"""Multi-rank path on CPU: world_size-2 gloo run of the slab partition + final gather.

The fit itself needs a GPU; here each rank fills its slab with a deterministic per-voxel function so that the
partition/scatter/gather logic is checked end to end: the gathered volumes must equal the single-rank result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_fit(sig):
    return sig.sum(axis=1) * 0.5 + 1.0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from multicomponent_t2_toolbox_b200 import pipeline
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    data = rng.uniform(0, 1, (5, 4, 3, 6))
    mask = (rng.uniform(0, 1, (5, 4, 3)) > 0.3).astype(np.int64)
    flat, sig = pipeline.masked_voxel_list(data, mask)
    lo, hi = pipeline.slab_bounds(len(flat), rank, world)
    vol = np.zeros(5 * 4 * 3)
    vol[flat[lo:hi]] = _fake_fit(sig[lo:hi])
    out = pipeline.gather_volumes({"MWF": vol.reshape(5, 4, 3), "T2s": np.arange(3.0)})
    full = np.zeros(5 * 4 * 3)
    full[flat] = _fake_fit(sig)
    ok = np.array_equal(out["MWF"].reshape(-1), full) and np.array_equal(out["T2s"], np.arange(3.0))
    # block-cyclic (over-decomposed) partition: same gather, same result
    mine = pipeline.cyclic_slab(len(flat), rank, world, chunk=4)
    vol2 = np.zeros(5 * 4 * 3)
    vol2[flat[mine]] = _fake_fit(sig[mine])
    out2 = pipeline.gather_volumes({"MWF": vol2.reshape(5, 4, 3)})
    ok = ok and np.array_equal(out2["MWF"].reshape(-1), full) and 0 < len(mine) < len(flat)
    q.put((rank, bool(ok), hi - lo))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) > 0
