"""Multi-rank path on CPU: world_size-2 gloo run of the slab partition + final gather.

The fit itself needs a GPU; here each rank fills its slab with a deterministic per-voxel function so that the
partition / all-gather-of-slabs logic (pipeline.gather_slabs) is checked end to end: the gathered arrays must equal
the single-rank result."""
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
    # this rank's per-voxel outputs for ITS slab only (a vector and a [V, 3] array, like reg / maps of the real fit)
    mine = {"reg": _fake_fit(sig[lo:hi]), "maps": np.stack([sig[lo:hi, 0], sig[lo:hi, 1], sig[lo:hi, 2]], 1),
            "idx": np.arange(lo, hi, dtype=np.int32)}
    out = pipeline.gather_slabs(mine, np.arange(lo, hi), len(flat))
    ok = np.array_equal(out["reg"], _fake_fit(sig)) and np.array_equal(out["maps"], sig[:, :3]) and \
        np.array_equal(out["idx"], np.arange(len(flat), dtype=np.int32))
    # over-decomposed partition (chunk_deal, the one MultiGpuFit uses): same gather, same result
    ranges = pipeline.chunk_deal(len(flat), world, chunks_per_part=3, align=2)[rank]
    sel = np.concatenate([np.arange(a, b) for a, b in ranges]) if ranges else np.zeros(0, dtype=np.int64)
    out2 = pipeline.gather_slabs({"reg": _fake_fit(sig[sel])}, sel, len(flat))
    ok = ok and np.array_equal(out2["reg"], _fake_fit(sig)) and 0 < len(sel) < len(flat)
    # block-cyclic index sets
    cyc = pipeline.cyclic_slab(len(flat), rank, world, chunk=4)
    out3 = pipeline.gather_slabs({"reg": _fake_fit(sig[cyc])}, cyc, len(flat))
    ok = ok and np.array_equal(out3["reg"], _fake_fit(sig))
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
