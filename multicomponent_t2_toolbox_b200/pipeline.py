"""Volume-level driver: what motor_recon_met2 does between "arrays are in host memory" and "arrays are handed to the
NIfTI writer" (motor/motor_recon_met2_real_data.py:167-182, 279, 349-373, 428-472), as gather -> two batched GPU calls
-> scatter.  Also holds the voxel-slab partition used for multi-GPU runs: voxels are independent, so
each rank fits a contiguous slab of the masked-voxel list and no collective runs during the fit.
"""
import numpy as np
import torch

from . import batched

MAP_NAMES = ("MWF", "IEWF", "FWF", "T2_M", "T2_IE", "TWC")


def slab_bounds(n_voxels, rank, world_size):
    """Contiguous slab [lo, hi) of rank `rank` out of `world_size`: ceil(V / W) voxels per rank, last ranks may be short
    or empty.  Every voxel belongs to exactly one rank; concatenating the slabs in rank order restores the list."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    per = -(-int(n_voxels) // int(world_size)) if n_voxels > 0 else 0
    lo = min(n_voxels, rank * per)
    hi = min(n_voxels, lo + per)
    return lo, hi


def masked_voxel_list(data, mask):
    """Apply the mask and the negative clamp like motor...:178-180,279 and return (flat voxel indices, signals[V,nTE])
    of the voxels with mask > 0, in C order of (x, y, z)."""
    nx, ny, nz, nt = data.shape
    m = np.asarray(mask).reshape(-1)
    flat = np.nonzero(m > 0)[0]
    sig = np.ascontiguousarray(data.reshape(-1, nt)[flat], dtype=np.float64)
    sig = sig * m[flat][:, None]
    sig[sig < 0.0] = 0.0
    return flat, sig


def fit_voxels(plan, sig, sig_fa=None, pinned_out=None):
    """Steps 2-4 on a list of voxels given as host array sig[V, nTE]; returns host arrays (dict)."""
    dev = plan.dev
    V = sig.shape[0]
    with torch.cuda.device(dev):
        d_sig = torch.as_tensor(sig).to(dev, non_blocking=True)
        if sig_fa is None:
            d_sig_fa = d_sig
        elif isinstance(sig_fa, torch.Tensor):
            d_sig_fa = sig_fa.to(dev)
        else:
            d_sig_fa = torch.as_tensor(sig_fa).to(dev, non_blocking=True)
        fa = plan.fa_fit(d_sig_fa)
        t2 = plan.t2_fit(d_sig, fa["fa_index"])
        out = dict(fa_index=fa["fa_index"], fa_deg=fa["fa_deg"], km=fa["km"], fsol_sum=fa["fsol_sum"],
                   fa_status=fa["status"], fsol=t2["fsol"], est_signal=t2["est_signal"], reg=t2["reg"], maps=t2["maps"],
                   status=t2["status"])
        host = {}
        for k, t in out.items():
            if pinned_out is not None and k in pinned_out:
                pinned_out[k][:t.shape[0]].copy_(t, non_blocking=True)
                host[k] = pinned_out[k][:t.shape[0]]
            else:
                host[k] = t.cpu()
        torch.cuda.synchronize(dev)
    return {k: v.numpy() for k, v in host.items()}


def recon_arrays(data, mask, TE_array, TR, reg_method, reg_matrix, FA_method, myelin_T2=40.0, data_fa=None, plan=None,
                 npc=None, n_alphas=None, device=None, rank=0, world_size=1):
    """Steps 2-4 of motor_recon_met2 on in-memory arrays.  Returns the ten output volumes (plus FA_index) as numpy.

    With world_size > 1 only this rank's slab of the masked voxels is fitted and the other voxels are left zero; the
    caller combines the ranks (sum of volumes, or `gather_volumes`).
    """
    data = np.asarray(data, dtype=np.float64)
    nx, ny, nz, nt = data.shape
    TE_array = np.asarray(TE_array, dtype=np.float64)
    if plan is None:
        plan = batched.Met2Plan(TE_array.shape[0], TE_array[1] - TE_array[0], TR, reg_method=reg_method,
                                reg_matrix=reg_matrix, FA_method=FA_method, myelin_T2=myelin_T2, npc=npc,
                                n_alphas=n_alphas, device=device)
    flat, sig = masked_voxel_list(data, mask)
    lo, hi = slab_bounds(len(flat), rank, world_size)
    sig_fa = None
    if data_fa is not None:
        if isinstance(data_fa, torch.Tensor):
            # smoothed volume already on the GPU (batched.gaussian_smooth): gather the slab's voxels there
            idx = torch.as_tensor(flat[lo:hi]).to(data_fa.device)
            m = torch.as_tensor(np.asarray(mask).reshape(-1)[flat[lo:hi]].astype(np.float64)).to(data_fa.device)
            sig_fa = (data_fa.reshape(-1, nt).index_select(0, idx) * m[:, None]).clamp_min(0.0)
        else:
            _, sig_fa_all = masked_voxel_list(np.asarray(data_fa, dtype=np.float64), mask)
            sig_fa = sig_fa_all[lo:hi]
    res = fit_voxels(plan, sig[lo:hi], sig_fa)
    sel = flat[lo:hi]
    nvox = nx * ny * nz
    vol = {}
    for i, name in enumerate(MAP_NAMES):
        a = np.zeros(nvox)
        a[sel] = res["maps"][:, i]
        vol[name] = a.reshape(nx, ny, nz)
    for name, key in (("FA", "fa_deg"), ("FA_index", "fa_index"), ("reg_param", "reg")):
        a = np.zeros(nvox)
        a[sel] = res[key]
        vol[name] = a.reshape(nx, ny, nz)
    f4 = np.zeros((nvox, plan.npc))
    f4[sel] = res["fsol"]
    vol["fsol_4D"] = f4.reshape(nx, ny, nz, plan.npc)
    s4 = np.zeros((nvox, nt))
    s4[sel] = res["est_signal"]
    vol["Est_Signal"] = s4.reshape(nx, ny, nz, nt)
    vol["mean_T2_dist"] = res["fsol_sum"]
    vol["status"] = np.zeros(nvox, dtype=np.int32)
    vol["status"][sel] = res["status"]
    vol["status"] = vol["status"].reshape(nx, ny, nz)
    vol["T2s"] = plan.T2s
    return vol


def gather_volumes(vol, group=None):
    """Final gather of a multi-rank run: every voxel was written by exactly one rank and is zero elsewhere, so the
    gather is a sum over ranks (torch.distributed all_reduce; NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return vol
    use_cuda = dist.get_backend(group) == "nccl"
    out = {}
    for k, a in vol.items():
        if k == "T2s":
            out[k] = a
            continue
        t = torch.as_tensor(np.ascontiguousarray(a))
        if use_cuda:
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out[k] = t.cpu().numpy()
    return out
