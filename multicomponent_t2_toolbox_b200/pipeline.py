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


def cyclic_slab(n_voxels, rank, world_size, chunk=2048):
    """Block-cyclic alternative to `slab_bounds` (over-decomposition for load balance — the cost of a
    voxel depends on its flip angle and spectrum, and both vary smoothly across the volume): chunks of `chunk`
    consecutive voxels are dealt to the ranks round-robin.  Returns this rank's voxel indices (sorted, int64); the
    ranks' index sets are disjoint and cover [0, n_voxels)."""
    if world_size < 1 or not (0 <= rank < world_size) or chunk < 1:
        raise ValueError("bad rank/world_size/chunk")
    idx = np.arange(int(n_voxels), dtype=np.int64)
    return idx[(idx // int(chunk)) % int(world_size) == rank]


def masked_voxel_list(data, mask, premasked=False):
    """Return (flat voxel indices, signals[V,nTE]) of the voxels with mask > 0, in C order of (x, y, z).  Unless
    `premasked` (the motor entry points have already done it, motor...:178-180,279 — applying a label mask twice would
    scale the data by mask squared), the mask multiplication and the negative clamp are applied here."""
    nx, ny, nz, nt = data.shape
    m = np.asarray(mask).reshape(-1)
    flat = np.nonzero(m > 0)[0]
    sig = np.ascontiguousarray(data.reshape(-1, nt)[flat], dtype=np.float64)
    if not premasked:
        sig = sig * m[flat][:, None]
        sig[sig < 0.0] = 0.0
    return flat, sig


def fit_voxels(plan, sig, sig_fa=None, pinned_out=None, in_mask=None, roi=None):
    """Steps 2-4 on a list of voxels given as host array sig[V, nTE]; returns host arrays (dict).
    in_mask[V] (optional): also return the mean-spectrum diagnostics of motor...:375-403 over the voxels with
    in_mask == 1 (key "diagnostics").  roi = (labels[V], values) (optional): also the ROI-based estimates of
    motor_recon_met2_real_data_ROI.py:405-445 (key "roi")."""
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
        extra = {}
        if in_mask is not None and V > 0:
            extra["diagnostics"] = mean_spectrum_diagnostics(plan, d_sig, fa["fa_index"], in_mask, fa["fsol_sum"])
        if roi is not None and V > 0:
            extra["roi"] = roi_estimates(plan, d_sig, fa["fa_index"], roi[0], roi[1])
        host = {}
        for k, t in out.items():
            if pinned_out is not None and k in pinned_out:
                pinned_out[k][:t.shape[0]].copy_(t, non_blocking=True)
                host[k] = pinned_out[k][:t.shape[0]]
            else:
                host[k] = t.cpu()
        torch.cuda.synchronize(dev)
    res = {k: v.numpy() for k, v in host.items()}
    res.update(extra)
    return res


def recon_arrays(data, mask, TE_array, TR, reg_method, reg_matrix, FA_method, myelin_T2=40.0, data_fa=None, plan=None,
                 npc=None, n_alphas=None, device=None, rank=0, world_size=1, diagnostics=False, rois=None,
                 premasked=False, fa_only=False):
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
    flat, sig = masked_voxel_list(data, mask, premasked)
    lo, hi = slab_bounds(len(flat), rank, world_size)
    sig_fa = None
    if data_fa is not None:
        if isinstance(data_fa, torch.Tensor):
            # smoothed volume already on the GPU (batched.gaussian_smooth): gather the slab's voxels there
            idx = torch.as_tensor(flat[lo:hi]).to(data_fa.device)
            sig_fa = data_fa.reshape(-1, nt).index_select(0, idx)
            if not premasked:
                m = torch.as_tensor(np.asarray(mask).reshape(-1)[flat[lo:hi]].astype(np.float64)).to(data_fa.device)
                sig_fa = (sig_fa * m[:, None]).clamp_min(0.0)
        else:
            _, sig_fa_all = masked_voxel_list(np.asarray(data_fa, dtype=np.float64), mask, premasked)
            sig_fa = sig_fa_all[lo:hi]
    if (diagnostics or rois is not None) and world_size != 1:
        raise ValueError("mean-spectrum diagnostics / ROI estimates are whole-volume reductions: run them unsharded")
    roi = None
    if rois is not None:
        # ROIs * mask, labels = np.unique without 0 (motor_recon_met2_real_data_ROI.py:166-183)
        rl = (np.asarray(rois).astype(np.int64) * np.asarray(mask).astype(np.int64)).reshape(-1)
        vals = np.unique(np.asarray(rois).astype(np.int64))
        roi = (rl[flat], vals[vals != 0])
    res = fit_voxels(plan, sig[lo:hi], sig_fa, in_mask=(np.asarray(mask).reshape(-1)[flat] if diagnostics else None),
                     roi=roi)
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
    for k in ("diagnostics", "roi"):
        if k in res:
            vol[k] = res[k]
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
        if k in ("T2s", "diagnostics", "roi"):
            out[k] = a
            continue
        t = torch.as_tensor(np.ascontiguousarray(a))
        if use_cuda:
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out[k] = t.cpu().numpy()
    return out


def _segment_x2(plan, mean_signal, mean_kernel, Laplac_plan, factor):
    """X2 fit (algorithms.py:211, no normalisation by M[0]) of every segment's mean signal against its mean kernel."""
    d = batched.Dictionary.from_device(mean_kernel)
    idx = torch.arange(mean_signal.shape[0], dtype=torch.int32, device=mean_signal.device)
    kest = Laplac_plan.t2_fit(mean_signal, idx, reg_method="X2", flags=2, dictionary=d, factor=float(factor))
    lam = Laplac_plan.t2_fit(mean_signal, idx, reg_method="X2", flags=3, dictionary=d, factor=float(factor))
    return d, idx, kest, lam


def mean_spectrum_diagnostics(plan, sig, fa_index, in_mask, fsol_sum, factor=1.01):
    """The three curves of the reference's 'Mean_spectrum_from_all_voxels' figure (motor...:375-403):
    mean_T2_dist (normalised sum of the FA-stage NNLS spectra), dist_T2_mean1 (NNLS of the mask-mean signal against the
    mask-mean kernel) and dist_T2_mean2 (X2 with factor 1.01 and the identity matrix on the same pair).
    sig[V, nTE] / fa_index[V] on the GPU; in_mask[V]: 1 where the reference's `mask == 1` test holds."""
    dev = plan.dev
    labels = torch.as_tensor(np.asarray(in_mask)).to(dev) if not isinstance(in_mask, torch.Tensor) else in_mask.to(dev)
    labels = torch.where(labels == 1, 0, -1).to(torch.int32)
    msig, mker, counts = batched.segment_means(sig, fa_index, labels, 1, plan.dict_hr)
    plan_I = plan if np.array_equal(plan.Laplac, np.eye(plan.npc)) else batched.Met2Plan(
        plan.nTE, plan.tau, plan.TR, reg_method="X2", reg_matrix="I", FA_method="brute-force", npc=plan.npc,
        Dic_3D=np.zeros((plan.nTE, plan.npc, 1)), T2s=plan.T2s, device=dev)
    d, idx, kest, lam = _segment_x2(plan, msig, mker, plan_I, factor)
    f1 = plan_I.t2_fit(msig, idx, reg_method="NNLS", flags=2, dictionary=d)["fsol"][0].cpu().numpy()
    f2 = kest["fsol"][0].cpu().numpy()
    fs = np.asarray(fsol_sum.cpu().numpy() if isinstance(fsol_sum, torch.Tensor) else fsol_sum, dtype=np.float64)
    with np.errstate(all="ignore"):
        return dict(T2s=plan.T2s, mean_T2_dist=fs / np.sum(fs), dist_T2_mean1=f1 / np.sum(f1),
                    dist_T2_mean2=f2 / np.sum(f2), total_signal=msig[0].cpu().numpy(),
                    total_Kernel=mker[0].cpu().numpy(), nv=int(counts[0]), reg_opt2=float(lam["reg"][0]),
                    k_est=float(kest["reg"][0]))


def roi_estimates(plan, sig, fa_index, roi_labels, roi_values, factor=1.01):
    """ROI-based estimator of motor/motor_recon_met2_real_data_ROI.py:405-445: for every ROI value, X2 (factor 1.01,
    the plan's regularisation matrix) of the ROI-mean signal against the ROI-mean kernel, then the Step-4 metrics of the
    normalised spectrum.  sig[V, nTE], fa_index[V], roi_labels[V] (integer label per voxel), roi_values: the labels to
    report (np.unique(ROIs) without 0 in the reference).  Returns host arrays: fsol_ROIs[nROI, nT2] (normalised),
    MWF/IEWF/FWF/T2M/T2IE/TWC[nROI], reg_opt[nROI], k_est[nROI], counts[nROI]."""
    dev = plan.dev
    roi_values = np.asarray(roi_values).astype(np.int64)
    lab = roi_labels if isinstance(roi_labels, torch.Tensor) else torch.as_tensor(np.asarray(roi_labels).astype(np.int64))
    lab = lab.to(dev).to(torch.int64)
    # label value -> segment id (position in roi_values); everything else -> -1
    seg = torch.full_like(lab, -1)
    for i, val in enumerate(roi_values.tolist()):
        seg = torch.where(lab == val, i, seg)
    n = len(roi_values)
    msig, mker, counts = batched.segment_means(sig, fa_index, seg.to(torch.int32), n, plan.dict_hr)
    d, idx, kest, lam = _segment_x2(plan, msig, mker, plan, factor)
    maps = kest["maps"].cpu().numpy()
    f = kest["fsol"].cpu().numpy()
    vt = maps[:, 5]
    return dict(roi_values=roi_values, fsol_ROIs=f / vt[:, None], MWF_ROIs=maps[:, 0], IEWF_ROIs=maps[:, 1],
                FWF_ROIs=maps[:, 2], T2M_ROIs=maps[:, 3], T2IE_ROIs=maps[:, 4], TWC_ROIs=vt,
                reg_opt=lam["reg"].cpu().numpy(), k_est=kest["reg"].cpu().numpy(), counts=counts.cpu().numpy(),
                mean_signal=msig.cpu().numpy(), mean_kernel=mker.cpu().numpy())
