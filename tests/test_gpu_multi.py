"""One volume over several GPUs (pipeline.MultiGpuFit, the `num_cores` -> GPUs path of motor_recon_met2): per-voxel
outputs must be byte-identical to the single-GPU fit — voxels are independent and the kernels do not depend on the
batch a voxel arrives in.  On a one-GPU box the same code runs with one device (chunked copies, pinned buffers); the
N > 1 case runs wherever torch.cuda.device_count() > 1 (gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu)."""
import numpy as np
import pytest
import torch

from multicomponent_t2_toolbox_b200 import batched, pipeline
from multicomponent_t2_toolbox_b200.phantom import make_phantom

pytestmark = pytest.mark.gpu

KEYS = ("fa_index", "fa_deg", "km", "fa_status", "fsol", "est_signal", "reg", "maps", "status")


@pytest.mark.parametrize("method,rm", [("X2", "I"), ("L_curve", "L2")])
def test_multi_gpu_fit_equals_single_fit(method, rm):
    ph = make_phantom((24, 20, 6), seed=12, fa_mode="b1")
    sig = ph["data"].reshape(-1, 32)
    V = sig.shape[0]
    plan = batched.Met2Plan(32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline", device="cuda:0")
    ref = pipeline.fit_voxels(plan, sig)
    n = torch.cuda.device_count()
    multi = pipeline.MultiGpuFit.create(n, 32, 10.0, 1000.0, reg_method=method, reg_matrix=rm, FA_method="spline",
                                        chunks_per_device=3)
    assert len(multi.plans) == n
    pin = torch.as_tensor(sig).pin_memory()
    for rep in range(2):                       # second call reuses the per-device buffers
        out = multi.fit(pin)
        for k in KEYS:
            assert np.array_equal(out[k], ref[k]), (k, rep)
        assert np.allclose(out["fsol_sum"], ref["fsol_sum"], rtol=1e-12, atol=0)
    # smoothed FA-stage input as a second host array
    sig_fa = np.ascontiguousarray(sig * 0.98 + 0.02 * sig.mean(axis=0))
    ref2 = pipeline.fit_voxels(plan, sig, sig_fa)
    out2 = multi.fit(pin, torch.as_tensor(sig_fa).pin_memory())
    for k in KEYS:
        assert np.array_equal(out2[k], ref2[k]), k
    assert V == sum(hi - lo for part in pipeline.chunk_deal(V, n, 3) for lo, hi in part)


def test_recon_arrays_all_gpus_equals_one_gpu():
    ph = make_phantom((12, 10, 8), seed=5, fa_mode="b1", mask_mode="ellipsoid")
    TE = 10.0 * np.arange(1, 33)
    args = (ph["data"], ph["mask"], TE, 1000.0, "X2", "I", "spline")
    one = pipeline.recon_arrays(*args, n_gpus=1, diagnostics=True)
    allg = pipeline.recon_arrays(*args, n_gpus=None, diagnostics=True)
    for k in ("MWF", "IEWF", "FWF", "T2_M", "T2_IE", "TWC", "FA", "FA_index", "reg_param", "fsol_4D", "Est_Signal", "status"):
        assert np.array_equal(one[k], allg[k]), k
    for k in ("mean_T2_dist", "dist_T2_mean1", "dist_T2_mean2"):
        assert np.allclose(one["diagnostics"][k], allg["diagnostics"][k], rtol=1e-9, atol=1e-12), k
