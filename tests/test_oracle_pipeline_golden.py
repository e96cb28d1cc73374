"""The oracle's restatements of the stages either side of the voxel fit (NESMA denoiser, FA-stage smoothing, volume
driver, mean-spectrum diagnostics) held to what the UNMODIFIED reference orchestrator produced end to end
(oracle/make_golden_pipeline.py -> tests/golden/pipeline_nesma_x2.npz: denoise=NESMA, FA_smooth=yes, spline, X2-I)."""
import os

import numpy as np
import pytest
import scipy.ndimage as ndi

import met2_oracle as O
from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "pipeline_nesma_x2.npz")))


@pytest.fixture(scope="module")
def oracle_run(gold):
    data = gold["data"].copy()
    mask = gold["mask"].astype(np.int64)
    for c in range(data.shape[3]):
        data[:, :, :, c] *= mask
    data[data < 0.0] = 0.0
    den = O.nesma_filter(data, mask)
    smooth = np.zeros_like(den)
    for c in range(den.shape[3]):
        smooth[:, :, :, c] = ndi.gaussian_filter(den[:, :, :, c], 2.0, 0)
    out = O.recon_volume(den, mask, gold["TE"], 1000.0, "X2", "I", "spline", data_fa=smooth, num_cores=4)
    out["den"] = den
    return out


def test_volume_outputs_match_reference(gold, oracle_run):
    for k in ("MWF", "IEWF", "FWF", "T2_M", "T2_IE", "TWC", "FA", "fsol_4D", "Est_Signal", "reg_param"):
        a, b = oracle_run[k], gold[k]
        assert a.shape == b.shape, k
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9 * np.abs(b).max()), (k, np.abs(a - b).max())
    assert np.array_equal(oracle_run["fsol_4D"] > 0, gold["fsol_4D"] > 0)
    # NESMA really averaged neighbours on this phantom (otherwise the test would not pin it)
    data = gold["data"] * gold["mask"][..., None]
    changed = np.abs(oracle_run["den"] - data).max(axis=3)[gold["mask"] == 1]
    assert (changed > 0).mean() > 0.5


def test_mean_spectrum_curves_match_reference(gold, oracle_run):
    Dic = O.create_Dic_3D(60, oracle_run["T2s"], 1000.0 * np.ones(60), 32, 10.0, oracle_run["alpha_values"], 1000.0)
    dg = O.mean_spectrum_diagnostics(oracle_run["den"], gold["mask"], oracle_run["FA_index"], Dic,
                                     oracle_run["mean_T2_dist"])
    for k in ("mean_T2_dist", "dist_T2_mean1", "dist_T2_mean2"):
        assert np.allclose(dg[k], gold[k], rtol=1e-9, atol=1e-12), (k, np.abs(dg[k] - gold[k]).max())


def test_roi_estimator_matches_reference_run(gold):
    """Oracle restatement of the ROI-based estimator (motor_recon_met2_real_data_ROI.py:405-445) against the outputs of
    the UNMODIFIED reference's motor_recon_met2_ROIs run end to end (oracle/make_golden_roi.py: spline FA, no
    smoothing, no denoising, reg_matrix L2)."""
    r = dict(np.load(os.path.join(GOLDEN, "roi_x2_l2.npz")))
    mask = gold["mask"].astype(np.int64)
    data = gold["data"] * mask[..., None]
    data[data < 0.0] = 0.0
    g = O._grids("X2", "L2", "spline", 40.0, 32, 10.0, 1000.0)
    Dic = O.create_Dic_3D(60, g["T2s"], g["T1s"], 32, 10.0, g["alpha_values"], 1000.0)
    DicLR = O.create_Dic_3D(60, g["T2s"], g["T1s"], 32, 10.0, g["alpha_spline"], 1000.0)
    nx, ny, nz = mask.shape
    FA_index = np.zeros((nx, ny, nz))
    for z in range(nz):
        for y in range(ny):
            FA_index[:, y, z] = O.fitting_slice_FA_spline_method(DicLR, Dic, data[:, y, z, :], mask[:, y, z],
                                                                 g["alpha_spline"], nx, g["alpha_values"])[1]
    out = O.roi_estimates(data, mask, r["rois"], FA_index, Dic, g["L"], g["T2s"], g["ind_m"], g["ind_t"], g["ind_csf"])
    assert np.array_equal(out["roi_values"], r["labels"].astype(np.int64))
    assert np.allclose(out["fsol_ROIs"], r["spectra"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(out["fsol_ROIs"] > 0, r["spectra"] > 0)
    assert np.allclose(out["MWF"], r["MWF"], rtol=1e-9, atol=1e-12)
    tv = r["table_values"]
    for col, k in enumerate(("MWF", "IEWF", "FWF", "T2M", "T2IE", "TWC")):
        assert np.allclose(out[k], tv[:, col], rtol=1e-9, atol=1e-12), k
