"""Drop-in surface on the GPU: the reference's row workers / per-voxel functions / motor_recon_met2 under their own
names, checked against the golden vectors of the unmodified reference and the oracle."""
import os

import numpy as np
import pytest

import met2_oracle as O
from multicomponent_t2_toolbox_b200 import nifti_io
from multicomponent_t2_toolbox_b200.epg.epg import create_Dic_3D
from multicomponent_t2_toolbox_b200.flip_angle_algorithms.fa_estimation import (
    compute_optimal_FA, fitting_slice_FA_brute_force, fitting_slice_FA_spline_method)
from multicomponent_t2_toolbox_b200.intravoxel_algorithms.algorithms import (nnls, nnls_gcv, nnls_lcurve_wrapper,
                                                                             nnls_tik, nnls_x2)
from multicomponent_t2_toolbox_b200.intravoxel_algorithms.bayesian_interpolation import BayesReg_nnls
from multicomponent_t2_toolbox_b200.motor.motor_recon_met2_real_data import (create_Laplacian_matrix, fitting_slice_T2,
                                                                             motor_recon_met2)
from multicomponent_t2_toolbox_b200.phantom import make_phantom

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dics(golden_voxels):
    T2s = golden_voxels["T2s"]
    T1s = 1000.0 * np.ones(60)
    a273, a15, a91 = np.linspace(90, 180, 273), np.linspace(90, 180, 15), np.linspace(90, 180, 91)
    mk = lambda a: create_Dic_3D(60, T2s, T1s, 32, 10.0, a, 1000.0)
    return dict(a273=a273, a15=a15, a91=a91, d273=mk(a273), d15=mk(a15), d91=mk(a91))


def test_create_Dic_3D_layout(golden_dictionary):
    g = golden_dictionary
    D = create_Dic_3D(60, g["T2s"], g["T1s"], int(g["nte"]), float(g["tau"]), g["alphas"], float(g["TR"]))
    assert D.shape == g["dic"].shape and np.abs(D - g["dic"]).max() <= 1e-13


def test_row_workers_against_reference_golden(golden_voxels, dics):
    g = golden_voxels
    nx = len(g["mask"])
    FA, idx, KM, fs = fitting_slice_FA_brute_force(g["mask"], g["sig"], nx, dics["d91"], dics["a91"])
    assert np.array_equal(idx, g["fa_brute_idx"]) and np.array_equal(FA, g["fa_brute_deg"])
    assert np.allclose(KM, g["fa_brute_km"], rtol=1e-7) and np.allclose(fs, g["fa_brute_fsum"], rtol=1e-6, atol=1e-6)
    FA, idx, KM, fs = fitting_slice_FA_spline_method(dics["d15"], dics["d273"], g["sig"], g["mask"], dics["a15"], nx,
                                                     dics["a273"])
    assert np.array_equal(idx, g["fa_spline_idx"]) and np.allclose(KM, g["fa_spline_km"], rtol=1e-7)
    for method, rm in [("X2", "L1"), ("L_curve", "I"), ("NNLS", "I")]:
        L = O._grids(method, rm, "spline", 40.0, 32, 10.0, 1000.0, npc=60)["L"]
        f, s, reg = fitting_slice_T2(g["mask"], g["sig"], g["fa_spline_idx"], nx, dics["d273"], g["lambda_reg"], 60, 32,
                                     method, L, None)
        gf = g["t2_%s_%s_f" % (method, rm)]
        assert np.array_equal(f > 0, gf > 0)
        assert np.abs(f - gf).max() < 1e-6 * np.abs(gf).max()
        assert np.allclose(reg, g["t2_%s_%s_reg" % (method, rm)], rtol=1e-6)
        assert np.abs(s - g["t2_%s_%s_s" % (method, rm)]).max() < 1e-6 * np.abs(gf).max()


def test_per_voxel_api(golden_voxels, dics):
    g = golden_voxels
    D = np.ascontiguousarray(dics["d273"][:, :, int(g["nnls_D_index"])])
    M = g["sig"][0] / g["sig"][0, 0]
    x, rn = nnls(D, M)
    assert np.array_equal(x > 0, g["nnls_x"] > 0) and np.allclose(x, g["nnls_x"], rtol=1e-6, atol=1e-9)
    assert abs(rn - float(g["nnls_rnorm"])) < 1e-8
    L = np.eye(60)
    for fun_g, fun_o in [(lambda: nnls_tik(D, M, L, 0.01), lambda: O.nnls_tik(D, M, L, 0.01))]:
        assert np.allclose(fun_g(), fun_o(), rtol=1e-6, atol=1e-9)
    f, lam, k = nnls_x2(D, M, L, 1.02)
    fo, lamo, ko = O.nnls_x2(D, M, L, 1.02)
    assert np.allclose(f, fo, rtol=1e-6, atol=1e-9) and abs(lam - lamo) < 1e-6 * lamo and abs(k - ko) < 1e-6
    assert nnls_lcurve_wrapper(D, M, L, g["lambda_reg"]) == O.nnls_lcurve_wrapper(D, M, L, g["lambda_reg"])
    fb, lb = BayesReg_nnls(D, M, L)
    fbo, lbo = O.BayesReg_nnls(D, M, L)
    assert np.allclose(fb, fbo, rtol=1e-6, atol=1e-9) and abs(lb - lbo) < 1e-6 * lbo
    fg, lg = nnls_gcv(D, M, L)                      # GCV: statistical parity only (see test_gpu_parity)
    assert fg.shape == (60,) and 1e-8 <= lg <= 10.0 and (fg >= 0).all()
    idx, alpha, km, sse, fsol = compute_optimal_FA(g["sig"][0], dics["d91"], dics["a91"])
    io, ao, kmo, sseo, fo = O.compute_optimal_FA(g["sig"][0], dics["d91"], dics["a91"])
    assert idx == io and alpha == ao and abs(km - kmo) < 1e-6 * kmo and abs(sse - sseo) < 1e-6 * sseo
    with pytest.raises(ValueError):
        nnls(D, np.full(32, np.nan))           # asarray_chkfinite, like algorithms.py:56
    # the stacked Tikhonov system the reference hands to nnls itself (algorithms.py:264): 92 x 60
    L2 = create_Laplacian_matrix(60, 2)
    A = np.concatenate((D, np.sqrt(0.05) * L2))
    b = np.concatenate((M, np.zeros(60)))
    xs, rs = nnls(A, b)
    xo, ro = O.nnls(A, b)
    assert np.array_equal(xs > 0, xo > 0) and np.allclose(xs, xo, rtol=1e-6, atol=1e-9) and abs(rs - ro) < 1e-8
    with pytest.raises(ValueError):
        nnls(np.concatenate((D, D, D)), np.concatenate((M, M, M)))      # 96 rows that are not [D; L]


def test_plan_cache_distinguishes_regularisation_matrices(golden_voxels, dics):
    """the cache key of the drop-in functions used to sample ~16 entries per array, which are equal
    for I, L1 and L2 — a call with L1 after the same call with I silently reused I's tables.  Same dictionary and
    method, three matrices in a row, each against the oracle; then again in another order."""
    g = golden_voxels
    D = np.ascontiguousarray(dics["d273"][:, :, int(g["nnls_D_index"])])
    M = g["sig"][0] / g["sig"][0, 0]
    mats = {"I": np.eye(60), "L1": create_Laplacian_matrix(60, 1), "L2": create_Laplacian_matrix(60, 2)}
    ref = {k: O.nnls_x2(D, M, L, 1.02) for k, L in mats.items()}
    assert abs(ref["I"][1] - ref["L1"][1]) > 1e-6 and abs(ref["L1"][1] - ref["L2"][1]) > 1e-6     # they DO differ
    for order in (("I", "L1", "L2"), ("L2", "I", "L1")):
        for k in order:
            f, lam, kest = nnls_x2(D, M, mats[k], 1.02)
            fo, lamo, ko = ref[k]
            assert np.array_equal(f > 0, fo > 0), k
            assert np.allclose(f, fo, rtol=1e-6, atol=1e-9) and abs(lam - lamo) < 1e-6 * lamo and abs(kest - ko) < 1e-6, k
            ft = nnls_tik(D, M, mats[k], 0.02)
            assert np.allclose(ft, O.nnls_tik(D, M, mats[k], 0.02), rtol=1e-6, atol=1e-9), k


def test_motor_recon_met2_files(tmp_path):
    """Entry point with NIfTI in/out, config-1-like phantom with an ellipsoidal mask, FA_smooth both ways."""
    ph = make_phantom((12, 10, 3), seed=6, mask_mode="ellipsoid")
    aff = np.diag([2.0, 2.0, 3.0, 1.0])
    nifti_io.save(ph["data"], str(tmp_path / "Data.nii.gz"), affine=aff)
    nifti_io.save(ph["mask"].astype(np.int16), str(tmp_path / "Mask.nii.gz"), affine=aff)
    out = str(tmp_path / "recon_all_X2-I") + "/"
    os.mkdir(out)
    TE = 10.0 * np.arange(1, 33)
    motor_recon_met2(TE, str(tmp_path / "Data.nii.gz"), str(tmp_path / "Mask.nii.gz"), out, 1000.0, "X2", "I", "None",
                     "brute-force", "no", 40.0, -1)
    ref = O.recon_volume(ph["data"], ph["mask"], TE, 1000.0, "X2", "I", "brute-force", num_cores=1)
    names = {"MWF": "MWF", "IEWF": "IEWF", "FWF": "FWF", "T2_M": "T2_M", "T2_IE": "T2_IE", "TWC": "TWC", "FA": "FA",
             "fsol_4D": "fsol_4D", "Est_Signal": "Est_Signal", "reg_param": "reg_param"}
    for fname, key in names.items():
        im = nifti_io.load(out + fname + ".nii.gz")
        assert np.allclose(im.affine, aff)
        got = im.get_fdata()
        assert got.shape == ref[key].shape
        if key in ("FA",):
            assert np.array_equal(got, ref[key])
        elif key in ("MWF", "IEWF", "FWF", "T2_M", "T2_IE"):
            assert np.abs(got - ref[key]).max() < 1e-4, key
        else:
            assert np.abs(got - ref[key]).max() <= 1e-6 * np.abs(ref[key]).max(), key
    # FA_smooth=yes only changes the FA-stage input (motor...:336-346)
    vol = motor_recon_met2(TE, str(tmp_path / "Data.nii.gz"), str(tmp_path / "Mask.nii.gz"), out, 1000.0, "NNLS", "I",
                           "None", "spline", "yes", 40.0, -1)
    assert vol["FA"][ph["mask"] > 0].min() >= 90.0 and not vol["FA"][ph["mask"] == 0].any()
    from scipy.ndimage import gaussian_filter
    masked = ph["data"] * ph["mask"][..., None]
    smooth = np.stack([gaussian_filter(masked[:, :, :, c], 2.0, 0) for c in range(32)], axis=-1)
    ref_s = O.recon_volume(ph["data"], ph["mask"], TE, 1000.0, "NNLS", "I", "spline", num_cores=1, data_fa=smooth)
    assert np.array_equal(vol["FA"], ref_s["FA"]) and np.array_equal(vol["fsol_4D"] > 0, ref_s["fsol_4D"] > 0)
    with pytest.raises(NotImplementedError):
        motor_recon_met2(TE, str(tmp_path / "Data.nii.gz"), str(tmp_path / "Mask.nii.gz"), out, 1000.0, "NNLS", "I",
                         "TV", "spline", "no", 40.0, -1)


def test_gaussian_smooth_bitwise_equal_to_scipy():
    """the FA-stage smoothing (motor...:336-346) on the GPU."""
    from scipy.ndimage import gaussian_filter

    from multicomponent_t2_toolbox_b200 import batched
    rng = np.random.default_rng(3)
    for shape in [(13, 9, 5, 4), (20, 17, 3, 2), (6, 6, 6, 1)]:      # includes axes shorter than the radius (8)
        data = rng.uniform(0.0, 1000.0, shape)
        data[rng.uniform(size=shape[:3]) < 0.3] = 0.0                # masked-out voxels are zeros in the orchestrator
        got = batched.gaussian_smooth(data, sigma=2.0).cpu().numpy()
        for c in range(shape[3]):
            ref = gaussian_filter(data[:, :, :, c], 2.0, 0)
            assert np.array_equal(got[:, :, :, c], ref), (shape, c, np.abs(got[:, :, :, c] - ref).max())


# ------------------------------------------------------------------------------------------------ SURVEY.md §8f rows
@pytest.fixture(scope="module")
def pipeline_gold():
    from conftest import GOLDEN
    return dict(np.load(os.path.join(GOLDEN, "pipeline_nesma_x2.npz")))


def test_nesma_filter_vs_numpy_restatement(pipeline_gold):
    """NESMA denoiser (motor...:305-333) on the GPU against the oracle's loop-for-loop restatement: same summation
    orders, so the result is bitwise equal wherever the 2.5 % similarity decisions agree (they all do here)."""
    from multicomponent_t2_toolbox_b200 import batched
    g = pipeline_gold
    mask = g["mask"].astype(np.int64)
    data = g["data"] * mask[..., None]
    got = batched.nesma_filter(data, mask).cpu().numpy()
    ref = O.nesma_filter(data, mask)
    assert got.shape == ref.shape and not got[mask == 0].any()
    assert np.array_equal(got, ref), np.abs(got - ref).max()
    # 48 echoes (two register slots), non-cubic volume smaller than the window, a mask value that is not 1, and a voxel
    # whose own signal is all zero (RE = NaN -> empty selection -> NaN, like np.mean of an empty array)
    rng = np.random.default_rng(5)
    base = rng.uniform(100.0, 1000.0, (1, 1, 1, 48)) * np.exp(-np.arange(48) / 10.0)
    vol = base * (1.0 + 0.004 * rng.standard_normal((7, 5, 16, 48)))
    m2 = np.ones((7, 5, 16), dtype=np.int64)
    m2[0, 0, 0] = 2
    m2[3, 2, 1] = 0
    vol[3, 2, 1] = 0.0
    vol[4, 4, 4] = 0.0
    got2 = batched.nesma_filter(vol, m2).cpu().numpy()
    ref2 = O.nesma_filter(vol, m2)
    assert np.isnan(ref2[4, 4, 4]).all() and np.isnan(got2[4, 4, 4]).all()
    assert not got2[0, 0, 0].any() and not got2[3, 2, 1].any()
    assert np.array_equal(np.nan_to_num(got2, nan=-1.0), np.nan_to_num(ref2, nan=-1.0))


def test_motor_recon_met2_nesma_smooth_against_reference_run(pipeline_gold, tmp_path):
    """The whole entry point (NIfTI in -> NESMA -> smoothing -> FA spline -> X2-I -> maps -> NIfTI out) against the
    outputs of the UNMODIFIED reference orchestrator run end to end (oracle/make_golden_pipeline.py)."""
    g = pipeline_gold
    nifti_io.save(g["data"], str(tmp_path / "Data.nii.gz"))
    nifti_io.save(g["mask"].astype(np.int16), str(tmp_path / "Mask.nii.gz"))
    out = str(tmp_path / "recon_all_X2-I") + "/"
    os.mkdir(out)
    vol = motor_recon_met2(g["TE"], str(tmp_path / "Data.nii.gz"), str(tmp_path / "Mask.nii.gz"), out, 1000.0, "X2", "I",
                           "NESMA", "spline", "yes", 40.0, -1)
    assert np.array_equal(nifti_io.load(out + "FA.nii.gz").get_fdata(), g["FA"])            # FA index bit-exact
    f = nifti_io.load(out + "fsol_4D.nii.gz").get_fdata()
    assert np.array_equal(f > 0, g["fsol_4D"] > 0)                                           # active sets bit-exact
    scale = np.abs(g["fsol_4D"]).max(axis=3, keepdims=True)
    scale[scale == 0] = 1.0
    assert (np.abs(f - g["fsol_4D"]) / scale).max() < 1e-6
    for k in ("MWF", "IEWF", "FWF", "T2_M", "T2_IE"):
        assert np.abs(nifti_io.load(out + k + ".nii.gz").get_fdata() - g[k]).max() < 1e-4, k
    for k in ("TWC", "Est_Signal", "reg_param"):
        got = nifti_io.load(out + k + ".nii.gz").get_fdata()
        assert np.abs(got - g[k]).max() <= 1e-6 * np.abs(g[k]).max(), k
    # the data behind the reference's mean-spectrum figure (motor...:375-403)
    tab = np.loadtxt(out + "Mean_spectrum_from_all_voxels.txt")
    for col, k in enumerate(("mean_T2_dist", "dist_T2_mean1", "dist_T2_mean2"), start=1):
        assert np.abs(tab[:, col] - g[k]).max() < 1e-6 * g[k].max(), k
        assert np.array_equal(tab[:, col] > 0, g[k] > 0), k
    assert vol["diagnostics"]["nv"] == int((g["mask"] == 1).sum())


def test_roi_estimator_against_oracle(pipeline_gold, tmp_path):
    """ROI-based estimator (motor_recon_met2_real_data_ROI.py:405-445) through its file-level entry point."""
    from multicomponent_t2_toolbox_b200.motor.motor_recon_met2_real_data_ROI import motor_recon_met2_ROIs
    g = pipeline_gold
    nx, ny, nz = g["mask"].shape
    rois = (1 + (np.arange(nx)[:, None, None] // 5) + 3 * (np.arange(ny)[None, :, None] // 6)
            + 0 * np.arange(nz)[None, None, :]).astype(np.int16)
    rois[:, :, :2] = 0
    rois[0, 0, 5] = 40            # a label that only exists outside the mask: empty after ROIs * mask
    nifti_io.save(g["data"], str(tmp_path / "Data.nii.gz"))
    nifti_io.save(g["mask"].astype(np.int16), str(tmp_path / "Mask.nii.gz"))
    nifti_io.save(rois, str(tmp_path / "ROIs.nii.gz"))
    out = str(tmp_path / "recon_all_X2-L2_ROI-based") + "/"
    os.mkdir(out)
    vol = motor_recon_met2_ROIs(g["TE"], str(tmp_path / "Data.nii.gz"), str(tmp_path / "Mask.nii.gz"),
                                str(tmp_path / "ROIs.nii.gz"), out, 1000.0, "L2", "None", "spline", "no", 40.0, -1)
    mask = g["mask"].astype(np.int64)
    data = g["data"] * mask[..., None]
    gr = O._grids("X2", "L2", "spline", 40.0, 32, 10.0, 1000.0)
    Dic = O.create_Dic_3D(60, gr["T2s"], gr["T1s"], 32, 10.0, gr["alpha_values"], 1000.0)
    keep = np.unique(rois)
    keep = keep[(keep != 0) & (keep != 40)]
    rois_ref = np.where(np.isin(rois, keep), rois, 0)
    ref = O.roi_estimates(data, mask, rois_ref, vol["FA_index"], Dic, gr["L"], gr["T2s"], gr["ind_m"], gr["ind_t"],
                          gr["ind_csf"])
    roi = vol["roi"]
    sel = np.isin(roi["roi_values"], keep)
    assert np.array_equal(roi["roi_values"][sel], ref["roi_values"]) and roi["roi_values"][~sel].tolist() == [40]
    assert roi["counts"][~sel].tolist() == [0] and not np.isfinite(roi["mean_signal"][~sel]).any()
    assert np.array_equal(roi["fsol_ROIs"][sel] > 0, ref["fsol_ROIs"] > 0)
    assert np.abs(roi["fsol_ROIs"][sel] - ref["fsol_ROIs"]).max() < 1e-6 * ref["fsol_ROIs"].max()
    for k in ("MWF", "IEWF", "FWF", "T2M", "T2IE"):
        assert np.abs(roi[k + "_ROIs"][sel] - ref[k]).max() < 1e-4, k
    assert np.allclose(roi["TWC_ROIs"][sel], ref["TWC"], rtol=1e-6)
    assert np.allclose(roi["reg_opt"][sel], ref["reg_opt"], rtol=1e-5) and np.allclose(roi["k_est"][sel], ref["k_est"], rtol=1e-6)
    # ... and against what the UNMODIFIED reference's motor_recon_met2_ROIs wrote for the same data, mask and the six
    # in-mask labels (oracle/make_golden_roi.py -> tests/golden/roi_x2_l2.npz)
    from conftest import GOLDEN
    gr = dict(np.load(os.path.join(GOLDEN, "roi_x2_l2.npz")))
    assert np.array_equal(gr["labels"].astype(np.int64), roi["roi_values"][sel])
    assert np.array_equal(roi["fsol_ROIs"][sel] > 0, gr["spectra"] > 0)
    assert np.abs(roi["fsol_ROIs"][sel] - gr["spectra"]).max() < 1e-6 * gr["spectra"].max()
    for col, k in enumerate(("MWF", "IEWF", "FWF", "T2M", "T2IE")):
        assert np.abs(roi[k + "_ROIs"][sel] - gr["table_values"][:, col]).max() < 1e-4, k
    assert np.allclose(roi["TWC_ROIs"][sel], gr["table_values"][:, 5], rtol=1e-6)
    assert np.allclose(np.loadtxt(out + "table_MWF.csv", delimiter=",")[sel], ref["MWF"], atol=1e-4)
    assert np.loadtxt(out + "table_Spectra.csv", delimiter=",").shape == (len(roi["roi_values"]), 60)
    assert os.path.exists(out + "ROI_%d/table_values.csv" % int(keep[0]))
