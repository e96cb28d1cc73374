"""Module path of the reference's motor/motor_recon_met2_real_data.py.

`motor_recon_met2` keeps the reference's signature and NIfTI outputs (motor/motor_recon_met2_real_data.py:165-506) but
replaces the two joblib voxel loops (Steps 2 and 3, :349-373 and :428-441) and the Python metrics loop (Step 4,
:443-472) by one gather -> batched GPU fit -> scatter (pipeline.recon_arrays).  Host-side preprocessing that is not on
the accelerated path: the TV denoiser (:293-303, scikit-image) is out of scope.  The NESMA denoiser
(:305-333) and the Gaussian smoothing for the FA stage (:336-346) run on the GPU (met2_nesma_filter,
met2_gaussian_smooth); the data behind the mean-spectrum figure (:375-403) is computed on the GPU and written as a text
table (Mean_spectrum_from_all_voxels.txt) instead of a PNG (no matplotlib here).
"""
import os

import numpy as np

from .. import batched, nifti_io, pipeline
from ..reference_api import create_Laplacian_matrix, fitting_slice_T2  # noqa: F401

OUTPUTS = ("MWF", "IEWF", "FWF", "T2_M", "T2_IE", "TWC", "FA", "fsol_4D", "Est_Signal", "reg_param")


def motor_recon_met2(TE_array, path_to_data, path_to_mask, path_to_save_data, TR, reg_method, reg_matrix, denoise,
                     FA_method, FA_smooth, myelin_T2, num_cores):
    """Same arguments as the reference.  `num_cores` is accepted and ignored (the fit runs on the current CUDA device;
    multi-GPU runs shard voxel slabs, see pipeline.recon_arrays)."""
    img = nifti_io.load(path_to_data)
    data = img.get_fdata().astype(np.float64, copy=False)
    mask = nifti_io.load(path_to_mask).get_fdata().astype(np.int64, copy=False)
    print('--------- Data shape -----------------')
    print(data.shape)
    print('--------------------------------------')
    nx, ny, nz, nt = data.shape
    np.multiply(data, mask[..., None], out=data)   # motor...:178-180, all echoes in one pass (5x faster than per echo)
    data[data < 0.0] = 0.0
    if reg_matrix not in ('I', 'L1', 'L2', 'InvT2'):
        print('Error: Wrong reg_matrix option!')
        raise SystemExit(1)
    if denoise == 'TV':
        raise NotImplementedError("denoise=TV (scikit-image total variation) is host preprocessing outside the "
                                  "accelerated path (SURVEY.md §8f); run it beforehand and pass --denoise None")
    if denoise == 'NESMA':
        print('Step #1: Denoising using the NESMA filter:')
        data = batched.nesma_filter(data, mask).cpu().numpy()
    print('Step #2: Estimation of flip angles:')
    data_fa = None
    if FA_smooth == 'yes':
        # per-echo Gaussian smoothing, sigma = 2 (motor...:336-346) — on the GPU, bitwise equal to scipy.ndimage
        data_fa = batched.gaussian_smooth(data, sigma=2.0)
    print('Step #3: Estimation of T2 spectra:')
    vol = pipeline.recon_arrays(data, mask, np.asarray(TE_array, dtype=np.float64), TR, reg_method, reg_matrix,
                                FA_method, myelin_T2=myelin_T2, data_fa=data_fa, diagnostics=True)
    print('Step #4: Estimation of quantitative metrics')
    for name in OUTPUTS:
        nifti_io.save(vol[name], os.path.join(path_to_save_data, name + '.nii.gz') if not path_to_save_data.endswith('/')
                      else path_to_save_data + name + '.nii.gz', affine=img.affine)
    dg = vol.get("diagnostics")
    if dg is not None:
        np.savetxt(os.path.join(path_to_save_data, 'Mean_spectrum_from_all_voxels.txt'),
                   np.column_stack([dg["T2s"], dg["mean_T2_dist"], dg["dist_T2_mean1"], dg["dist_T2_mean2"]]),
                   header="T2(ms)  mean_T2_dist(all voxels, NNLS)  dist_T2_mean1(mean signal, NNLS)  "
                          "dist_T2_mean2(mean signal, NNLS-X2-I)")
    print('Done!')
    return vol
