"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference orchestrator end to end and stores what it produced.

`motor.motor_recon_met2_real_data.motor_recon_met2` (reference :165-506) is imported read-only through
oracle/ref_shim.py and executed on a small seeded volume with denoise='NESMA', FA_smooth='yes', FA_method='spline',
reg_method='X2', reg_matrix='I'.  What the reference needs and this image lacks is replaced by recording stand-ins
that do no arithmetic: `nibabel` (load/save of .npy files, affine passed through), `matplotlib` (a MagicMock whose
`plot` calls are recorded — that is how the three mean-spectrum curves of :397-399 are captured), `progressbar`.
Everything numerical (NESMA loop, Gaussian smoothing, joblib row workers, metrics loop, mean-spectrum NNLS/X2) is
the reference's own code.  Output: tests/golden/pipeline_nesma_x2.npz (inputs + all ten output volumes + the curves).

    python oracle/make_golden_pipeline.py
"""
import os
import sys
import tempfile
import types
import warnings
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

warnings.simplefilter("ignore")
OUT = os.path.join(ROOT, "tests", "golden", "pipeline_nesma_x2.npz")


def piecewise_phantom(shape=(14, 12, 10), seed=7):
    """Three tissue classes in blocks + mild noise (SNR ~ 400), so that the NESMA similarity test (RE < 2.5 %) accepts
    many neighbours; ellipsoidal mask."""
    from multicomponent_t2_toolbox_b200.phantom import epg_signal_batch
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    cls = (np.arange(nx)[:, None, None] // 5 + np.arange(ny)[None, :, None] // 6 + np.arange(nz)[None, None, :] // 5) % 3
    mwf = np.array([0.08, 0.15, 0.22])[cls]
    t2m = np.array([18.0, 25.0, 30.0])[cls]
    t2ie = np.array([65.0, 75.0, 85.0])[cls]
    fa = np.array([150.0, 165.0, 175.0])[cls] + rng.uniform(-1.0, 1.0, shape)
    V = nx * ny * nz
    T1 = np.full(V, 1000.0)
    sig = mwf.ravel()[:, None] * epg_signal_batch(32, 10.0, T1, t2m.ravel(), fa.ravel())
    sig += (0.97 - mwf.ravel())[:, None] * epg_signal_batch(32, 10.0, T1, t2ie.ravel(), fa.ravel())
    sig += 0.03 * epg_signal_batch(32, 10.0, T1, np.full(V, 2000.0), fa.ravel())
    sig *= 1000.0 * (1.0 - np.exp(-1.0))
    sigma = sig[:, 0] / 400.0
    noisy = np.sqrt((sig + rng.standard_normal(sig.shape) * sigma[:, None]) ** 2 +
                    (rng.standard_normal(sig.shape) * sigma[:, None]) ** 2)
    gx = np.linspace(-1, 1, nx)[:, None, None]
    gy = np.linspace(-1, 1, ny)[None, :, None]
    gz = np.linspace(-1, 1, nz)[None, None, :]
    mask = ((gx ** 2 + gy ** 2 + gz ** 2) <= 1.0).astype(np.float64)
    return noisy.reshape(nx, ny, nz, 32), mask


class _Img:
    def __init__(self, arr, affine):
        self._a, self.affine = arr, affine

    def get_fdata(self):
        return np.array(self._a, dtype=np.float64)


def main():
    saved = {}
    nib = types.ModuleType("nibabel")
    nib.load = lambda path: _Img(np.load(path), np.eye(4))
    nib.Nifti1Image = lambda arr, affine: _Img(np.array(arr), affine)
    nib.save = lambda img, path: saved.__setitem__(os.path.basename(path).replace(".nii.gz", ""), img._a)
    sys.modules["nibabel"] = nib
    mpl = mock.MagicMock()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):
        sys.modules[name] = mpl
    mpl.pyplot = mpl
    mpl.rcParams = {}
    import ref_shim
    R = ref_shim.load_reference()
    motor = R["motor"]
    plt = motor.plt
    data, mask = piecewise_phantom()
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, "data.npy"), data)
    np.save(os.path.join(tmp, "mask.npy"), mask)
    TE = 10.0 * np.arange(1, 33)
    motor.motor_recon_met2(TE, os.path.join(tmp, "data.npy"), os.path.join(tmp, "mask.npy"), tmp + "/", 1000.0, "X2", "I",
                           "NESMA", "spline", "yes", 40.0, 4)
    curves = [np.asarray(c.args[1], dtype=np.float64) for c in plt.plot.call_args_list
              if len(c.args) >= 2 and np.ndim(c.args[1]) == 1 and len(c.args[1]) == 60]
    assert len(curves) >= 3, len(curves)
    out = dict(data=data, mask=mask, TE=TE, mean_T2_dist=curves[0], dist_T2_mean1=curves[1], dist_T2_mean2=curves[2])
    for k in ("MWF", "IEWF", "FWF", "T2_M", "T2_IE", "TWC", "FA", "fsol_4D", "Est_Signal", "reg_param"):
        out[k] = saved[k]
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB; masked voxels", int(mask.sum()), "saved:", sorted(saved))


if __name__ == "__main__":
    main()
