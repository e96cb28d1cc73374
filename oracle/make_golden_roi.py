"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference ROI-based estimator end to end and stores what it produced.

`motor.motor_recon_met2_real_data_ROI.motor_recon_met2_ROIs` (reference :152-504) is imported read-only and executed on
the seeded volume of tests/golden/pipeline_nesma_x2.npz with a 6-label ROI image, FA_method='spline', FA_smooth='no',
denoise='None', reg_matrix='L2'.  Stand-ins that do no arithmetic replace what this image lacks: `nibabel` (load/save
of .npy files), `matplotlib` / `joypy` (MagicMock), `progressbar`; `np.int` (removed from NumPy 1.24, used at :386 and
:415) is aliased to `int` like the `xrange` alias of oracle/ref_shim.py.  The per-ROI mean signals / kernels, the X2
fits (factor 1.01) and the metrics are the reference's own code; its CSV outputs are stored in
tests/golden/roi_x2_l2.npz.

    python oracle/make_golden_roi.py
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
OUT = os.path.join(ROOT, "tests", "golden", "roi_x2_l2.npz")


def roi_image(shape):
    nx, ny, nz = shape
    rois = (1 + (np.arange(nx)[:, None, None] // 5) + 3 * (np.arange(ny)[None, :, None] // 6)
            + 0 * np.arange(nz)[None, None, :]).astype(np.float64)
    rois[:, :, :2] = 0
    return rois


class _Img:
    def __init__(self, arr, affine):
        self._a, self.affine = arr, affine

    def get_fdata(self):
        return np.array(self._a, dtype=np.float64)


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "pipeline_nesma_x2.npz"))
    data, mask, TE = g["data"], g["mask"], g["TE"]
    rois = roi_image(mask.shape)
    nib = types.ModuleType("nibabel")
    nib.load = lambda path: _Img(np.load(path), np.eye(4))
    nib.Nifti1Image = lambda arr, affine: _Img(np.array(arr), affine)
    nib.save = lambda img, path: None
    sys.modules["nibabel"] = nib
    mpl = mock.MagicMock()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "joypy"):
        sys.modules[name] = mpl
    mpl.pyplot = mpl
    mpl.rcParams = {}
    mpl.joyplot.return_value = (mock.MagicMock(), mock.MagicMock())   # fig, axes of the ridge plot
    mpl.xticks.return_value = ([], [])
    np.int = int                      # NumPy < 1.24 alias used by the reference
    import ref_shim
    ref_shim.load_reference()
    import motor.motor_recon_met2_real_data_ROI as roi_mod
    tmp = tempfile.mkdtemp() + "/"
    for name, arr in (("data", data), ("mask", mask), ("rois", rois)):
        np.save(tmp + name + ".npy", arr)
    roi_mod.motor_recon_met2_ROIs(TE, tmp + "data.npy", tmp + "mask.npy", tmp + "rois.npy", tmp, 1000.0, "L2", "None",
                                  "spline", "no", 40.0, 4)
    out = dict(rois=rois, MWF=np.loadtxt(tmp + "table_MWF.csv", delimiter=","),
               spectra=np.loadtxt(tmp + "table_Spectra.csv", delimiter=","),
               labels=np.loadtxt(tmp + "ROI_labels.csv", delimiter=","))
    vals = []
    for lab in out["labels"].astype(int):
        tab = np.genfromtxt(tmp + "ROI_%d/table_values.csv" % lab, delimiter=",", dtype=str)
        vals.append([float(v) for v in tab[:, 1]])
    out["table_values"] = np.array(vals)        # MWF, IEWF, FWF, T2M, T2IE, TWC per ROI
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "labels", out["labels"], "MWF", out["MWF"])


if __name__ == "__main__":
    main()
