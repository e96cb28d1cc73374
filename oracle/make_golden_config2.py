"""TEST INFRASTRUCTURE ONLY — writes tests/golden/config2_subset.npz from OUTPUTS OF THE UNMODIFIED REFERENCE.

BASELINE.json configs[1] (96x96x60 phantom, FA spline + X2, reg_matrix I) is too large for a CPU oracle run inside the
test-suite, so SURVEY.md §8d asks for parity on a fixed random subset of >= 20 000 voxels.  This script draws that
subset (20 480 voxels, seeded) from the seed-2 config-2 phantom, runs the reference's own row workers on it
(flip_angle_algorithms/fa_estimation.py:35 and motor/motor_recon_met2_real_data.py:113, imported read-only through
oracle/ref_shim.py) on all local cores, and stores inputs + outputs:

    sig       [S, 32]  float64   raw signals of the picked voxels (the noise makes them incompressible: 5 MB)
    pick      [S]      int64     flat voxel index in the 96x96x60 volume
    fa_idx    [S]      int16     reference FA index on the 273-grid
    km        [S]      float64   sum of the FA-stage spectrum
    reg       [S]      float64   reg_param (= k_est for X2, motor...:141-143)
    support   [S, 8]   uint8     np.packbits(f > 0) over the 60 T2 bins (padded to 64)
    f_nz      [nnz]    float64   non-zero spectrum coefficients, voxel after voxel, bin order

Runs only where /root/reference exists (the build container).

    python oracle/make_golden_config2.py
"""
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

warnings.simplefilter("ignore")
OUT = os.path.join(ROOT, "tests", "golden", "config2_subset.npz")
S = 20480
CHUNK = 64
_G = {}


def _work(lo):
    R = _G["R"]
    sig = _G["sig"][lo:lo + CHUNK]
    nx = sig.shape[0]
    ok = np.ones(nx)
    FA, idx, KM, _ = R["fa"].fitting_slice_FA_spline_method(_G["Dic15"], _G["Dic273"], sig, ok, _G["a15"], nx, _G["a273"])
    f, s, reg = R["motor"].fitting_slice_T2(ok, sig, idx, nx, _G["Dic273"], _G["lam"], 60, 32, "X2", np.eye(60), None)
    return lo, idx, KM, f, reg


def main():
    t0 = time.time()
    R = ref_shim.load_reference()
    nte, tau, TR = 32, 10.0, 1000.0
    T2s = np.logspace(np.log10(10.0), np.log10(2000.0), 60)
    T1s = 1000.0 * np.ones(60)
    a273 = np.linspace(90.0, 180.0, 273)
    a15 = np.linspace(90.0, 180.0, 15)
    ph = make_phantom((96, 96, 60), n_echoes=nte, tau=tau, TR=TR, seed=2, fa_mode="b1")
    allsig = ph["data"].reshape(-1, nte)
    pick = np.sort(np.random.default_rng(20480).choice(allsig.shape[0], S, replace=False))
    sig = np.ascontiguousarray(allsig[pick])
    print("phantom %.0f s" % (time.time() - t0), flush=True)
    lam = np.zeros(50)
    lam[1:] = np.logspace(-8, 1, 49)
    _G.update(R=R, sig=sig, a15=a15, a273=a273, lam=lam,
              Dic273=R["epg"].create_Dic_3D(60, T2s, T1s, nte, tau, a273, TR),
              Dic15=R["epg"].create_Dic_3D(60, T2s, T1s, nte, tau, a15, TR))
    print("dictionaries %.0f s" % (time.time() - t0), flush=True)
    fa_idx = np.zeros(S, dtype=np.int16)
    km = np.zeros(S)
    reg = np.zeros(S)
    f = np.zeros((S, 60))
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        for lo, idx, KM, fc, rc in pool.imap_unordered(_work, range(0, S, CHUNK)):
            n = len(idx)
            fa_idx[lo:lo + n], km[lo:lo + n], reg[lo:lo + n], f[lo:lo + n] = idx.astype(np.int16), KM, rc, fc
    sup = f > 0
    np.savez_compressed(OUT, sig=sig, pick=pick, fa_idx=fa_idx, km=km, reg=reg,
                        support=np.packbits(np.pad(sup, ((0, 0), (0, 4))), axis=1), f_nz=f[sup],
                        seed=2, shape=np.array([96, 96, 60]))
    print("wrote %s (%.1f MB) in %.0f s; mean support %.1f" % (OUT, os.path.getsize(OUT) / 1e6, time.time() - t0,
                                                               sup.sum(1).mean()))


if __name__ == "__main__":
    main()
