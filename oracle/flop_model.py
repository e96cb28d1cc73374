"""TEST INFRASTRUCTURE ONLY — algorithmic flops per voxel F_alg of the REFERENCE formulation.

Counts, with the instrumented Lawson-Hanson restatement `met2_oracle.lh_nnls` (Householder/Givens QR on the m x n or
(m+n) x n system, 2 flops per multiply-add), every NNLS solve the reference performs for one voxel of the benchmark
workload (config 2: spline FA = 15 + 1 plain solves; X2-I = 1 plain + Brent evaluations + 1 final augmented solves),
on seeded phantom voxels.  The Brent abscissae are taken from the real SciPy path so the solve sequence is the
reference's.  Writes oracle/F_ALG.json; bench.py quotes the numbers.

    python oracle/flop_model.py [n_voxels = 1024]      (all local cores; ~10 minutes for 1 024 voxels on 8 cores)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import met2_oracle as O  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402
from scipy.optimize import fminbound  # noqa: E402


_G = {}


def _voxel(v):
    g, Dic, DicLR, L, sig = _G["g"], _G["Dic"], _G["DicLR"], _G["L"], _G["sig"]
    M = sig[v]
    c = O.FlopCounter()
    for i in range(15):
        O.lh_nnls(np.ascontiguousarray(DicLR[:, :, i]), M, counter=c)
    idx, _, _, _ = O.spline_optimal_FA(M, DicLR, Dic, g["alpha_spline"], g["alpha_values"])
    D = np.ascontiguousarray(Dic[:, :, idx])
    O.lh_nnls(D, M, counter=c)
    fa = c.flops
    c = O.FlopCounter()
    Mn = M / M[0]
    f0, _, _ = O.lh_nnls(D, Mn, counter=c)
    SSE = np.sum((D @ f0 - Mn) ** 2)
    Maug = np.concatenate((Mn, np.zeros(60)))

    def obj(x):
        f, _, _ = O.lh_nnls(np.concatenate((D, np.sqrt(x) * L)), Maug, counter=c)
        return abs(np.sum((D @ f - Mn) ** 2) - 1.02 * SSE) / SSE
    reg = fminbound(obj, 0.0, 10.0, xtol=1e-5, maxfun=300)
    O.lh_nnls(np.concatenate((D, np.sqrt(reg) * L)), Maug, counter=c)
    return fa, c.flops, c.solves, c.outer


def main():
    import multiprocessing as mp
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ph = make_phantom((96, 96, 1), seed=2, fa_mode="b1")
    sig = ph["data"].reshape(-1, 32)
    rng = np.random.default_rng(0)
    pick = rng.choice(sig.shape[0], nv, replace=False)
    g = O._grids("X2", "I", "spline", 40.0, 32, 10.0, 1000.0)
    _G.update(g=g, sig=sig, L=g["L"],
              Dic=O.create_Dic_3D(60, g["T2s"], g["T1s"], 32, 10.0, g["alpha_values"], 1000.0),
              DicLR=O.create_Dic_3D(60, g["T2s"], g["T1s"], 32, 10.0, g["alpha_spline"], 1000.0))
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        res = np.array(pool.map(_voxel, [int(v) for v in pick], chunksize=4), dtype=np.float64)
    fa_flops, t2_flops, t2_solves, t2_outer = res.T
    out = dict(workload="config 2: spline FA + X2-I, nTE=32, nT2=60", n_voxels=nv,
               F_alg_fa_spline=float(np.mean(fa_flops)), F_alg_fa_spline_sd=float(np.std(fa_flops)),
               F_alg_t2_x2_I=float(np.mean(t2_flops)), F_alg_t2_x2_I_sd=float(np.std(t2_flops)),
               F_alg_t2_x2_I_min=float(np.min(t2_flops)), F_alg_t2_x2_I_max=float(np.max(t2_flops)),
               nnls_solves_x2=float(np.mean(t2_solves)), outer_iterations_x2=float(np.mean(t2_outer)),
               model="SURVEY.md 8d: gradient 2(m-p)|Z|, Householder build 3(m-p) + apply 4(m-p) per column, "
                     "back-substitution p^2, Givens removal (6n+12) per rotated row, interpolation 3p")
    with open(os.path.join(HERE, "F_ALG.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
