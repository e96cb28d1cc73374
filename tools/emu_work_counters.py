"""Work per voxel of the T2 / FA kernels from the SIMT emulator's counters (tests/emu): explicit fma() calls,
shared-memory accesses, __syncwarp and warp collectives (shuffles, votes, reductions, FP64 MMAs), warp level
(per-thread events / 32), averaged over config-2 voxels.  A cost model for comparing kernel variants without a GPU —
not a timing.  Usage: python tools/emu_work_counters.py [n_voxels] [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import met2_oracle as O  # noqa: E402
from emu import emu  # noqa: E402


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "config2_subset.npz")))
    gr = O._grids("X2", "I", "spline", 40.0, 32, 10.0, 1000.0)
    Dic = O.create_Dic_3D(60, gr["T2s"], gr["T1s"], 32, 10.0, gr["alpha_values"], 1000.0)
    DicLR = O.create_Dic_3D(60, gr["T2s"], gr["T1s"], 32, 10.0, gr["alpha_spline"], 1000.0)
    sel = np.arange(0, len(g["sig"]), len(g["sig"]) // nv)[:nv]
    sig, fa = g["sig"][sel], g["fa_idx"][sel].astype(np.int32)
    uniq, inv = np.unique(fa, return_inverse=True)
    Dc = np.ascontiguousarray(Dic[:, :, uniq])
    LI = gr["L"]
    Linv = O._grids("X2", "InvT2", "spline", 40.0, 32, 10.0, 1000.0)["L"]
    rows = {}

    def rec(name, out):
        c = out["counters"]
        rows[name] = {k: round(c[k] / 32.0 / nv, 1) for k in ("fma", "smem", "syncwarp", "collectives")}
        rows[name]["launches"] = c["launches"]
        print("%-34s fma %8.0f  smem %8.0f  syncwarp %7.0f  collectives %7.0f" % (
            name, rows[name]["fma"], rows[name]["smem"], rows[name]["syncwarp"], rows[name]["collectives"]), flush=True)

    t2 = lambda method, L, **kw: emu.t2_fit(sig, inv.astype(np.int32), Dc, L, gr["T2s"], method, lambdas=gr["lambda_reg"], **kw)
    rec("X2-I default (tables, warm)", t2("X2", LI, flags=16))
    rec("X2-I no tables", t2("X2", LI, flags=0))
    rec("X2-I cold starts (reference path)", t2("X2", LI, flags=4))
    rec("X2-I echo space (experimental)", t2("X2", LI, flags=16, echo=True))
    rec("X2-InvT2 default", t2("X2", Linv))
    rec("X2-InvT2 echo space", t2("X2", Linv, echo=True))
    rec("NNLS", t2("NNLS", LI))
    rec("T2SPARC-I (60 bins)", t2("T2SPARC", LI))
    rec("L_curve-I", t2("L_curve", LI))
    rec("BayesReg-I", t2("BayesReg", LI))
    rec("GCV-I", t2("GCV", LI))
    rec("FA spline (15 knots -> 273)", emu.fa_fit(sig, Dic, gr["alpha_values"], DicLR, gr["alpha_spline"]))
    res = dict(voxels=nv, unit="warp-level events per voxel (per-thread count / 32); table kernels included", rows=rows)
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
