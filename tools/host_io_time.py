"""Host-side wall time of the file-level entry point around the GPU fit, no GPU needed:
load Data/Mask (.nii.gz) -> mask + clamp -> gather of the masked voxel list | (fit) | scatter -> ten NIfTI outputs.
Usage: python tools/host_io_time.py [out.json]   (config-2 sized volume: 96x96x60x32, 60 T2 bins)"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multicomponent_t2_toolbox_b200 import nifti_io, pipeline  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    nx, ny, nz, nt, npc = 96, 96, 60, 32, 60
    mask = (rng.random((nx, ny, nz)) < 0.6).astype(np.int16)
    data = (1000.0 * np.exp(-np.arange(nt) / 8.0) * (1 + 0.01 * rng.standard_normal((nx, ny, nz, nt)))).astype(np.float32)
    d = tempfile.mkdtemp()
    nifti_io.save(data, d + "/Data.nii.gz")
    nifti_io.save(mask, d + "/Mask.nii.gz")
    T = {}

    def lap(name, t0):
        T[name] = round(time.time() - t0, 3)

    t = time.time()
    img = nifti_io.load(d + "/Data.nii.gz")
    vol = img.get_fdata()
    m = nifti_io.load(d + "/Mask.nii.gz").get_fdata().astype(np.int64)
    lap("load_s", t)
    t = time.time()
    np.multiply(vol, m[..., None], out=vol)
    vol[vol < 0.0] = 0.0
    flat, sig = pipeline.masked_voxel_list(vol, m)
    lap("mask_and_gather_s", t)
    V = len(flat)
    fsol = rng.random((V, npc)) * (rng.random((V, npc)) < 0.1)
    est = rng.random((V, nt))
    t = time.time()
    f4 = np.zeros((nx * ny * nz, npc))
    f4[flat] = fsol
    s4 = np.zeros((nx * ny * nz, nt))
    s4[flat] = est
    maps = [np.zeros(nx * ny * nz) for _ in range(8)]
    for a in maps:
        a[flat] = est[:, 0]
    lap("scatter_s", t)
    t = time.time()
    nifti_io.save(f4.reshape(nx, ny, nz, npc), d + "/fsol_4D.nii.gz", affine=img.affine)
    nifti_io.save(s4.reshape(nx, ny, nz, nt), d + "/Est_Signal.nii.gz", affine=img.affine)
    for i, a in enumerate(maps):
        nifti_io.save(a.reshape(nx, ny, nz), d + "/map%d.nii.gz" % i, affine=img.affine)
    lap("save_10_outputs_s", t)
    t = time.time()
    nifti_io.save(f4.reshape(nx, ny, nz, npc), d + "/one.nii.gz", affine=img.affine, threads=1)
    lap("save_fsol_4D_one_thread_s", t)
    T.update(voxels=int(V), cores=os.cpu_count(), volume="%dx%dx%dx%d, %d T2 bins" % (nx, ny, nz, nt, npc),
             note="host side only (no GPU in this container); deflate level 1, one gzip member, 4 MiB chunks")
    print(json.dumps(T))
    if len(sys.argv) > 1:
        json.dump(T, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
