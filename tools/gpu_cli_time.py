"""Wall time of the file-level entry points on the config-2 volume (96x96x60x32, brain-shaped mask): the CLI as a
subprocess (interpreter start + imports + load + fit + ten NIfTI outputs) and motor_recon_met2 called twice in-process
(first call: CUDA contexts, library load; second: steady state), with the load / fit / save split of the second call.

    gpurun --timeout 600 -- 'timeout 500 python tools/gpu_cli_time.py > gpurun_out/cli_time.log 2>&1'
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multicomponent_t2_toolbox_b200 import nifti_io  # noqa: E402
from multicomponent_t2_toolbox_b200.phantom import make_phantom  # noqa: E402

d = tempfile.mkdtemp() + "/"
ph = make_phantom((96, 96, 60), seed=2, fa_mode="b1", backend="gpu")
data = ph["data"].astype(np.float32)
xx, yy, zz = np.meshgrid(np.linspace(-1, 1, 96), np.linspace(-1, 1, 96), np.linspace(-1, 1, 60), indexing="ij")
mask = ((xx / 0.85) ** 2 + (yy / 0.95) ** 2 + (zz / 0.9) ** 2 < 1.0).astype(np.int16)      # ~55 % of the box
nifti_io.save(data, d + "Data.nii.gz")
nifti_io.save(mask, d + "Mask.nii.gz")
rep = dict(volume="96x96x60x32", masked_voxels=int(mask.sum()), cores=os.cpu_count())
cmd = [sys.executable, os.path.join(ROOT, "run_real_data_script.py"), "--path_to_folder", d, "--input", "Data.nii.gz",
       "--mask", "Mask.nii.gz", "--minTE", "10", "--nTE", "32", "--TR", "1000", "--FA_method", "spline", "--FA_smooth", "yes",
       "--denoise", "None", "--reg_method", "X2", "--reg_matrix", "I", "--numcores", "1", "--myelin_T2_cutoff", "40",
       "--savefig", "no", "--savefig_slice", "30"]
for i in range(2):
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    rep["cli_subprocess_wall_s_run%d" % (i + 1)] = round(time.time() - t, 3)
    if r.returncode != 0:
        rep["cli_error"] = (r.stdout + r.stderr)[-1500:]
        break
from multicomponent_t2_toolbox_b200.motor.motor_recon_met2_real_data import motor_recon_met2  # noqa: E402
TE = 10.0 * np.arange(1, 33)
os.makedirs(d + "out", exist_ok=True)
for i in range(2):
    t = time.time()
    motor_recon_met2(TE, d + "Data.nii.gz", d + "Mask.nii.gz", d + "out/", 1000.0, "X2", "I", "None", "spline", "yes", 40.0, 1)
    rep["motor_recon_met2_wall_s_call%d" % (i + 1)] = round(time.time() - t, 3)
t = time.time()
img = nifti_io.load(d + "Data.nii.gz")
vol = img.get_fdata()
m = nifti_io.load(d + "Mask.nii.gz").get_fdata()
rep["load_only_s"] = round(time.time() - t, 3)
print(json.dumps(rep, indent=1))
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "cli_time.json"), "w"), indent=1)
