"""Seeded Monte-Carlo sanity run against the ONLY numbers the reference itself holds for this path:
scripts_synthetic_data_evaluation/Paper_Comparison/Results/SNRs_{50_150,150_300}/All_methods_10000iters/
table_errors.txt:3-12 (MWF mean absolute error) and table_regularization.txt:3-12 (mean lambda) — ten methods, 10 000
two-lobe voxels per SNR band, generated and fitted as in evaluate_all_methods_two_lobes_SNR*.py (tools/montecarlo.py).
The tables come from unseeded draws, so the bands are statistical: MAE within 5 % (standard error of a 10 000-voxel
mean: < 1 %); mean lambda within 5 % or four standard errors of the table's own STD / sqrt(N), whichever is larger."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_montecarlo_tables():
    import montecarlo
    N = 10000
    rec = montecarlo.run(N, seed=0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "montecarlo.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    bad = []
    for band, b in rec["bands"].items():
        for name, r in b["methods"].items():
            assert r["status_nonzero"] == 0, (band, name, r)
            if abs(r["MAE"] / r["reference_MAE"] - 1.0) > 0.05:
                bad.append((band, name, "MAE", r["MAE"], r["reference_MAE"]))
            if r["reference_mean_lambda"] > 0:
                tol = max(0.05 * r["reference_mean_lambda"], 4.0 * r["reference_std_lambda"] / np.sqrt(N) * np.sqrt(2.0))
                if abs(r["mean_lambda"] - r["reference_mean_lambda"]) > tol:
                    bad.append((band, name, "mean_lambda", r["mean_lambda"], r["reference_mean_lambda"], tol))
    assert not bad, bad
