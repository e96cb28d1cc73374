#!/bin/bash
# Round-2 GPU call #23: per-kernel launch bounds of the echo kernels (T2SPARC and BayesReg-60 at 1024 threads) — parity
# tests of the methods concerned, times of all configs; L-curve with fewer resident warps (MET2_T2_WARPS sweep).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "echo or methods_subset or config4 or t2sparc or regularised_fit_vs_oracle or golden_vectors or lcurve or montecarlo or full_size" > $O/r23_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r23_pytest.log
timeout 900 python tools/gpu_configs.py > $O/r23_configs.log 2>&1
cp $O/configs.json $O/r23_configs.json
for w in 8 12 16; do
  MET2_T2_WARPS=$w METHOD=L_curve SHAPE=96,96,60 timeout 200 python tools/gpu_time.py > $O/r23_lcurve_warps_$w.log 2>&1
done
tail -3 $O/r23_pytest.log; cat $O/r23_configs.log | cut -c1-150; tail -2 $O/r23_lcurve_warps_*.log
