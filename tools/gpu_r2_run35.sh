#!/bin/bash
# Round-2 GPU call #35: GCV Jacobi without identity rotations (skipped pairs and rounds) — GCV tests and times.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "gcv or GCV or config3b or methods_subset or montecarlo" > $O/r35_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r35_pytest.log
WHICH=3a,3b timeout 600 python tools/gpu_configs.py > $O/r35_configs.log 2>&1
METHOD=GCV timeout 200 python tools/gpu_time.py > $O/r35_gcv_I.log 2>&1
tail -n 3 $O/r35_pytest.log; cut -c1-150 $O/r35_configs.log; grep "^{'fa_ms'" $O/r35_gcv_I.log | tail -n 1
