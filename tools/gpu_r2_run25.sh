#!/bin/bash
# Round-2 GPU call #25: + unrolled T update and Ct^T v product in the X2 / T2SPARC kernels — parity subset, bench, times.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "echo or methods_subset or config4_subset or config2_subset or lcurve_corner" > $O/r25_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r25_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $O/r25_bench.json 2> $O/r25_bench.err
WHICH=2x,4 timeout 600 python tools/gpu_configs.py > $O/r25_configs.log 2>&1
tail -n 3 $O/r25_pytest.log; cut -c1-150 $O/r25_configs.log
