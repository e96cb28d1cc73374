#!/bin/bash
# Round-2 GPU call #24: unrolled triangular products of the echo kernels (echo_tmul / echo_tmul_t) — parity subset, bench, times.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "echo or methods_subset or config4_subset or config2_subset or lcurve_corner" > $O/r24_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r24_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $O/r24_bench.json 2> $O/r24_bench.err
WHICH=2x,4 timeout 600 python tools/gpu_configs.py > $O/r24_configs.log 2>&1
tail -n 3 $O/r24_pytest.log; cut -c1-150 $O/r24_configs.log
