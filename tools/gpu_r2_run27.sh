#!/bin/bash
# Round-2 GPU call #27: A/B of a compile-time switch (full-block ldl_8x8 + unrolled M_P update in the X2 / T2SPARC kernels):
# bench + config-2 method times + echo parity subset.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --no-cpu-baseline > $O/r27_bench.json 2> $O/r27_bench.err
WHICH=2x timeout 600 python tools/gpu_configs.py > $O/r27_configs.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "echo or config2_subset" > $O/r27_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r27_pytest.log
tail -n 3 $O/r27_pytest.log; cut -c1-150 $O/r27_configs.log
