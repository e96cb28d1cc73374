#!/bin/bash
# Round-2 GPU call #11: FA thread kernel (contiguous per-thread state, one solve site, compact loops), full-size ncu.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "fa_search or config2_subset or config1 or methods_subset or golden_vectors or plain_nnls_wide or config4" > $O/r11_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r11_pytest.log
MET2_FA_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > $O/r11_bench.json 2> $O/r11_bench.err
MET2_FA_DEBUG=1 WHICH=1,4,3a,3b timeout 900 python tools/gpu_configs.py > $O/r11_configs.log 2>&1
export SHAPE=96,96,60
timeout 200 python tools/prof_one.py > $O/r11_plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fa_search_thread' -c 1 \
    -o $O/r11_prof_fa python tools/prof_one.py > $O/r11_ncu_prof.log 2>&1
ls -la $O | tail -4
