#!/bin/bash
# Round-2 GPU call #7: thread-per-voxel FA search, G out of shared memory for the wide grids, warp spline weights.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r7_pytest.log
MET2_FA_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > $O/r7_bench.json 2> $O/r7_bench.err
MET2_FA_SEARCH=warp timeout 300 python bench.py --no-cpu-baseline > $O/r7_bench_warp.json 2> $O/r7_bench_warp.err
MET2_FA_DEBUG=1 WHICH=1,4,3a,3b,2x timeout 900 python tools/gpu_configs.py > $O/r7_configs.log 2>&1
cp $O/configs.json $O/r7_configs.json
timeout 200 python tools/prof_one.py > $O/r7_plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fa_search_thread' -c 1 \
    -o $O/r7_prof_fa python tools/prof_one.py > $O/r7_ncu_prof.log 2>&1
ls -la $O | tail -12
