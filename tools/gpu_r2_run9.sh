#!/bin/bash
# Round-2 GPU call #9: FA thread kernel with warp-synchronised phases.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "fa_search or config2_subset or config1 or methods_subset or golden_vectors" > $O/r9_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r9_pytest.log
MET2_FA_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > $O/r9_bench.json 2> $O/r9_bench.err
timeout 200 SHAPE=96,96,60 python tools/prof_one.py > $O/r9_plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fa_search_thread' -c 1 \
    -o $O/r9_prof_fa env SHAPE=96,96,60 python tools/prof_one.py > $O/r9_ncu_prof.log 2>&1
MET2_FA_DEBUG=1 WHICH=1,4,3a timeout 900 python tools/gpu_configs.py > $O/r9_configs.log 2>&1
ls -la $O | tail -8
