#!/bin/bash
# Round-2 GPU call #10: full-size ncu capture of the thread-per-voxel FA search.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export SHAPE=96,96,60
timeout 200 python tools/prof_one.py > $O/r10_plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fa_search_thread' -c 1 \
    -o $O/r10_prof_fa python tools/prof_one.py > $O/r10_ncu_prof.log 2>&1
ls -la $O | tail -4
