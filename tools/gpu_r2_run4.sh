#!/bin/bash
# Round-2 GPU call #4: 20-warp reduced-echo kernel, balanced rank-1, dynamic FA scheduling; new bench.py; warp sweep.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r4_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r4_pytest.log
timeout 600 python bench.py > $O/r4_bench.json 2> $O/r4_bench.err
MET2_LIB_VARIANT=norescue timeout 300 python bench.py --no-cpu-baseline > $O/r4_bench_norescue.json 2> $O/r4_bench_norescue.err
for w in 16 18; do MET2_T2_WARPS=$w timeout 300 python bench.py --no-cpu-baseline --steps 3 > $O/r4_bench_w$w.json 2>/dev/null; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r4_bench_ref.json 2> $O/r4_bench_ref.err
METHOD=T2SPARC RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r4_echo_t2sparc.log 2>&1
RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r4_echo_x2_invt2.log 2>&1
WHICH=1,4,5a,5b,3a,3b,2x timeout 900 python tools/gpu_configs.py > $O/r4_configs.log 2>&1
cp $O/configs.json $O/r4_configs.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r4_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r4_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r4_ncu_launches.log 2>&1
timeout 200 python tools/prof_one.py > $O/r4_plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_echo|fa_search_kernel|fa_select_kernel' -c 3 \
    -o $O/r4_prof python tools/prof_one.py > $O/r4_ncu_prof.log 2>&1
ls -la $O | tail -30
