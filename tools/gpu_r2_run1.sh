#!/bin/bash
# Round-2 GPU call #1: full GPU suite (incl. the new wide-grid / config-4 / echo-space tests), bench A/Bs, fresh ncu.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/r1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r1_pytest.log
timeout 600 python bench.py > $O/r1_bench.json 2> $O/r1_bench.err; echo "bench rc=$?"
MET2_LIB_VARIANT=norescue timeout 300 python bench.py --no-cpu-baseline > $O/r1_bench_norescue.json 2> $O/r1_bench_norescue.err
timeout 300 python bench.py --no-cpu-baseline --t2-flags 64 > $O/r1_bench_echo.json 2> $O/r1_bench_echo.err
METHOD=T2SPARC RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r1_echo_t2sparc.log 2>&1
RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r1_echo_x2_invt2.log 2>&1
WHICH=4,2x timeout 600 python tools/gpu_configs.py > $O/r1_configs.log 2>&1
# ncu: launch list of the bench command, then full captures of the three hot kernels on a 55 296-voxel slab
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r1_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r1_ncu_launches.log 2>&1
timeout 200 python tools/prof_one.py > $O/r1_plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_fit_kernel|fa_search_kernel|fa_select_kernel' -c 3 \
    -o $O/r1_prof python tools/prof_one.py > $O/r1_ncu_prof.log 2>&1
ls -la $O
