#!/bin/bash
# Round-2 GPU call #12: full validation of HEAD — whole GPU suite, bench (both arms), launch list, full-size counters and
# ncu --set full of the three config-2 kernels, every other config.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r12_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r12_pytest.log
timeout 600 python bench.py > $O/r12_bench.json 2> $O/r12_bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r12_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r12_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r12_ncu_launches.log 2>&1
export SHAPE=96,96,60
timeout 200 python tools/prof_one.py > $O/r12_plain_prof.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum \
    --clock-control none -k regex:'fa_search|fa_select|t2_echo|spline_weights|reduce_partials' -c 8 --csv --log-file $O/r12_counters.csv \
    python tools/prof_one.py > $O/r12_ncu_counters.log 2>&1
export SHAPE=96,96,6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_echo|fa_search_thread|fa_select_kernel' -c 3 \
    -o $O/r12_prof python tools/prof_one.py > $O/r12_ncu_prof.log 2>&1
unset SHAPE
WHICH=1,4,5a,5b,3a,3b,2x timeout 900 python tools/gpu_configs.py > $O/r12_configs.log 2>&1
cp $O/configs.json $O/r12_configs.json
ls -la $O | tail -14
