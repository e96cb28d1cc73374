#!/bin/bash
# Round-2 GPU call #26 (final kernels: per-kernel launch bounds, unrolled inner products in the X2 / T2SPARC kernels): full records of HEAD — whole GPU suite, bench (both arms), launch list, full-size counters and
# ncu --set full of the config-2 kernels (rank-16 echo space), of the config-4 BayesReg kernel and the L-curve kernel
# (t2_echo_reg_kernel), every other config, final A/B of the L-curve / BayesReg kernel families.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T=r26
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 600 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_launches.log 2>&1
export SHAPE=96,96,60
timeout 200 python tools/prof_one.py > $O/${T}_plain_prof.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum \
    --clock-control none -k regex:'fa_search|fa_select|t2_echo|spline_weights|reduce_partials' -c 8 --csv --log-file $O/${T}_counters.csv \
    python tools/prof_one.py > $O/${T}_ncu_counters.log 2>&1
export SHAPE=96,96,6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_echo|fa_search_thread|fa_select_kernel' -c 3 \
    -o $O/${T}_prof python tools/prof_one.py > $O/${T}_ncu_prof.log 2>&1
export SHAPE=96,96,2
METHOD=BayesReg RM=InvT2 FA=brute-force NTE=48 TAU=8.0 NPC=100 timeout 200 python tools/prof_one.py > $O/${T}_plain_c4.log 2>&1 &&
METHOD=BayesReg RM=InvT2 FA=brute-force NTE=48 TAU=8.0 NPC=100 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'t2_echo_reg' -c 1 -o $O/${T}_prof_c4 python tools/prof_one.py > $O/${T}_ncu_c4.log 2>&1
METHOD=L_curve timeout 200 python tools/prof_one.py > $O/${T}_plain_lc.log 2>&1 &&
METHOD=L_curve timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'t2_echo_reg' -c 1 -o $O/${T}_prof_lc python tools/prof_one.py > $O/${T}_ncu_lc.log 2>&1
unset SHAPE
timeout 900 python tools/gpu_configs.py > $O/${T}_configs.log 2>&1
cp $O/configs.json $O/${T}_configs.json
timeout 800 python tools/gpu_ab_echo_reg.py > $O/${T}_ab_echo_reg.log 2>&1
timeout 400 python tools/gpu_ab_echo_rank.py > $O/${T}_ab_echo_rank.log 2>&1
cp $O/ab_echo_reg.json $O/${T}_ab_echo_reg.json
du -sh $O; ls -la $O | tail -24
