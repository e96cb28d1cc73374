#!/bin/bash
# Round-2 GPU call #2: suite with the BayesReg rework, rescue by value A/B, ncu of the echo-space and BayesReg kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $O/r2_bench.json 2> $O/r2_bench.err
MET2_LIB_VARIANT=norescue timeout 300 python bench.py --no-cpu-baseline > $O/r2_bench_norescue.json 2> $O/r2_bench_norescue.err
WHICH=4 timeout 600 python tools/gpu_configs.py > $O/r2_configs.log 2>&1
cp $O/configs.json $O/r2_configs.json
export T2FLAGS=64
timeout 200 python tools/prof_one.py > $O/r2_plain_echo.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_echo' -c 1 \
    -o $O/r2_prof_echo_x2 python tools/prof_one.py > $O/r2_ncu_echo_x2.log 2>&1
export METHOD=T2SPARC RM=InvT2
timeout 200 python tools/prof_one.py > $O/r2_plain_echo_tik.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_echo' -c 1 \
    -o $O/r2_prof_echo_tik python tools/prof_one.py > $O/r2_ncu_echo_tik.log 2>&1
export T2FLAGS=0 METHOD=BayesReg RM=InvT2 NTE=48 TAU=8.0 NPC=100 FA=brute-force SHAPE=96,96,2
timeout 200 python tools/prof_one.py > $O/r2_plain_bayes.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_fit_kernel' -c 1 \
    -o $O/r2_prof_bayes4 python tools/prof_one.py > $O/r2_ncu_bayes4.log 2>&1
ls -la $O | tail -30
