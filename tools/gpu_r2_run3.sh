#!/bin/bash
# Round-2 GPU call #3: reduced-echo-space kernels as default; A/B vs Gram; rescue variants; ncu of the new kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r3_pytest.log
timeout 600 python bench.py > $O/r3_bench.json 2> $O/r3_bench.err
timeout 300 python bench.py --no-cpu-baseline --gram > $O/r3_bench_gram.json 2> $O/r3_bench_gram.err
MET2_LIB_VARIANT=norescue timeout 300 python bench.py --no-cpu-baseline > $O/r3_bench_norescue.json 2> $O/r3_bench_norescue.err
MET2_LIB_VARIANT=inl timeout 300 python bench.py --no-cpu-baseline > $O/r3_bench_inl.json 2> $O/r3_bench_inl.err
for w in 8 12 14; do MET2_T2_WARPS=$w timeout 300 python bench.py --no-cpu-baseline --steps 3 > $O/r3_bench_w$w.json 2>/dev/null; done
METHOD=T2SPARC RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r3_echo_t2sparc.log 2>&1
RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r3_echo_x2_invt2.log 2>&1
timeout 200 python tools/prof_one.py > $O/r3_plain_echo.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_echo' -c 1 \
    -o $O/r3_prof_echo_x2 python tools/prof_one.py > $O/r3_ncu_echo_x2.log 2>&1
ls -la $O | tail -30
