#!/bin/bash
# Round-2 GPU call #5: ldl lane code, position-slot loops (PS), 16 warps/128 regs vs 20 warps/96 regs; Monte-Carlo; multi tests.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r5_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r5_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $O/r5_bench.json 2> $O/r5_bench.err
MET2_LIB_VARIANT=w16 timeout 300 python bench.py --no-cpu-baseline > $O/r5_bench_w16.json 2> $O/r5_bench_w16.err
MET2_T2_WARPS=18 timeout 300 python bench.py --no-cpu-baseline --steps 3 > $O/r5_bench_w18.json 2>/dev/null
MET2_LIB_VARIANT=w16 METHOD=T2SPARC RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r5_echo_t2sparc_w16.log 2>&1
METHOD=T2SPARC RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r5_echo_t2sparc.log 2>&1
timeout 300 python tools/montecarlo.py 10000 0 > $O/r5_montecarlo.log 2>&1
ls -la $O | tail -12
