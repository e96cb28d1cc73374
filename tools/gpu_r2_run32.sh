#!/bin/bash
# Round-2 GPU call #32: final check of HEAD — smoke(), the whole GPU suite, the default bench line and the reference arm.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/r32_smoke.log 2>&1; echo "rc=$?" >> $O/r32_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r32_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r32_pytest.log
timeout 600 python bench.py > $O/r32_bench.json 2> $O/r32_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r32_bench_ref.json 2> $O/r32_bench_ref.err
tail -n 2 $O/r32_smoke.log; tail -n 3 $O/r32_pytest.log; cut -c1-400 $O/r32_bench.json; cut -c1-200 $O/r32_bench_ref.json
