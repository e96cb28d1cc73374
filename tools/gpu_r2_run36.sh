#!/bin/bash
# Round-2 GPU call #36: final check of HEAD — smoke(), the whole GPU suite, the default bench line and the reference arm.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/r36_smoke.log 2>&1; echo "rc=$?" >> $O/r36_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r36_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r36_pytest.log
timeout 600 python bench.py > $O/r36_bench.json 2> $O/r36_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r36_bench_ref.json 2> $O/r36_bench_ref.err
tail -n 2 $O/r36_smoke.log; tail -n 3 $O/r36_pytest.log; cut -c1-400 $O/r36_bench.json; cut -c1-200 $O/r36_bench_ref.json
