#!/bin/bash
# Round-2 GPU call #21: smoke(), the file-level tests after the CLI changes (warm-up thread, concurrent output writers),
# wall time of the CLI on the config-2 volume.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/r21_smoke.log 2>&1; echo "rc=$?" >> $O/r21_smoke.log
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_abi.py -m gpu -q > $O/r21_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r21_pytest.log
timeout 500 python tools/gpu_cli_time.py > $O/r21_cli_time.log 2>&1
tail -3 $O/r21_smoke.log; tail -3 $O/r21_pytest.log; tail -14 $O/r21_cli_time.log
