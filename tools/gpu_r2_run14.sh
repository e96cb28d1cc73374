#!/bin/bash
# Round-2 GPU call #14: rank 16 / 24 selection of the reduced echo space by the measured residual — the echo tests, the
# full-volume A/B of both ranks against the Gram-domain kernels, bench of the new default.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "echo or config2_subset or methods_subset or golden_vectors or t2sparc or full_size or montecarlo or dropin" > $O/r14_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r14_pytest.log
timeout 400 python tools/gpu_ab_echo_rank.py > $O/r14_ab_echo_rank.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $O/r14_bench.json 2> $O/r14_bench.err
ls -la $O | tail -6
