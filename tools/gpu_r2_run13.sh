#!/bin/bash
# Round-2 GPU call #13: A/B of the reduced echo rank (24 = product, 20, 16): parity on the reference fixtures + bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
for v in r16 r20; do
  export MET2_LIB_VARIANT=$v MET2_ECHO_TAIL_MAX=1e-11
  timeout 900 python -m pytest tests -m gpu -q -k "echo_space or config2_subset or methods_subset or golden_vectors or t2sparc or full_size or montecarlo" > $O/r13_pytest_$v.log 2>&1; echo "pytest rc=$?" >> $O/r13_pytest_$v.log
  cp $O/parity_r2_echo_space.json $O/r13_parity_echo_$v.json; cp $O/parity_config2_subset.json $O/r13_parity_config2_$v.json
  timeout 300 python bench.py --no-cpu-baseline > $O/r13_bench_$v.json 2> $O/r13_bench_$v.err
  METHOD=T2SPARC RM=InvT2 timeout 300 python tools/gpu_check_echo.py > $O/r13_echo_t2sparc_$v.log 2>&1
done
ls -la $O | tail -8
