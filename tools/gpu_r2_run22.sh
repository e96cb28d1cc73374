#!/bin/bash
# Round-2 GPU call #22: A/B of the echo kernels' launch bound (MET2_ECHO_MAX_THREADS 640 = product, 768, 896, 1024:
# 20 / 24 / 28 / 30 warps per SM at 96 / 80 / 72 / 64 registers), variants libmet2_t<N>.so.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
for v in "" t768 t896 t1024; do
  export MET2_LIB_VARIANT=$v; [ -z "$v" ] && unset MET2_LIB_VARIANT
  tag=${v:-t640}
  timeout 300 python bench.py --no-cpu-baseline > $O/r22_bench_$tag.json 2> $O/r22_bench_$tag.err
  WHICH=2x timeout 300 python tools/gpu_configs.py > $O/r22_configs_$tag.log 2>&1
done
export MET2_LIB_VARIANT=t768
timeout 600 python -m pytest tests -m gpu -q -k "echo_space or config2_subset or methods_subset or lcurve_corner" > $O/r22_pytest_t768.log 2>&1; echo "rc=$?" >> $O/r22_pytest_t768.log
export MET2_LIB_VARIANT=t1024
timeout 600 python -m pytest tests -m gpu -q -k "echo_space or config2_subset or methods_subset or lcurve_corner" > $O/r22_pytest_t1024.log 2>&1; echo "rc=$?" >> $O/r22_pytest_t1024.log
grep -h "L_curve\|X2 I\|BayesReg\|T2SPARC" $O/r22_configs_*.log | cut -c1-120
