#!/bin/bash
# Round-2 GPU call #29: A/B — Gram-domain grid points of the L-curve kernel cold-started (variant libmet2_cold.so).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
for v in "" cold; do
  export MET2_LIB_VARIANT=$v; [ -z "$v" ] && unset MET2_LIB_VARIANT
  WHICH=config2_L_curve_I,config2_L_curve_InvT2 timeout 500 python tools/gpu_ab_echo_reg.py > $O/r29_ab_${v:-base}.log 2>&1
done
grep -h -o '"config[^"]*": {"voxels": [0-9]*, "echo_rank": [0-9]*, "t2_ms_echo": [0-9.]*\|"lambda_differs": [0-9]*' $O/r29_ab_base.log $O/r29_ab_cold.log
