#!/bin/bash
# Round-2 GPU call #28: A/B — unrolled factor update also in the L-curve / BayesReg kernel (variant libmet2_upd.so).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
for v in "" upd; do
  export MET2_LIB_VARIANT=$v; [ -z "$v" ] && unset MET2_LIB_VARIANT
  WHICH=config2_L_curve_I,config2_BayesReg_I,config2_BayesReg_InvT2,config4_BayesReg_InvT2 timeout 500 python tools/gpu_ab_echo_reg.py > $O/r28_ab_${v:-base}.log 2>&1
done
grep -h -o '"config[^"]*": {"voxels": [0-9]*, "echo_rank": [0-9]*, "t2_ms_echo": [0-9.]*' $O/r28_ab_base.log $O/r28_ab_upd.log
