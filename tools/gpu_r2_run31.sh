#!/bin/bash
# Round-2 GPU call #31: X2 kernel at 22 warps / 88 registers (variant libmet2_t704.so) against the product (20 / 96).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
for v in "" t704; do
  export MET2_LIB_VARIANT=$v; [ -z "$v" ] && unset MET2_LIB_VARIANT
  METHOD=X2 timeout 200 python tools/gpu_time.py > $O/r31_x2_${v:-base}.log 2>&1
  echo "${v:-base}: $(grep "^{'fa_ms'" $O/r31_x2_${v:-base}.log | tail -n 1)"
  METHOD=L_curve timeout 200 python tools/gpu_time.py > $O/r31_lc_${v:-base}.log 2>&1
  echo "${v:-base} L_curve: $(grep "^{'fa_ms'" $O/r31_lc_${v:-base}.log | tail -n 1)"
done
