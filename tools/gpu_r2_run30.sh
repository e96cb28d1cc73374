#!/bin/bash
# Round-2 GPU call #30: X2 kernel with fewer resident warps (MET2_T2_WARPS 14 / 16 / 18; product = 20) at the final code size.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
for w in 14 16 18 20; do
  MET2_T2_WARPS=$w METHOD=X2 timeout 200 python tools/gpu_time.py > $O/r30_x2_warps_$w.log 2>&1
  echo "warps $w: $(grep "^{'fa_ms'" $O/r30_x2_warps_$w.log | tail -n 1)"
done
