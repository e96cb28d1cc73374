#!/bin/bash
# diagnostic: T2 stage time as a function of warps per CTA (MET2_T2_WARPS only lowers the limit)
for w in 4 6 8 10; do
  echo -n "warps=$w "
  MET2_T2_WARPS=$w REPS=2 timeout 200 python tools/gpu_time.py 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["rep1"])'
done
