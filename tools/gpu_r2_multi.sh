#!/bin/bash
# Round-2 multi-GPU call: N = $1 GPUs.  Byte-identity of the gathered / multi-device results, then the timings.
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/m${N}_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/m${N}_pytest_multi.log 2>&1; echo "rc=$?" >> $O/m${N}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    tools/gpu_multi_check.py > $O/m${N}_multi_check.log 2>&1; echo "rc=$?" >> $O/m${N}_multi_check.log
MET2_MULTI_TRACE=1 timeout 600 python tools/multi_time.py > $O/m${N}_multi_time.log 2>&1; cp $O/multi_time.json $O/m${N}_multi_time.json
TILE=2 REPS=3 METHOD=L_curve OUT=m${N}_multi_time_c5_lcurve.json timeout 600 python tools/multi_time.py > $O/m${N}_multi_time_c5_lcurve.log 2>&1
TILE=2 REPS=3 METHOD=T2SPARC RM=InvT2 OUT=m${N}_multi_time_c5_t2sparc.json timeout 600 python tools/multi_time.py > $O/m${N}_multi_time_c5_t2sparc.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/m${N}_bench.json 2> $O/m${N}_bench.err; echo "bench rc=$?" >> $O/m${N}_bench.err
tail -3 $O/m${N}_bench.err; tail -2 $O/m${N}_multi_time.log
