#!/bin/bash
# Round-2 GPU call #6: HEAD 2db7f11 — full GPU suite with the rank-pinned GCV test, bench both arms, launch list,
# ncu --set full of the three config-2 kernels, and of the slow configs (GCV-I, config-4 BayesReg + brute-force FA).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r6_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r6_pytest.log
timeout 600 python bench.py > $O/r6_bench.json 2> $O/r6_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r6_bench_ref.json 2> $O/r6_bench_ref.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r6_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r6_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r6_ncu_launches.log 2>&1
timeout 200 python tools/prof_one.py > $O/r6_plain_prof.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_echo|fa_search_kernel|fa_select_kernel' -c 3 \
    -o $O/r6_prof python tools/prof_one.py > $O/r6_ncu_prof.log 2>&1
METHOD=GCV SHAPE=96,96,2 timeout 200 python tools/prof_one.py > $O/r6_plain_gcv.log 2>&1 &&
METHOD=GCV SHAPE=96,96,2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'t2_fit_kernel' -c 1 \
    -o $O/r6_prof_gcv python tools/prof_one.py > $O/r6_ncu_gcv.log 2>&1
METHOD=BayesReg RM=InvT2 FA=brute-force NTE=48 TAU=8.0 NPC=100 SHAPE=96,96,2 timeout 200 python tools/prof_one.py > $O/r6_plain_c4.log 2>&1 &&
METHOD=BayesReg RM=InvT2 FA=brute-force NTE=48 TAU=8.0 NPC=100 SHAPE=96,96,2 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'t2_fit_kernel|fa_search_kernel|fa_select_kernel' -c 3 -o $O/r6_prof_c4 python tools/prof_one.py > $O/r6_ncu_c4.log 2>&1
ls -la $O | tail -20
