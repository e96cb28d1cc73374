#!/bin/bash
# Round-2 GPU call #18: t2_echo_reg_kernel as a state machine with one Gram-domain and one echo-space call site
# (11 k instead of 18.8 k SASS instructions for the L-curve): parity tests, corner agreement, A/B times; L-curve switch sweep.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "lcurve or L_curve or BayesReg or methods_subset or config4 or golden_vectors or warm_start" > $O/r18_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r18_pytest.log
timeout 800 python tools/gpu_ab_echo_reg.py > $O/r18_ab_echo_reg.log 2>&1
cp $O/ab_echo_reg.json $O/r18_ab_echo_reg.json
for sw in 0 1e-6 1e-5 1e-4; do
  MET2_LCURVE_SWITCH=$sw WHICH=config2_L_curve_I,config2_L_curve_InvT2 timeout 300 python tools/gpu_ab_echo_reg.py > $O/r18_ab_lcurve_switch_$sw.log 2>&1
done
WHICH=4,5a timeout 600 python tools/gpu_configs.py > $O/r18_configs.log 2>&1
cp $O/configs.json $O/r18_configs.json
ls -la $O | tail -6
