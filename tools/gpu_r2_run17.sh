#!/bin/bash
# Round-2 GPU call #17: hybrid L-curve (grid points below 1e-3 in the Gram domain, the rest in echo space) — corner
# agreement over the whole volume, arbitration by the oracle, parity tests, A/B times.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "lcurve or L_curve or methods_subset or config4 or golden_vectors or warm_start" > $O/r17_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r17_pytest.log
NMAX=40 timeout 600 python tools/gpu_lcurve_arbiter.py > $O/r17_lcurve_arbiter.log 2>&1
WHICH=config2_L_curve_I,config2_L_curve_InvT2 timeout 600 python tools/gpu_ab_echo_reg.py > $O/r17_ab_echo_reg.log 2>&1
cp $O/ab_echo_reg.json $O/r17_ab_echo_reg.json
WHICH=5a,2x timeout 600 python tools/gpu_configs.py > $O/r17_configs.log 2>&1
cp $O/configs.json $O/r17_configs.json
ls -la $O | tail -6
