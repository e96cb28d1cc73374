#!/bin/bash
# Round-2 GPU call #15: L-curve / BayesReg in the reduced echo space (t2_echo_reg_kernel) — parity tests of both kernel
# families against the reference fixtures / oracle, full-volume A/B with times (config 2 methods, config 4).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -k "methods_subset or config4 or regularised_fit_vs_oracle or golden_vectors or warm_start or full_size or dropin or montecarlo" > $O/r15_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r15_pytest.log
timeout 800 python tools/gpu_ab_echo_reg.py > $O/r15_ab_echo_reg.log 2>&1
WHICH=4,5a timeout 600 python tools/gpu_configs.py > $O/r15_configs.log 2>&1
cp $O/configs.json $O/r15_configs.json
ls -la $O | tail -6
