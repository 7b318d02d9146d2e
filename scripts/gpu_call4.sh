#!/bin/bash
# round-2 call 4: dense work queue + tile ring default; policy A/B with repeats; ncu source profile of both solver kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r04_pytest.log; tail -3 gpurun_out/r04_pytest.log
timeout 600 python scripts/tune_policy.py 3 "" "screen_rounds=5" "screen_rounds=5,bail_round=3" "screen_rounds=8,bail_round=3" "screen_rounds=6,bail_round=3,dense_warm_rounds=4" "screen_rounds=8,bail_round=3,dense_warm_rounds=4" "screen_rounds=4" 2>&1 | tee gpurun_out/r04_policy.txt
GROUPS=16 timeout 300 python scripts/tune_policy.py 2 "screen_rounds=5" "screen_rounds=8,bail_round=3" 2>&1 | tee gpurun_out/r04_policy_g16.txt
GROUPS=4 timeout 300 python scripts/tune_policy.py 2 "screen_rounds=5" "screen_rounds=8,bail_round=3" 2>&1 | tee gpurun_out/r04_policy_g4.txt
timeout 300 python scripts/diag_transient.py 30 > gpurun_out/r04_transient.txt 2>&1; tail -8 gpurun_out/r04_transient.txt
STEPS=14 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qmpc_ipm|qmpc_dense" -s 24 -c 2 -f -o gpurun_out/r04_prof_solver python scripts/profile_step.py > gpurun_out/r04_ncu.log 2>&1; tail -2 gpurun_out/r04_ncu.log
