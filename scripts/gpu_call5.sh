#!/bin/bash
# round-2 call 5: adaptive (busy-step) screening policy sweep + new GPU tests (device reference generators, cold handle, ...)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r05_pytest.log; tail -6 gpurun_out/r05_pytest.log
timeout 900 python scripts/tune_policy.py 3 "" "screen_rounds_busy=-1" "screen_busy_pct=12" "screen_busy_pct=16" "screen_busy_pct=25" "screen_busy_pct=35" "screen_busy_pct=16,screen_rounds=4" "screen_busy_pct=16,screen_rounds_busy=6" "screen_busy_pct=16,screen_rounds_busy=12" 2>&1 | tee gpurun_out/r05_policy.txt
