#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r08_pytest.log; tail -3 gpurun_out/r08_pytest.log
timeout 900 python scripts/tune_policy.py 3 "" "screen_rounds_busy=-1" "screen_busy_pct=25" "screen_busy_pct=25,screen_rounds=4" "screen_busy_pct=30,screen_rounds=4,screen_rounds_busy=6" 2>&1 | tee gpurun_out/r08_policy.txt
EVERY=4 timeout 300 python scripts/diag_transient.py 100 "" > gpurun_out/r08_trans.txt 2>&1; tail -26 gpurun_out/r08_trans.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r08_bench.json 2> gpurun_out/r08_bench.err; tail -c 3000 gpurun_out/r08_bench.json; tail -5 gpurun_out/r08_bench.err
