#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r09_pytest.log; tail -3 gpurun_out/r09_pytest.log
timeout 900 python scripts/tune_policy.py 3 "" "screen_rounds_busy=-1" "screen_busy_pct=25" 2>&1 | tee gpurun_out/r09_policy.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r09_bench.json 2> gpurun_out/r09_bench.err; tail -c 4500 gpurun_out/r09_bench.json; tail -5 gpurun_out/r09_bench.err
