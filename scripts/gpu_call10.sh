#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r10_pytest.log; tail -3 gpurun_out/r10_pytest.log
timeout 900 python scripts/tune_policy.py 3 "" "screen_rounds_busy=-1" "screen_busy_pct=25" 2>&1 | tee gpurun_out/r10_policy.txt
EVERY=4 timeout 300 python scripts/diag_transient.py 100 "" > gpurun_out/r10_trans.txt 2>&1; tail -26 gpurun_out/r10_trans.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; python scripts/show_bench.py gpurun_out/r10_bench.json; tail -5 gpurun_out/r10_bench.err
