#!/bin/bash
# Runs on the GPU box (via gpurun): tests, smoke, a short bench; logs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps ${STEPS:-50} --warmup ${WARMUP:-5} > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
