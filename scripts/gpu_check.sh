#!/bin/bash
# Runs on the GPU box (via gpurun): tests, smoke, a short bench, optional ncu; logs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1200 python -m pytest tests -m gpu -q -s ${PYTEST_ARGS:-} 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -5 gpurun_out/smoke.log
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
timeout 900 python bench.py --steps ${STEPS:-50} --warmup ${WARMUP:-5} ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
if [ "${NCU:-0}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 90 --csv --log-file gpurun_out/launches.csv env STEPS=42 python scripts/profile_step.py > gpurun_out/ncu_launches.log 2>&1
STEPS=42 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qmpc_ipm|qmpc_dense" -s 80 -c 2 -f -o gpurun_out/prof_ipm python scripts/profile_step.py > gpurun_out/ncu_ipm.log 2>&1
tail -3 gpurun_out/ncu_ipm.log
fi
