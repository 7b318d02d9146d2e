#!/bin/bash
# round-2 call 1: baseline numbers in the driver's window, transient diagnostics, sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python scripts/diag_transient.py 32 > gpurun_out/r02_transient.txt 2>&1
tail -34 gpurun_out/r02_transient.txt
for v in 2 0; do
  QMPC_IPM_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02_base_v$v.json 2> gpurun_out/r02_base_v$v.err
  python scripts/show_bench.py gpurun_out/r02_base_v$v.json 2>/dev/null || tail -c 600 gpurun_out/r02_base_v$v.json
done
for g in 4 16; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --groups $g > gpurun_out/r02_base_g$g.json 2> gpurun_out/r02_base_g$g.err
  python scripts/show_bench.py gpurun_out/r02_base_g$g.json 2>/dev/null
done
# sanitizer: odd batch, a few closed-loop steps (cold first step -> dense IPM for all, then screening + dense)
B=67 STEPS=4 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/profile_step.py > gpurun_out/r02_memcheck.txt 2>&1
tail -5 gpurun_out/r02_memcheck.txt
B=67 STEPS=3 timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python scripts/profile_step.py > gpurun_out/r02_racecheck.txt 2>&1
tail -5 gpurun_out/r02_racecheck.txt
