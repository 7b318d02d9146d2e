#!/bin/bash
# kernel-tuning sweep on the GPU box: same bench, different builds / options; one JSON line each in gpurun_out/tune.log
mkdir -p gpurun_out
: > gpurun_out/tune.log
run() { echo "## $*" >> gpurun_out/tune.log; "$@" 2>> gpurun_out/tune.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print(json.dumps({k:d[k] for k in ('value','ms_per_step')}|{k:r[k] for k in ('ms_per_launch','ms_linearize_per_launch','n_ipm_mean','n_refine_rounds_mean','warm_start_success_frac','frac')}|{'p99':d['latency_ms']['p99'],'bad':d['solver']['status_not_ok_last_step']}))
" >> gpurun_out/tune.log; }
B="python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-e2e"
run $B
run $B --cold
for lib in $TUNE_LIBS; do
  QMPC_LIB=$PWD/mpc_quad_ros_b200/csrc/$lib run $B
done
run $B --batch 16384
run $B --workload lemniscate
cat gpurun_out/tune.log
