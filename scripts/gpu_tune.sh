#!/bin/bash
# kernel-tuning sweep on the GPU box: same bench, different builds / options; one JSON line each in gpurun_out/tune.log
mkdir -p gpurun_out
: > gpurun_out/tune.log
run() { echo "## $QMPC_LIB $*" >> gpurun_out/tune.log; "$@" 2>> gpurun_out/tune.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print(json.dumps({k:d[k] for k in ('value','ms_per_step')}|{k:r[k] for k in ('ms_per_launch','ms_dense_per_launch','ms_linearize_per_launch','n_ipm_mean','n_refine_rounds_mean','warm_start_success_frac','frac')}|{'p99':d['latency_ms']['p99'],'bad':d['solver']['status_not_ok_last_step']}))
" >> gpurun_out/tune.log; }
B="python bench.py --steps ${STEPS:-40} --warmup ${WARMUP:-10} --no-cpu-baseline --no-e2e"
if [ "${TUNE_BASE:-1}" = "1" ]; then run $B; fi
for lib in $TUNE_LIBS; do
  export QMPC_LIB=$PWD/mpc_quad_ros_b200/csrc/$lib
  run $B
  unset QMPC_LIB
done
IFS=';' read -ra EX <<< "$TUNE_EXTRA"
for extra in "${EX[@]}"; do [ -n "$extra" ] && run $B $extra; done
for pad in $TUNE_PADS; do export QMPC_IPM_SMEM_PAD=$pad; echo "# pad $pad" >> gpurun_out/tune.log; run $B $TUNE_PAD_ARGS; unset QMPC_IPM_SMEM_PAD; done
cat gpurun_out/tune.log
if [ "${NCU:-0}" = "1" ]; then
STEPS=14 timeout 900 ncu --set full --clock-control none --import-source on -k regex:qmpc_ipm -s 12 -c 1 -f -o gpurun_out/prof_ipm python scripts/profile_step.py > gpurun_out/ncu_ipm.log 2>&1
tail -3 gpurun_out/ncu_ipm.log
fi
