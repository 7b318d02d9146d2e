#!/bin/bash
# round-2 call 3: new dense kernel (chunked condensing, TMA tiles) + tile-ring variants of the Riccati kernel
mkdir -p gpurun_out
C=$PWD/mpc_quad_ros_b200/csrc
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r03_pytest_default.log; tail -3 gpurun_out/r03_pytest_default.log
for v in r2 r2p; do
  QMPC_LIB=$C/libqmpc_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fp64_vs_oracle or variants or closed_loop or grouped" 2>&1 | tail -5 > gpurun_out/r03_pytest_$v.log; tail -2 gpurun_out/r03_pytest_$v.log
done
QMPC_LIB=$C/libqmpc_prof.so B=512 STEPS=8 timeout 300 python scripts/profile_step.py 2>&1 | grep "dense ocp" | tail -12 > gpurun_out/r03_dense_phases.txt; cat gpurun_out/r03_dense_phases.txt
run() { # name lib window opts
  if [ "$2" != "default" ]; then export QMPC_LIB=$C/libqmpc_$2.so; else unset QMPC_LIB; fi
  timeout 300 python bench.py $3 --no-cpu-baseline --no-e2e --solver-opts "$4" > gpurun_out/r03_$1.json 2> gpurun_out/r03_$1.err
  echo "## $1 [$2] [$3] [$4]"; python scripts/show_bench.py gpurun_out/r03_$1.json 2>/dev/null | head -2 || tail -3 gpurun_out/r03_$1.err
  unset QMPC_LIB
}
D="--steps 20 --warmup 5"; S="--steps 60 --warmup 40"
for lib in default p r2 r2p r3p; do
  run d_${lib}_s5 $lib "$D" "screen_rounds=5"
  run s_${lib}_s5 $lib "$S" "screen_rounds=5"
done
run d_default_s3 default "$D" ""
run d_default_s8b3 default "$D" "screen_rounds=8,bail_round=3"
run d_r2_s8b3 r2 "$D" "screen_rounds=8,bail_round=3"
run s_default_s3 default "$S" ""
B=67 STEPS=3 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python scripts/profile_step.py > gpurun_out/r03_racecheck_default.txt 2>&1; tail -3 gpurun_out/r03_racecheck_default.txt
QMPC_LIB=$C/libqmpc_r2.so B=67 STEPS=3 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python scripts/profile_step.py > gpurun_out/r03_racecheck_r2.txt 2>&1; tail -3 gpurun_out/r03_racecheck_r2.txt
QMPC_LIB=$C/libqmpc_r2.so B=67 STEPS=3 timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python scripts/profile_step.py > gpurun_out/r03_memcheck_r2.txt 2>&1; tail -3 gpurun_out/r03_memcheck_r2.txt
