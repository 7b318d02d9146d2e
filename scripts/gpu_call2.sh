#!/bin/bash
# round-2 call 2: GPU tests after the config refactor + screening-policy sweep in the driver's window and in steady state
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu.log
tail -6 gpurun_out/r02_pytest_gpu.log
run() { # name, window args, solver opts
  timeout 300 python bench.py $2 --no-cpu-baseline --no-e2e --solver-opts "$3" > gpurun_out/r02_pol_$1.json 2> gpurun_out/r02_pol_$1.err
  echo "## $1 [$2] [$3]"; python scripts/show_bench.py gpurun_out/r02_pol_$1.json 2>/dev/null | head -2 || tail -3 gpurun_out/r02_pol_$1.err
}
D="--steps 20 --warmup 5"; S="--steps 60 --warmup 40"
run d_s3 "$D" ""
run d_s5 "$D" "screen_rounds=5"
run d_s8 "$D" "screen_rounds=8"
run d_s8b3 "$D" "screen_rounds=8,bail_round=3"
run d_s12 "$D" "screen_rounds=12"
run d_s5w4 "$D" "screen_rounds=5,dense_warm_rounds=4"
run d_v1 "$D" "solver_variant=1"
run s_s3 "$S" ""
run s_s5 "$S" "screen_rounds=5"
run s_s8 "$S" "screen_rounds=8"
run s_s12 "$S" "screen_rounds=12"
