#!/bin/bash
# Round-2 measurement pass on one B200: tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of
# the three hot kernels in the start-up transient, sanitizer runs, timelines.  Everything lands in gpurun_out/; then
# `python scripts/make_profiles_r02.py` (on the CPU box) turns it into profiles/r02_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_N1.json 2> gpurun_out/r02_bench_N1.err; python scripts/show_bench.py gpurun_out/r02_bench_N1.json
timeout 900 python bench.py --steps 100 --warmup 40 --no-cpu-baseline --no-extra-legs > gpurun_out/r02_bench_N1_steady.json 2> gpurun_out/r02_bench_N1_steady.err; python scripts/show_bench.py gpurun_out/r02_bench_N1_steady.json | head -2
# launch list of the bench command itself (serialised, cold caches: compare SHARES)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extra-legs --latency-steps 20 > gpurun_out/r02_ncu_launches.log 2>&1; tail -1 gpurun_out/r02_ncu_launches.log | cut -c1-200
# full captures: step 12 (busy transient step) and step 60 (steady)
STEPS=13 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qmpc_ipm|qmpc_dense|qmpc_linearize" -s 36 -c 3 -f -o gpurun_out/r02_prof_step12 python scripts/profile_step.py > gpurun_out/r02_ncu_step12.log 2>&1; tail -1 gpurun_out/r02_ncu_step12.log
STEPS=61 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:qmpc_ipm|qmpc_dense|qmpc_linearize" -s 180 -c 3 -f -o gpurun_out/r02_prof_step60 python scripts/profile_step.py > gpurun_out/r02_ncu_step60.log 2>&1; tail -1 gpurun_out/r02_ncu_step60.log
# sanitizer on an odd batch: fp64 default path (K1, screening, dense, RGP, plant), fp32 path, shared-swarm kernels
B=67 STEPS=4 timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python scripts/profile_step.py > gpurun_out/r02_memcheck_fp64.txt 2>&1; tail -2 gpurun_out/r02_memcheck_fp64.txt
B=67 STEPS=3 timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python scripts/profile_step.py > gpurun_out/r02_racecheck_fp64.txt 2>&1; tail -2 gpurun_out/r02_racecheck_fp64.txt
# racecheck does not model the mbarrier completion flags of the column-per-warp Cholesky (profiles/r02_sanitizer.txt): the same library with
# the factorisation synchronised by CTA barriers (build it first: bash scripts/build_variants.sh f0="-DQMPC_DENSE_FACTOR=0") must be clean
F0=$PWD/mpc_quad_ros_b200/csrc/libqmpc_f0.so
[ -f $F0 ] && QMPC_LIB=$F0 B=67 STEPS=3 timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python scripts/profile_step.py > gpurun_out/r02_racecheck_fp64_barrier_build.txt 2>&1; tail -2 gpurun_out/r02_racecheck_fp64_barrier_build.txt
B=67 STEPS=3 PREC=32 timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python scripts/profile_step.py > gpurun_out/r02_racecheck_fp32.txt 2>&1; tail -2 gpurun_out/r02_racecheck_fp32.txt
B=67 STEPS=3 PREC=32 timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python scripts/profile_step.py > gpurun_out/r02_memcheck_fp32.txt 2>&1; tail -2 gpurun_out/r02_memcheck_fp32.txt
[ -f $F0 ] && export QMPC_LIB=$F0
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "shared_swarm_rgp_vs_sequential or device_reference or rgp_learn" > gpurun_out/r02_racecheck_rgp.txt 2>&1; tail -3 gpurun_out/r02_racecheck_rgp.txt
unset QMPC_LIB
# per-OCP timelines: busy transient step and steady state
timeout 300 python scripts/diag_timeline.py 8 > gpurun_out/r02_timeline_step8.txt 2>&1
timeout 300 python scripts/diag_timeline.py 60 > gpurun_out/r02_timeline_step60.txt 2>&1
EVERY=2 timeout 300 python scripts/diag_transient.py 40 "" > gpurun_out/r02_transient.txt 2>&1
