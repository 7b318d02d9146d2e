#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests -m gpu -q -s --tb=short -k "two_ranks or free_running or odometry" 2>&1 | grep -v "^  \|^$" | tail -20 > gpurun_out/r12_pytest.log; cat gpurun_out/r12_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r12_bench_N2.json 2> gpurun_out/r12_bench_N2.err
tail -c 2500 gpurun_out/r12_bench_N2.json; tail -8 gpurun_out/r12_bench_N2.err
