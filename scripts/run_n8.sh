mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_N8.json 2> gpurun_out/r02_bench_N8.err
tail -c 1800 gpurun_out/r02_bench_N8.json; tail -3 gpurun_out/r02_bench_N8.err
