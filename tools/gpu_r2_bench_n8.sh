#!/bin/bash
# the driver's own command at N = 8 (replicas + the sharded C5 leg)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29728 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err
tail -c 300 gpurun_out/r2g_bench_n8.json; tail -3 gpurun_out/r2g_bench_n8.err
