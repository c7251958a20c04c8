#!/bin/bash
# the driver's own command at N = 2 (replicas + the sharded C5 leg)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
tail -c 600 gpurun_out/r2g_bench_n2.json; tail -3 gpurun_out/r2g_bench_n2.err
