#!/bin/bash
# round 2, GPU call 14 (8 GPUs): the driver's own command at N = 8 (replicas + sharded C5 at 36 qubits with verify at 30)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29714"
timeout 1500 $T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c14_bench_n8.json 2> gpurun_out/r2c14_bench_n8.err
tail -c 2500 gpurun_out/r2c14_bench_n8.json; tail -5 gpurun_out/r2c14_bench_n8.err
