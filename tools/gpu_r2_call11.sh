#!/bin/bash
# round 2, GPU call 11 (2 GPUs): C5 at 34 qubits with qubit relabelling
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712"
timeout 1200 $T bench.py --gpus 2 --workload c5 --steps 2 --warmup 1 --verify > gpurun_out/r2c11_c5_n34_g2_relabel.json 2> gpurun_out/r2c11_c5_n34_g2_relabel.err
tail -c 1500 gpurun_out/r2c11_c5_n34_g2_relabel.json; tail -5 gpurun_out/r2c11_c5_n34_g2_relabel.err
