#!/bin/bash
# round 2, GPU call 9 (2 GPUs): sharded tests on two real GPUs, C5 at 34 qubits with and without the gather prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q --durations=4 > gpurun_out/r2c9_pytest_sharded.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c9_pytest_sharded.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
VQE_GATHER_OVERLAP=1 timeout 900 $T bench.py --gpus 2 --workload c5 --steps 2 --warmup 1 --verify > gpurun_out/r2c9_c5_n34_g2_overlap.json 2> gpurun_out/r2c9_c5_n34_g2_overlap.err
VQE_GATHER_OVERLAP=0 timeout 900 $T bench.py --gpus 2 --workload c5 --steps 2 --warmup 1 > gpurun_out/r2c9_c5_n34_g2_serial.json 2> gpurun_out/r2c9_c5_n34_g2_serial.err
tail -2 gpurun_out/r2c9_pytest_sharded.log; tail -c 600 gpurun_out/r2c9_c5_n34_g2_overlap.json
