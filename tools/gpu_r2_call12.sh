#!/bin/bash
# round 2, GPU call 12 (2 GPUs): sharded tests (relabelling + real layout on shards), C5 at 34 qubits
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q --durations=4 > gpurun_out/r2c12_pytest_sharded.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c12_pytest_sharded.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713"
timeout 1200 $T bench.py --gpus 2 --workload c5 --steps 2 --warmup 1 --verify > gpurun_out/r2c12_c5_n34_g2.json 2> gpurun_out/r2c12_c5_n34_g2.err
tail -3 gpurun_out/r2c12_pytest_sharded.log; tail -c 1200 gpurun_out/r2c12_c5_n34_g2.json; tail -5 gpurun_out/r2c12_c5_n34_g2.err
