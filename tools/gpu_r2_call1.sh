#!/bin/bash
# round 2, GPU call 1: 24-qubit parity of the round-1 kernels + full ncu captures (source level) of the two hot kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt 2>&1
nproc >> gpurun_out/r2c1_smi.txt
timeout 900 python -m pytest tests/test_parity_24q_gpu.py -x -q --durations=0 > gpurun_out/r2c1_parity24.log 2>&1
echo "parity exit $?" >> gpurun_out/r2c1_parity24.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_tile_rot -s 147 -c 1 -o gpurun_out/r2c1_rot24 python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c1_ncu_rot.log 2>&1
timeout 600 $NCU -k regex:k_tile_expect -s 137 -c 1 -o gpurun_out/r2c1_exp_mixed python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c1_ncu_exp1.log 2>&1
timeout 600 $NCU -k regex:k_tile_expect -s 148 -c 1 -o gpurun_out/r2c1_exp_flat python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c1_ncu_exp2.log 2>&1
ls -la gpurun_out/*.ncu-rep
