#!/bin/bash
# round 2, GPU call 10: full test suite (LiH, sharded exponential, qubit relabelling on virtual ranks), bench, ncu of the ADAPT kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2c10_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c10_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_apply_lean -s 20 -c 1 -o gpurun_out/r2c10_apply python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep > gpurun_out/r2c10_ncu_apply.log 2>&1
timeout 600 $NCU -k regex:k_tile_pool -s 20 -c 1 -o gpurun_out/r2c10_pool python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep > gpurun_out/r2c10_ncu_pool.log 2>&1
timeout 600 $NCU -k regex:k_tile_col -s 300 -c 1 -o gpurun_out/r2c10_col_heavy python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c10_ncu_col.log 2>&1
tail -5 gpurun_out/r2c10_pytest.log
