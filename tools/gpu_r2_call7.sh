#!/bin/bash
# round 2, GPU call 7: real layout of purely real states (8-byte tiles), gather prefetch, H2O workload -- tests, A/B benches, ncu
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c7_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c7_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu"
for mol in h2o h12; do
  for rl in 1 0; do
    for lb in 4 5; do
      VQE_REAL_LAYOUT=$rl VQE_LOW_BITS=$lb timeout 400 $B --molecule $mol > gpurun_out/r2c7_bench_${mol}_rl${rl}_lb${lb}.json 2> gpurun_out/r2c7_bench_${mol}_rl${rl}_lb${lb}.err
    done
  done
done
VQE_EXP_LEAN_THREADS=384 timeout 400 $B --molecule h12 > gpurun_out/r2c7_bench_h12_rl1_t384.json 2> gpurun_out/r2c7_bench_h12_rl1_t384.err
VQE_EXP_LEAN_THREADS=512 timeout 400 $B --molecule h12 > gpurun_out/r2c7_bench_h12_rl1_t512.json 2> gpurun_out/r2c7_bench_h12_rl1_t512.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_tile_col -s 147 -c 1 -o gpurun_out/r2c7_col24 python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --molecule h12 > gpurun_out/r2c7_ncu_col.log 2>&1
timeout 600 $NCU -k regex:k_expect_lean -s 100 -c 1 -o gpurun_out/r2c7_exp python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --molecule h12 > gpurun_out/r2c7_ncu_exp.log 2>&1
tail -3 gpurun_out/r2c7_pytest.log
