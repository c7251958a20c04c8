#!/bin/bash
# round 2, GPU call 8: new tests (LiH, sharded exponential), RL expectation occupancy A/B, launch list, full bench record
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2c8_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c8_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-pool"
for ct in 3 4 5 6; do
  VQE_EXP_RL_CTAS=$ct timeout 400 $B > gpurun_out/r2c8_bench_h2o_ctas${ct}.json 2> gpurun_out/r2c8_bench_h2o_ctas${ct}.err
done
VQE_LOW_BITS=5 timeout 400 $B > gpurun_out/r2c8_bench_h2o_lb5.json 2> gpurun_out/r2c8_bench_h2o_lb5.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/r2c8_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c8_ncu.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c8_bench_full.json 2> gpurun_out/r2c8_bench_full.err
tail -3 gpurun_out/r2c8_pytest.log
