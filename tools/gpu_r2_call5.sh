#!/bin/bash
# round 2, GPU call 5: 128-byte-swizzled TMA tiles, prefetched entries, lite run descriptors -- tests, A/B benches, ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2c5_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c5_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-pool --no-cpu"
timeout 300 $B > gpurun_out/r2c5_bench_swz.json 2> gpurun_out/r2c5_bench_swz.err
VQE_SWIZZLE=0 timeout 300 $B > gpurun_out/r2c5_bench_noswz.json 2> gpurun_out/r2c5_bench_noswz.err
VQE_PIPE=1 timeout 300 $B > gpurun_out/r2c5_bench_pipe.json 2> gpurun_out/r2c5_bench_pipe.err
VQE_LOW_BITS=4 timeout 300 $B > gpurun_out/r2c5_bench_lb4.json 2> gpurun_out/r2c5_bench_lb4.err
VQE_LOW_BITS=3 timeout 300 $B > gpurun_out/r2c5_bench_lb3.json 2> gpurun_out/r2c5_bench_lb3.err
VQE_PIPE=1 VQE_LOW_BITS=3 timeout 300 $B > gpurun_out/r2c5_bench_lb3_pipe.json 2> gpurun_out/r2c5_bench_lb3_pipe.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/r2c5_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c5_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_tile_col -s 147 -c 1 -o gpurun_out/r2c5_col24 python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c5_ncu_col.log 2>&1
timeout 600 $NCU -k regex:k_expect_lean -s 127 -c 1 -o gpurun_out/r2c5_exp_heavy python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c5_ncu_exp.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-sweep > gpurun_out/r2c5_bench_full.json 2> gpurun_out/r2c5_bench_full.err
tail -3 gpurun_out/r2c5_pytest.log
