#!/bin/bash
# round 2, GPU call 3: tile-ring kernels -- tests, A/B benches, skeleton timing, ncu source-level captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2c3_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c3_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-pool --no-cpu"
timeout 300 $B > gpurun_out/r2c3_bench_pipe.json 2> gpurun_out/r2c3_bench_pipe.err
VQE_PIPE=0 timeout 300 $B > gpurun_out/r2c3_bench_nopipe.json 2> gpurun_out/r2c3_bench_nopipe.err
VQE_LOW_BITS=4 timeout 300 $B > gpurun_out/r2c3_bench_lb4.json 2> gpurun_out/r2c3_bench_lb4.err
VQE_LOW_BITS=3 timeout 300 $B > gpurun_out/r2c3_bench_lb3.json 2> gpurun_out/r2c3_bench_lb3.err
VQE_EXP_THREADS=256 timeout 300 $B > gpurun_out/r2c3_bench_t256.json 2> gpurun_out/r2c3_bench_t256.err
VQE_DEBUG_SKELETON=1 timeout 300 $B > gpurun_out/r2c3_bench_skel.json 2> gpurun_out/r2c3_bench_skel.err
VQE_DEBUG_SKELETON=1 VQE_LOW_BITS=3 timeout 300 $B > gpurun_out/r2c3_bench_skel_lb3.json 2> gpurun_out/r2c3_bench_skel_lb3.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/r2c3_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c3_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_col_pipe -s 147 -c 1 -o gpurun_out/r2c3_col24 python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c3_ncu_col.log 2>&1
timeout 600 $NCU -k regex:k_expect_pipe -s 127 -c 1 -o gpurun_out/r2c3_exp_heavy python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c3_ncu_exp.log 2>&1
tail -3 gpurun_out/r2c3_pytest.log
