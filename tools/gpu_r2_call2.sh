#!/bin/bash
# round 2, GPU call 2: full GPU test-suite on the new kernels, A/B benches, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c2_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-pool --no-cpu"
VQE_COL_KERNEL=0 VQE_EXP_LEAN=0 timeout 300 $B > gpurun_out/r2c2_bench_old.json 2> gpurun_out/r2c2_bench_old.err
timeout 300 $B > gpurun_out/r2c2_bench_new.json 2> gpurun_out/r2c2_bench_new.err
VQE_COL_KERNEL=0 timeout 300 $B > gpurun_out/r2c2_bench_newexp_oldrot.json 2> gpurun_out/r2c2_bench_a.err
VQE_LOW_BITS=4 timeout 300 $B > gpurun_out/r2c2_bench_lb4.json 2> gpurun_out/r2c2_bench_lb4.err
VQE_LOW_BITS=3 timeout 300 $B > gpurun_out/r2c2_bench_lb3.json 2> gpurun_out/r2c2_bench_lb3.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c2_bench_full.json 2> gpurun_out/r2c2_bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2c2_bench_ref.json 2> gpurun_out/r2c2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/r2c2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool > gpurun_out/r2c2_ncu.log 2>&1
tail -3 gpurun_out/r2c2_pytest.log
