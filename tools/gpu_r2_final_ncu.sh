#!/bin/bash
# round 2, final profiles: launch lists (time + DRAM bytes of every launch of one evaluation; ncu's default flushes the caches
# before every kernel = cold, --cache-control none = what the kernels see in the stream) and --set full captures of a light
# rotation pass, a 25-run rotation pass, an expectation pass and the diagonal kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_boundary_gpu.py tests/test_sharded_gpu.py -m gpu -q > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
( time timeout 1500 python bench.py ) > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv"
timeout 600 ncu $M --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2g_ncu.log 2>&1
timeout 600 ncu $M --cache-control none --log-file gpurun_out/r2g_launches_warm.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2g_ncu_warm.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_col_stab -s 10 -c 3 -o gpurun_out/r2g_col python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2g_ncu_col.log 2>&1
timeout 600 $NCU -k regex:k_expect_rlp -s 20 -c 1 -o gpurun_out/r2g_exp python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2g_ncu_exp.log 2>&1
timeout 600 $NCU -k regex:k_expect_diag2 -s 1 -c 1 -o gpurun_out/r2g_diag python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2g_ncu_diag.log 2>&1
tail -3 gpurun_out/r2g_pytest.log; tail -c 600 gpurun_out/r2g_bench_n1.json; ls -la gpurun_out/r2g_*
