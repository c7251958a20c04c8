#!/bin/bash
# round 2, final record (2): 24-qubit parity tests on the final code, the driver's bench command, launch lists and the ncu capture
# of the expectation kernel with the lane-table entries
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_24q_gpu.py -m gpu -q --durations=5 > gpurun_out/r2h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
( time timeout 1500 python bench.py ) > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv"
timeout 600 ncu $M --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2h_ncu.log 2>&1
timeout 600 ncu $M --cache-control none --log-file gpurun_out/r2h_launches_warm.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2h_ncu_warm.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_expect_rlp -s 20 -c 1 -o gpurun_out/r2h_exp python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2h_ncu_exp.log 2>&1
tail -8 gpurun_out/r2h_pytest.log; tail -c 300 gpurun_out/r2h_bench_n1.json
