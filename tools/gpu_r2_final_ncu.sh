#!/bin/bash
# round 2, final profiles: launch list (time + DRAM bytes of every launch of one evaluation) and --set full captures of a light
# rotation pass, a 25-run rotation pass and an expectation pass
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2f_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_col_stab -s 10 -c 3 -o gpurun_out/r2f_col python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2f_ncu_col.log 2>&1
timeout 600 $NCU -k regex:k_expect_rlp -s 20 -c 1 -o gpurun_out/r2f_exp python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2f_ncu_exp.log 2>&1
timeout 600 $NCU -k regex:k_expect_diag2 -s 1 -c 1 -o gpurun_out/r2f_diag python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2f_ncu_diag.log 2>&1
ls -la gpurun_out/r2f_*
