#!/bin/bash
# round 2, GPU call 20: launch list (time + DRAM bytes of every launch of one evaluation) for the 13-bit real-layout tiles
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
   --log-file gpurun_out/r2c20_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2c20_ncu.log 2>&1
tail -2 gpurun_out/r2c20_ncu.log
