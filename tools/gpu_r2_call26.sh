#!/bin/bash
# round 2, GPU call 26: factored segment tables (96 bytes per segment: every pass keeps 3 CTAs per SM)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_boundary_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r2c26_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c26_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-pool"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2c26_bench_${name}.json 2> gpurun_out/r2c26_bench_${name}.err; }
run new VQE_X=0
run t256 VQE_STAB_THREADS2=256 VQE_STAB_THREADS3=256
run new_h12 VQE_BENCH_MOLECULE=h12
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r2c26_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2c26_ncu.log 2>&1
tail -3 gpurun_out/r2c26_pytest.log
for f in gpurun_out/r2c26_bench_*.json; do python - $f <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; o=d['roofline_other']
    print(sys.argv[1], 'ms',round(d['ms_per_step'],2),'E',d['energy_first_step'],'rot',r['launches_per_step'],round(r['avg_launch_us'],1),'exp',o['launches_per_step'],round(o['avg_launch_us'],1))
except Exception as e: print(sys.argv[1], 'FAILED', e)
P
done
