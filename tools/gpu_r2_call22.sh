#!/bin/bash
# round 2, GPU call 22: shared-memory item table (k_col_stab) -- small-register tests, A/B against k_tile_col, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_boundary_gpu.py tests/test_sharded_gpu.py tests/test_widen_gpu.py -m gpu -x -q --durations=3 > gpurun_out/r2c22_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c22_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-pool"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2c22_bench_${name}.json 2> gpurun_out/r2c22_bench_${name}.err; }
run stab VQE_X=0
run old VQE_COL_TAB=0
run stab_h12 VQE_BENCH_MOLECULE=h12
run old_h12 VQE_BENCH_MOLECULE=h12 VQE_COL_TAB=0
run skeleton VQE_DEBUG_SKELETON=1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r2c22_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2c22_ncu.log 2>&1
tail -3 gpurun_out/r2c22_pytest.log
for f in gpurun_out/r2c22_bench_*.json; do python - $f <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; o=d['roofline_other']
    print(sys.argv[1], 'ms',round(d['ms_per_step'],2),'E',d['energy_first_step'],'rot',r['launches_per_step'],round(r['avg_launch_us'],1),'exp',o['launches_per_step'],round(o['avg_launch_us'],1))
except Exception as e: print(sys.argv[1], 'FAILED', e)
P
done
