#!/bin/bash
# round 2, GPU call 23: ncu --set full of a 25-run and a 28-run pass of k_col_stab; butterfly form of the expectation pair sum
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-pool"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2c23_bench_${name}.json 2> gpurun_out/r2c23_bench_${name}.err; }
run new VQE_X=0
run old VQE_COL_TAB=0 VQE_EXP_RL2=0
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_boundary_gpu.py -m gpu -x -q > gpurun_out/r2c23_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c23_pytest.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_col_stab -s 12 -c 2 -o gpurun_out/r2c23_stab python bench.py --steps 1 --warmup 1 --no-cpu --no-pool --no-sweep > gpurun_out/r2c23_ncu_stab.log 2>&1
tail -3 gpurun_out/r2c23_pytest.log
for f in gpurun_out/r2c23_bench_*.json; do python - $f <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; o=d['roofline_other']
    print(sys.argv[1], 'ms',round(d['ms_per_step'],2),'E',d['energy_first_step'],'rot',r['launches_per_step'],round(r['avg_launch_us'],1),'exp',o['launches_per_step'],round(o['avg_launch_us'],1))
except Exception as e: print(sys.argv[1], 'FAILED', e)
P
done
