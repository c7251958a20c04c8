#!/bin/bash
# round 2, GPU call 33: real-layout twin of the expectation entries with a lane table and conflict-free lane order -- tests + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_boundary_gpu.py tests/test_sharded_gpu.py tests/test_widen_gpu.py -m gpu -q > gpurun_out/r2c33_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c33_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-pool"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2c33_bench_${name}.json 2> gpurun_out/r2c33_bench_${name}.err; }
run lane VQE_X=0
run nolane VQE_EXP_LANE_TAB=0
run lane_h12 VQE_BENCH_MOLECULE=h12
run nolane_h12 VQE_BENCH_MOLECULE=h12 VQE_EXP_LANE_TAB=0
tail -3 gpurun_out/r2c33_pytest.log
for f in gpurun_out/r2c33_bench_*.json; do python - $f <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; o=d['roofline_other']
    print(sys.argv[1], 'ms',round(d['ms_per_step'],2),'E',d['energy_first_step'],'rot',r['launches_per_step'],round(r['avg_launch_us'],1),'exp',o['launches_per_step'],round(o['avg_launch_us'],1), 'launches', d['gpu_launches'])
except Exception as e: print(sys.argv[1], 'FAILED', e)
P
done
