#!/bin/bash
# round 2, GPU call 21: item-table rotation kernel (k_col_tab), two-slot expectation kernel (k_expect_rl2), programmatic
# dependent launch -- tests (without the two slowest 24-qubit oracle runs) + A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=4 -k "not full_uccsd_energy and not sigma_and_pool" > gpurun_out/r2c21_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c21_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-pool"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2c21_bench_${name}.json 2> gpurun_out/r2c21_bench_${name}.err; }
run new VQE_X=0
run old VQE_COL_TAB=0 VQE_EXP_RL2=0
run nopdl VQE_PDL=0
run tab_only VQE_EXP_RL2=0
run skeleton VQE_DEBUG_SKELETON=1
run exp384 VQE_EXP_LEAN_THREADS=384
run expctas2 VQE_EXP_RL_CTAS=2
run h12 VQE_BENCH_MOLECULE=h12
tail -3 gpurun_out/r2c21_pytest.log
for f in gpurun_out/r2c21_bench_*.json; do python - $f <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; o=d['roofline_other']
    print(sys.argv[1], 'ms',round(d['ms_per_step'],2),'E',d['energy_first_step'],'rot',r['launches_per_step'],round(r['avg_launch_us'],1),'exp',o['launches_per_step'],round(o['avg_launch_us'],1))
except Exception as e: print(sys.argv[1], 'FAILED', e)
P
done
