#!/bin/bash
# round 2, final record: full GPU test suite, the driver's bench command (timed), the reference arm, smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
( time timeout 1500 python bench.py ) > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
( time timeout 1500 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
tail -4 gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_bench_n1.err; tail -4 gpurun_out/r2f_bench_ref.err; cat gpurun_out/r2f_smoke.log
