#!/bin/bash
# round 2, last call: smoke + boundary tests + one quick bench on the final build
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
timeout 60 python -m pytest tests/test_boundary_gpu.py tests/test_engine_gpu.py -m gpu -q -x > gpurun_out/r2z_pytest.log 2>&1
timeout 60 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-pool > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
cat gpurun_out/r2z_smoke.log; tail -2 gpurun_out/r2z_pytest.log; head -c 200 gpurun_out/r2z_bench.json
