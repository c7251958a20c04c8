#!/bin/bash
# round 2, GPU call 13: sharded + engine tests after the real-layout expectation on shards
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sharded_gpu.py tests/test_engine_gpu.py tests/test_boundary_gpu.py -m gpu -x -q --durations=4 > gpurun_out/r2c13_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c13_pytest.log
tail -4 gpurun_out/r2c13_pytest.log
