#!/bin/bash
# round 2, GPU call 15: sharded tests after the split expectation
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q --durations=4 > gpurun_out/r2c15_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c15_pytest.log
tail -4 gpurun_out/r2c15_pytest.log
