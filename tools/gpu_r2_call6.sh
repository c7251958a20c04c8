#!/bin/bash
# round 2, GPU call 6: Lanczos ground state + QPU plugin tests, balanced lean passes A/B, low-bit floor A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c6_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c6_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu"
for lb in 5 4 3; do
  for bal in 1 0; do
    VQE_LOW_BITS=$lb VQE_EXP_BALANCE=$bal timeout 400 $B > gpurun_out/r2c6_bench_lb${lb}_bal${bal}.json 2> gpurun_out/r2c6_bench_lb${lb}_bal${bal}.err
  done
done
tail -3 gpurun_out/r2c6_pytest.log
