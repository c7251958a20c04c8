#!/bin/bash
# round 2, GPU call 18: 13-bit real-layout tiles for rotation passes -- tests + A/B
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=4 > gpurun_out/r2c18_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c18_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --no-pool"
for tb in 13 12; do for mol in h2o h12; do
  VQE_RL_TILE_BITS=$tb timeout 400 $B --molecule $mol > gpurun_out/r2c18_bench_${mol}_tb${tb}.json 2> gpurun_out/r2c18_bench_${mol}_tb${tb}.err
done; done
tail -3 gpurun_out/r2c18_pytest.log
