#!/bin/bash
# programmatic dependent launch A/B: tests with PDL on, single-game latency and the bench line with RZ_PDL=1 / 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_run34_pytest_gpu.log
for pdl in 1 0; do
  RZ_PDL=$pdl timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run34_single_game_pdl$pdl.log 2>&1
  RZ_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-exchange > gpurun_out/r2_run34_bench_pdl$pdl.json 2> gpurun_out/r2_run34_bench_pdl$pdl.err
done
