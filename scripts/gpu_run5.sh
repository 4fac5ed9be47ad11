#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/net_microbench.py > gpurun_out/net_microbench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rz_(heads|stem)' -c 6 \
  -o gpurun_out/heads_full python scripts/net_microbench.py > gpurun_out/ncu_heads.log 2>&1
timeout 900 python bench.py --steps 800 --warmup 8 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
