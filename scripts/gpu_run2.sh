#!/bin/bash
# GPU session 2: all parity tests with conv v2 as the default trunk, bench, ncu of conv v2.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 800 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rz_conv3x3_tc2' \
  -s 205 -c 3 -o gpurun_out/conv2_full python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
