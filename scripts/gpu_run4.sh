#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 50 --csv \
  --log-file gpurun_out/launches2.csv python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rz_(conv3x3_tc2|heads|gomoku_encode)' \
  -s 2400 -c 24 -o gpurun_out/wave2_full python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
