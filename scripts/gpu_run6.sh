#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 48 --csv \
  --log-file gpurun_out/launches3.csv python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rz_(heads|stem|expand|select)|Lb1' \
  -s 20 -c 10 -o gpurun_out/wave3_full python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
