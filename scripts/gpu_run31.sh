#!/bin/bash
# Validation of the state at the end of round 1: full GPU suite, smoke, both bench arms, every configuration, launch list.
set -x
mkdir -p gpurun_out
python -m oracle.build_oracle
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_run50_pytest_gpu.log 2>&1
tail -3 gpurun_out/r1_run50_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_run50_smoke.log 2>&1
tail -1 gpurun_out/r1_run50_smoke.log
timeout 600 python bench.py > gpurun_out/r1_run50_bench.json 2> gpurun_out/r1_run50_bench.err
tail -c 700 gpurun_out/r1_run50_bench.json
timeout 300 python bench.py --impl reference --cpu-seconds 10 > gpurun_out/r1_run50_bench_reference.json 2>&1
timeout 900 python scripts/bench_configs.py 1 2 4 5 stock15 > gpurun_out/r1_run50_bench_configs.log 2>&1
cut -c1-260 gpurun_out/r1_run50_bench_configs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 72 --csv \
  --log-file gpurun_out/r1_run50_wave_launches.csv python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
