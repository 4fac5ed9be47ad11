#!/bin/bash
# Full GPU suite + bench arms + config 5 (MuZero) + the launch list of the headline step.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_run17_pytest_gpu.log 2>&1
tail -3 gpurun_out/r1_run17_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_run17_smoke.log 2>&1
tail -1 gpurun_out/r1_run17_smoke.log
timeout 600 python scripts/bench_configs.py 5 > gpurun_out/r1_run17_bench_config5_muzero.log 2>&1
tail -2 gpurun_out/r1_run17_bench_config5_muzero.log
timeout 600 python bench.py > gpurun_out/r1_run17_bench.json 2> gpurun_out/r1_run17_bench.err
tail -c 2500 gpurun_out/r1_run17_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 72 --csv \
  --log-file gpurun_out/r1_run17_wave_launches.csv python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
