#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_run42_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-exchange > gpurun_out/r2_run42_bench.json 2> gpurun_out/r2_run42_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rz_(select|stem|conv|heads|expand)" -s 2400 -c 72 --csv \
  --log-file gpurun_out/r2_run42_wave_launches.csv python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e --no-configs --no-exchange > gpurun_out/r2_run42_ncu.log 2>&1
