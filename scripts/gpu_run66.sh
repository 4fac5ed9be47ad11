#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_run66_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-exchange --no-cpu-baseline > gpurun_out/r2_run66_bench.json 2> gpurun_out/r2_run66_bench.err
