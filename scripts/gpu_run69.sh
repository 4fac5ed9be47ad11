#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_selfplay.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4 > gpurun_out/r2_run69_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-exchange --no-cpu-baseline > gpurun_out/r2_run69_bench.json 2> gpurun_out/r2_run69_bench.err
