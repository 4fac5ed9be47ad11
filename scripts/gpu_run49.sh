#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_go.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6 > gpurun_out/r2_run49_tests.log
timeout 600 python scripts/wave_timeline.py > gpurun_out/r2_run49_wave_timeline_8192.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-exchange --no-configs --no-cpu-baseline > gpurun_out/r2_run49_bench.json 2> gpurun_out/r2_run49_bench.err
