#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -x -q -k "one_launch_trunk" 2>&1 | tail -15 > gpurun_out/r2_run36_trunk_small.log
timeout 600 python -m pytest tests/test_gpu_net.py -x -q 2>&1 | tail -8 > gpurun_out/r2_run36_net.log
timeout 300 python scripts/small_batch_probe.py > gpurun_out/r2_run36_small_batch.log 2>&1
