#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -x -q -k "one_launch or small_and_large" 2>&1 | tail -4 > gpurun_out/r2_run73_tests.log
timeout 300 python scripts/small_kernel_probe.py > gpurun_out/r2_run73_small_kernels.log 2>&1
