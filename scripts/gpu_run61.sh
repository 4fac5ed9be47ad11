#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -x -q -k "one_launch or small_and_large" 2>&1 | tail -12 > gpurun_out/r2_run61_tests.log
