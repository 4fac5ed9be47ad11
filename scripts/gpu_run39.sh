#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_go.py -x -q 2>&1 | tail -8 > gpurun_out/r2_run39_net.log
timeout 300 python scripts/small_kernel_probe.py > gpurun_out/r2_run39_small_kernels.log 2>&1
timeout 300 python scripts/small_batch_probe.py > gpurun_out/r2_run39_small_batch.log 2>&1
