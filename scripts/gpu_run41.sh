#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_go.py -x -q 2>&1 | tail -12 > gpurun_out/r2_run41_net.log
timeout 300 python scripts/small_kernel_probe.py 2>&1 | grep -v probe_ns > gpurun_out/r2_run41_small_kernels.log
timeout 300 python scripts/small_batch_probe.py > gpurun_out/r2_run41_small_batch.log 2>&1
