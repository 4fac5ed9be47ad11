#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/small_kernel_probe.py > gpurun_out/r2_run40_small_kernels.log 2>&1
