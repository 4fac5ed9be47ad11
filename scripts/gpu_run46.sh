#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -x -q -k "one_launch_trunk or batch_invariant or search_with_native" 2>&1 | tail -12 > gpurun_out/r2_run46_trunk_pairs.log
timeout 300 python scripts/small_kernel_probe.py > gpurun_out/r2_run46_small_kernels.log 2>&1
timeout 300 python scripts/small_batch_probe.py > gpurun_out/r2_run46_small_batch.log 2>&1
