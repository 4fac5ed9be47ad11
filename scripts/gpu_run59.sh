#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/soak_small.py > gpurun_out/r2_run59_soak_small.log 2>&1
echo "exit $?" >> gpurun_out/r2_run59_soak_small.log
