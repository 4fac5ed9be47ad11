#!/bin/bash
mkdir -p gpurun_out
for m in 74 1000; do RZ_GS=100,128,192,256,296,384,512 RZ_SMALL_BATCH_MAX=$m timeout 300 python scripts/small_batch_probe.py >> gpurun_out/r2_run43_crossover.log 2>&1; done
