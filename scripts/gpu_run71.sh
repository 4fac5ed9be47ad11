#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-exchange --no-cpu-baseline --no-configs > gpurun_out/r2_run71_bench.json 2> gpurun_out/r2_run71_bench.err
