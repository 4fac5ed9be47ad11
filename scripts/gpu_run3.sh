#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/pytest_net.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_net.log
timeout 300 python scripts/conv2_probe.py one 2 0 > gpurun_out/conv2_probe_cg2.log 2>&1
timeout 300 python scripts/net_microbench.py > gpurun_out/net_microbench.log 2>&1
timeout 900 python bench.py --steps 800 --warmup 8 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
