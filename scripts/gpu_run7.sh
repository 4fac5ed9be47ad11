#!/bin/bash
# Re-validation of HEAD on a fresh box: GPU parity suite, smoke, both bench arms.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_run15_pytest_gpu.log 2>&1
tail -3 gpurun_out/r1_run15_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_run15_smoke.log 2>&1
tail -2 gpurun_out/r1_run15_smoke.log
timeout 600 python bench.py > gpurun_out/r1_run15_bench.json 2> gpurun_out/r1_run15_bench.err
tail -c 3000 gpurun_out/r1_run15_bench.json
timeout 300 python bench.py --impl reference --cpu-seconds 10 > gpurun_out/r1_run15_bench_reference.json 2>&1
tail -c 1500 gpurun_out/r1_run15_bench_reference.json
