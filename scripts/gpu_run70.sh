#!/bin/bash
# state of the tree near the end of round 2: GPU suite, smoke, the bench line (both arms), launch list of the same command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_run70_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_run70_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_run70_bench.json 2> gpurun_out/r2_run70_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_run70_bench_reference.json 2> gpurun_out/r2_run70_bench_reference.err
