#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_run63_pytest_gpu.log
timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run63_single_game.log 2>&1
timeout 600 python scripts/bench_configs.py 1 > gpurun_out/r2_run63_config1.log 2>&1
