#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_api.py tests/test_gpu_tree.py tests/test_rollout.py tests/test_gpu_train.py -m gpu -x -q > gpurun_out/r1_run43_pytest.log 2>&1
tail -5 gpurun_out/r1_run43_pytest.log | cut -c1-300
timeout 300 python scripts/bench_configs.py stock15 1 > gpurun_out/r1_run43_stock.log 2>&1; cat gpurun_out/r1_run43_stock.log | cut -c1-330
timeout 300 python scripts/single_game_latency.py > gpurun_out/r1_run43_single_game.log 2>&1; cat gpurun_out/r1_run43_single_game.log | cut -c1-200
