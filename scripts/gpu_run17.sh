#!/bin/bash
# Leaf-parallel (virtual loss) mode: parity against the oracle wave, the untouched K = 1 path, single-game throughput.
set -x
mkdir -p gpurun_out
python -m oracle.build_oracle
timeout 600 python -m pytest tests/test_gpu_leaf_parallel.py -m gpu -q > gpurun_out/r1_run35_pytest_leaf_parallel.log 2>&1
tail -30 gpurun_out/r1_run35_pytest_leaf_parallel.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_api.py tests/test_gpu_selfplay.py tests/test_gpu_fullsize.py tests/test_gpu_dm.py tests/test_gpu_go_search.py -m gpu -x -q > gpurun_out/r1_run35_pytest_tree.log 2>&1
tail -4 gpurun_out/r1_run35_pytest_tree.log | cut -c1-250
timeout 600 python scripts/single_game_latency.py > gpurun_out/r1_run35_single_game.log 2>&1
cat gpurun_out/r1_run35_single_game.log | cut -c1-250
