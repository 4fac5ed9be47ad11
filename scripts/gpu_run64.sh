#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_tree.py tests/test_gpu_dm.py tests/test_gpu_round2.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -6 > gpurun_out/r2_run64_tests.log
timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run64_single_game.log 2>&1
