#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_api.py tests/test_gpu_selfplay.py -x -q 2>&1 | tail -10 > gpurun_out/r2_run62_tests.log
timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run62_single_game.log 2>&1
