#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_run58_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_run58_smoke.log 2>&1
timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run58_single_game.log 2>&1
RZ_FUSE_SELECT=0 timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run58_single_game_unfused.log 2>&1
