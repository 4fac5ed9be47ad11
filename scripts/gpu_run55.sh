#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_selfplay.py -x -q 2>&1 | tail -6 > gpurun_out/r2_run55_tests.log
timeout 300 python scripts/small_kernel_probe.py 2>&1 | grep -v probe_ns > gpurun_out/r2_run55_small_kernels.log
timeout 300 python scripts/small_batch_probe.py > gpurun_out/r2_run55_small_batch.log 2>&1
timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run55_single_game.log 2>&1
timeout 600 python scripts/wave_timeline.py 2>&1 | grep -E "games|heads" > gpurun_out/r2_run55_wave_timeline_8192.log
