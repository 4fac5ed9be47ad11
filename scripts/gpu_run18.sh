#!/bin/bash
# fp32 conv split over blocks (single-game path): stock-net parity tests + single-game throughput again.
set -x
mkdir -p gpurun_out
python -m oracle.build_oracle
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_api.py tests/test_gpu_leaf_parallel.py tests/test_gpu_train.py tests/test_rollout.py -m gpu -x -q > gpurun_out/r1_run36_pytest.log 2>&1
tail -4 gpurun_out/r1_run36_pytest.log | cut -c1-250
timeout 600 python scripts/single_game_latency.py > gpurun_out/r1_run36_single_game.log 2>&1
cat gpurun_out/r1_run36_single_game.log | cut -c1-250
timeout 300 python scripts/bench_configs.py 1 > gpurun_out/r1_run36_bench_config1.log 2>&1
cat gpurun_out/r1_run36_bench_config1.log | cut -c1-300
