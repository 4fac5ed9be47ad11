#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/r2_run68_smi.log
timeout 300 python scripts/small_batch_probe.py > gpurun_out/r2_run68_small_batch.log 2>&1
timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run68_single_game.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-exchange --no-cpu-baseline > gpurun_out/r2_run68_bench.json 2> gpurun_out/r2_run68_bench.err
