#!/bin/bash
# single-game latency through AlphaZeroMCTS.simulate after the small-batch work + ncu --set full of the new kernels
mkdir -p gpurun_out
timeout 300 python scripts/single_game_latency.py > gpurun_out/r2_run44_single_game.log 2>&1
RZ_EAGER=1 RZ_G=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rz_(trunk_small|heads_tc4)" -s 4 -c 2 \
  -o gpurun_out/r2_run44_small_full python scripts/small_batch_probe.py > gpurun_out/r2_run44_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_run44_small_full.ncu-rep gpurun_out/r2_run44_small_ncu_full_summary.csv >> gpurun_out/r2_run44_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rz_heads_tc4" -s 100 -c 1 \
  -o gpurun_out/r2_run44_heads_full python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e --no-configs --no-exchange > gpurun_out/r2_run44_ncu2.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_run44_heads_full.ncu-rep gpurun_out/r2_run44_heads8192_ncu_full_summary.csv >> gpurun_out/r2_run44_ncu2.log 2>&1
