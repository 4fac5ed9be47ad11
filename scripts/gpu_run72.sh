#!/bin/bash
# ncu --set full of the dominant kernels of the headline wave on the final tree (round-2 evidence for roofline.traffic)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rz_(conv3x3_tc2|stem_tc|heads_tc4|expand_backup|select)" -s 240 -c 8 \
  -o gpurun_out/r2_run72_wave_full python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e --no-configs --no-exchange > gpurun_out/r2_run72_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_run72_wave_full.ncu-rep gpurun_out/r2_run72_wave_ncu_full_summary.csv >> gpurun_out/r2_run72_ncu.log 2>&1
