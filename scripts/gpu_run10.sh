#!/bin/bash
# ncu --set full of the kernels added for Go / DeepMindMCTS / MuZero (one capture each, small -c).
set -x
mkdir -p gpurun_out
# Go at full size: tree kernels with the rules inside, the fused observation + stem, conv rev. 3 at stride 20
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'rz_(select|expand_backup|stem_go|heads)' -s 200 -c 8 -o gpurun_out/r1_run20_go_full \
  python scripts/bench_configs.py 4 > gpurun_out/ncu_go_full.log 2>&1
# MuZero: latent-space tree kernels and the hidden-state gather
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'rz_mz_' -s 60 -c 9 -o gpurun_out/r1_run20_mz_full \
  python scripts/bench_configs.py 5 > gpurun_out/ncu_mz_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/r1_run20_go_full.ncu-rep gpurun_out/r1_run20_go_ncu_full_summary.csv
python scripts/ncu_summary.py gpurun_out/r1_run20_mz_full.ncu-rep gpurun_out/r1_run20_mz_ncu_full_summary.csv
ls -la gpurun_out/*.ncu-rep
