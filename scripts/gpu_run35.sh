#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 0; do RZ_PDL=$pdl timeout 300 python scripts/small_batch_probe.py > gpurun_out/r2_run35_small_batch_pdl$pdl.log 2>&1; done
RZ_EAGER=1 RZ_G=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 72 --csv \
  --log-file gpurun_out/r2_run35_g1_wave_launches.csv python scripts/small_batch_probe.py > gpurun_out/r2_run35_ncu.log 2>&1
