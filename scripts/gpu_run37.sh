#!/bin/bash
mkdir -p gpurun_out
RZ_EAGER=1 RZ_G=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rz_(select|stem|trunk|heads|expand)" -s 20 -c 15 --csv \
  --log-file gpurun_out/r2_run37_g1_wave_launches.csv python scripts/small_batch_probe.py > gpurun_out/r2_run37_ncu.log 2>&1
