#!/bin/bash
mkdir -p gpurun_out
RZ_CFG=c4 RZ_WARM=3000 timeout 600 python scripts/wave_timeline.py > gpurun_out/r2_run51_wave_timeline_c4.log 2>&1
