#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/wave_timeline.py > gpurun_out/r2_run48_wave_timeline_8192.log 2>&1
RZ_G=1 RZ_WARM=200 timeout 300 python scripts/wave_timeline.py > gpurun_out/r2_run48_wave_timeline_1.log 2>&1
