#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_go.py tests/test_gpu_go_search.py tests/test_gpu_selfplay.py tests/test_gpu_muzero.py -x -q 2>&1 | tail -6 > gpurun_out/r2_run53_tests.log
RZ_CFG=c4 RZ_WARM=3000 timeout 600 python scripts/wave_timeline.py > gpurun_out/r2_run53_wave_timeline_c4.log 2>&1
