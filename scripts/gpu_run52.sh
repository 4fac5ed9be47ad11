#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_go.py tests/test_gpu_go_search.py tests/test_gpu_selfplay.py tests/test_gpu_fullsize.py tests/test_gpu_muzero.py -x -q 2>&1 | tail -6 > gpurun_out/r2_run52_tests.log
