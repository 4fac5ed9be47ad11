#!/bin/bash
# Per-tap weight barriers + activation-first load order in the conv kernels: parity, then timing.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -x -q -k "conv3x3" > gpurun_out/r1_run47_pytest_conv.log 2>&1
tail -4 gpurun_out/r1_run47_pytest_conv.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_connect4.py tests/test_gpu_muzero.py tests/test_gpu_fullsize.py tests/test_gpu_go_search.py -m gpu -x -q > gpurun_out/r1_run47_pytest.log 2>&1
tail -4 gpurun_out/r1_run47_pytest.log | cut -c1-250
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1_run47_bench.json 2> gpurun_out/r1_run47_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r1_run47_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['launch_ms'], d['roofline']['launch_ms_after_sustained_run'])"
timeout 600 python scripts/bench_configs.py 2 4 stock15 > gpurun_out/r1_run47_bench_configs.log 2>&1
cut -c1-230 gpurun_out/r1_run47_bench_configs.log
