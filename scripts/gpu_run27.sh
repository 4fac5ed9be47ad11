#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_connect4.py tests/test_gpu_muzero.py tests/test_gpu_selfplay.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r1_run41_pytest.log 2>&1
tail -5 gpurun_out/r1_run41_pytest.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:rz_stem_tc -s 10 -c 4 python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | grep -E "gpu__time|dram__bytes" | head
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1_run41_bench.json 2> gpurun_out/r1_run41_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r1_run41_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
