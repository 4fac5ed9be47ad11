#!/bin/bash
# GS Gamma sampler in the expansion kernel: noise distribution test, noisy self-play tests, expand/backup timing, bench.
set -x
mkdir -p gpurun_out
python -m oracle.build_oracle
timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_selfplay.py tests/test_gpu_fullsize.py tests/test_gpu_dm.py tests/test_connect4.py tests/test_gpu_go_search.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r1_run38_pytest.log 2>&1
tail -6 gpurun_out/r1_run38_pytest.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1_run38_bench.json 2> gpurun_out/r1_run38_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r1_run38_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print(json.dumps(d['tree_roofline'], indent=1))"
tail -3 gpurun_out/r1_run38_bench.err
