#!/bin/bash
# compute-sanitizer memcheck over the kernels on small problems (slow: minutes)
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_tree.py tests/test_rollout.py tests/test_connect4.py -x -q -k "not fullsize and not 15x15 and not 19" > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck_pytest.log
tail -5 gpurun_out/memcheck_pytest.log; grep -c "Invalid\|out of bounds" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log
