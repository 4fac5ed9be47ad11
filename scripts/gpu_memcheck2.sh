#!/bin/bash
# compute-sanitizer memcheck + synccheck over the Go / DeepMindMCTS / MuZero kernels (small problems).
set -x
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r1_run21_memcheck.log \
  python -m pytest tests/test_gpu_go.py tests/test_gpu_go_search.py tests/test_gpu_dm.py tests/test_gpu_muzero.py -x -q \
  -k "not 19 and not 15-2-2" > gpurun_out/r1_run21_memcheck_pytest.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r1_run21_memcheck_pytest.log
tail -4 gpurun_out/r1_run21_memcheck_pytest.log; tail -3 gpurun_out/r1_run21_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file gpurun_out/r1_run21_synccheck.log \
  python -m pytest tests/test_gpu_go.py tests/test_gpu_dm.py -x -q -k "not 19 and not 9-8" > gpurun_out/r1_run21_synccheck_pytest.log 2>&1
echo "synccheck exit $?" >> gpurun_out/r1_run21_synccheck_pytest.log
tail -4 gpurun_out/r1_run21_synccheck_pytest.log; tail -3 gpurun_out/r1_run21_synccheck.log
