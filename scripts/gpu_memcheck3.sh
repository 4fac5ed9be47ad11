#!/bin/bash
# compute-sanitizer over the kernels added late in round 1: leaf-parallel select / expand+backup (memcheck + racecheck is
# not applicable: no shared memory), tensor-core heads and stride-8 kernels (synccheck).
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r1_run39_memcheck.log \
  python -m pytest tests/test_gpu_leaf_parallel.py tests/test_gpu_tree.py -q -k "not 15 and not 19 and not fullsize" > gpurun_out/r1_run39_memcheck_pytest.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r1_run39_memcheck_pytest.log
tail -4 gpurun_out/r1_run39_memcheck_pytest.log; tail -3 gpurun_out/r1_run39_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file gpurun_out/r1_run39_synccheck.log \
  python -m pytest tests/test_gpu_net.py -q -k "heads_on_the_tensor or row_stride_8" > gpurun_out/r1_run39_synccheck_pytest.log 2>&1
echo "synccheck exit $?" >> gpurun_out/r1_run39_synccheck_pytest.log
tail -4 gpurun_out/r1_run39_synccheck_pytest.log; tail -3 gpurun_out/r1_run39_synccheck.log
