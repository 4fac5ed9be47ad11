#!/bin/bash
# Full GPU suite after the 8-stride layout and the tensor-core heads; MuZero throughput with the TC heads; memcheck of the new kernels.
set -x
mkdir -p gpurun_out
python -m oracle.build_oracle
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_run34_pytest_gpu.log 2>&1
tail -4 gpurun_out/r1_run34_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_run34_smoke.log 2>&1
tail -1 gpurun_out/r1_run34_smoke.log
timeout 600 python scripts/bench_configs.py 5 > gpurun_out/r1_run34_bench_config5.log 2>&1
cat gpurun_out/r1_run34_bench_config5.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r1_run34_memcheck.log \
  python -m pytest tests/test_gpu_net.py -q -k "heads_on_the_tensor or row_stride_8 or fused_head" > gpurun_out/r1_run34_memcheck_pytest.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r1_run34_memcheck_pytest.log
tail -3 gpurun_out/r1_run34_memcheck_pytest.log; tail -3 gpurun_out/r1_run34_memcheck.log
