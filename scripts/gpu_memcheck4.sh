#!/bin/bash
# compute-sanitizer over the kernels of round 2's small-batch work: one-launch trunk (DSMEM seam writes, weight ring),
# cluster heads kernel (DSMEM exchange), PDL launches -- memcheck and synccheck
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_run45_memcheck.log \
  python -m pytest tests/test_gpu_net.py -q -k "one_launch_trunk or heads_on_the_tensor or fused_head or batch_invariant" > gpurun_out/r2_run45_memcheck_pytest.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2_run45_memcheck_pytest.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file gpurun_out/r2_run45_synccheck.log \
  python -m pytest tests/test_gpu_net.py -q -k "one_launch_trunk or heads_on_the_tensor" > gpurun_out/r2_run45_synccheck_pytest.log 2>&1
echo "synccheck exit $?" >> gpurun_out/r2_run45_synccheck_pytest.log
tail -3 gpurun_out/r2_run45_memcheck_pytest.log gpurun_out/r2_run45_memcheck.log gpurun_out/r2_run45_synccheck_pytest.log gpurun_out/r2_run45_synccheck.log
