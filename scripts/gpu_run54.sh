#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -x -q -k "small_and_large or one_launch" 2>&1 | tail -6 > gpurun_out/r2_run54_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-configs > gpurun_out/r2_run54_bench_2gpu.json 2> gpurun_out/r2_run54_bench_2gpu.err
