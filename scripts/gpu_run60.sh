#!/bin/bash
# 4 x B200 under torchrun: the bench line with its exchange block on the last commit
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_run60_bench_4gpu.json 2> gpurun_out/r2_run60_bench_4gpu.err
