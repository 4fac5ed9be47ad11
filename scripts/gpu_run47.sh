#!/bin/bash
# 2 x B200: the bench line under torchrun (exchange block, shard hash) and the generation loop, on the final kernels
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-configs > gpurun_out/r2_run47_bench_2gpu.json 2> gpurun_out/r2_run47_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  scripts/generation_loop.py > gpurun_out/r2_run47_generation_loop_2gpu.log 2>&1
tail -3 gpurun_out/r2_run47_generation_loop_2gpu.log
