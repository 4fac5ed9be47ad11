#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_run74_bench_2gpu.json 2> gpurun_out/r2_run74_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_run74_bench_reference_2gpu.json 2> gpurun_out/r2_run74_bench_reference_2gpu.err
