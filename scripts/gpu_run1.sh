#!/bin/bash
# GPU session: parity tests, bench line, ncu launch list, ncu full capture of one wave.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py --steps 800 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 800 --warmup 8 --cpu-seconds 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch list of one steady-state wave (skip 300 waves x 25 kernels)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 7500 -c 75 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 300 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
# full capture of one wave's kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rz_(conv3x3_tc|select|expand_backup|heads|gomoku_encode)' \
  -s 7500 -c 25 -o gpurun_out/wave_full python bench.py --steps 4 --warmup 300 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
