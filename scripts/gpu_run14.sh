#!/bin/bash
# Stride-8 padded layout (boards up to 7x7: Connect Four, TicTacToe): parity tests, configs 1/2 throughput, wave launch list.
set -x
mkdir -p gpurun_out
python -m oracle.build_oracle
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_connect4.py tests/test_gpu_muzero.py tests/test_gpu_selfplay.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r1_run31_pytest_stride8.log 2>&1
tail -5 gpurun_out/r1_run31_pytest_stride8.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_run31_smoke.log 2>&1
tail -1 gpurun_out/r1_run31_smoke.log
timeout 600 python scripts/bench_configs.py 1 2 > gpurun_out/r1_run31_bench_configs_1_2.log 2>&1
cat gpurun_out/r1_run31_bench_configs_1_2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 64 --csv \
  --log-file gpurun_out/r1_run31_config2_wave_launches.csv python scripts/bench_configs.py 2 > gpurun_out/ncu_launch_c2.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r1_run31_config2_wave_launches.csv')) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hdr]; k = H.index('Kernel Name'); v = H.index('Metric Value'); u = H.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    t = float(r[v].replace(',', ''))
    if r[u] == 'ns': t /= 1000.0
    elif r[u] == 'ms': t *= 1000.0
    agg[r[k][:70]].append(t)
tot=sum(sum(x) for x in agg.values())
for name, ts in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print('%-72s n=%3d mean=%8.1f us share=%5.1f%%' % (name, len(ts), sum(ts) / len(ts), 100*sum(ts)/tot))
P
