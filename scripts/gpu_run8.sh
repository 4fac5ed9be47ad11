#!/bin/bash
# Go at full size: config 4 throughput + the launch list of its wave (tree kernels with Go rules inside).
set -x
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py 4 > gpurun_out/r1_run16_bench_config4_go.log 2>&1
tail -3 gpurun_out/r1_run16_bench_config4_go.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 140 --csv \
  --log-file gpurun_out/r1_run16_go_wave_launches.csv python scripts/bench_configs.py 4 > gpurun_out/ncu_go.log 2>&1
tail -3 gpurun_out/ncu_go.log
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r1_run16_go_wave_launches.csv')) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hdr]; k = H.index('Kernel Name'); v = H.index('Metric Value'); u = H.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    t = float(r[v].replace(',', ''))
    if r[u] == 'ns': t /= 1000.0
    elif r[u] == 'ms': t *= 1000.0
    agg[r[k][:60]].append(t)
for name, ts in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print('%-62s n=%3d mean=%9.1f us total=%10.1f us' % (name, len(ts), sum(ts) / len(ts), sum(ts)))
P
