#!/bin/bash
# Evidence for the kernels added in the last third of the round: ncu --set full of the tensor-core heads (15x15 and
# Go), the 8-stride stem / conv (Connect Four wave), and the final headline wave launch list.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 72 --csv \
  --log-file gpurun_out/r1_run37_wave_launches.csv python bench.py --steps 4 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r1_run37_wave_launches.csv')) if len(r) > 5]
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
print('total per wave %.1f us' % (tot/3))
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rz_heads_tc|rz_stem_tc|rz_expand_backup|rz_select' -s 40 -c 8 -o gpurun_out/r1_run37_heads_tc_full \
  python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/r1_run37_heads_tc_full.ncu-rep gpurun_out/r1_run37_heads_tc_ncu_full_summary.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rz_conv3x3_tc3|rz_stem_tc|rz_heads_tc' -s 60 -c 8 -o gpurun_out/r1_run37_c4_full \
  python scripts/bench_configs.py 2 > gpurun_out/ncu_full_c4.log 2>&1
python scripts/ncu_summary.py gpurun_out/r1_run37_c4_full.ncu-rep gpurun_out/r1_run37_c4_ncu_full_summary.csv
cat gpurun_out/r1_run37_heads_tc_ncu_full_summary.csv gpurun_out/r1_run37_c4_ncu_full_summary.csv | cut -c1-400
rm -f gpurun_out/r1_run37_c4_full.ncu-rep
