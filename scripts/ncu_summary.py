#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into a small CSV of the metrics the roofline uses.

    python scripts/ncu_summary.py gpurun_out/wave_full.ncu-rep profiles/r1_wave_ncu_summary.csv
"""
import csv
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex.sum',
    'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
]


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [m for m in METRICS if m in hdr]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['id', 'kernel'] + ['%s [%s]' % (m, units[hdr.index(m)]) for m in cols])
        for r in data:
            name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
            w.writerow([r[hdr.index('ID')], name] + [r[hdr.index(m)] for m in cols])
    print('wrote', out, len(data), 'launches')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
