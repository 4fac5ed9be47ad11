#!/bin/bash
# sustained A/B on one box: staged TMA-store epilogue (flags 0) vs direct-store epilogue (flags 2)
mkdir -p gpurun_out
for f in 0 2 0 2; do
  RZ_CONV_FLAGS=$f timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'conv_flags': $f, 'sims_per_s': d['value'], 'ms_per_wave': d['ms_per_step'], 'sm_mhz': d['clocks']['sm_mhz'], 'reasons': d['clocks']['reasons'], 'conv_cold_ms': d['roofline']['launch_ms'], 'conv_hot_ms': d['roofline']['launch_ms_after_sustained_run']}))"
done | tee gpurun_out/r1_run24_bench_ab.log
