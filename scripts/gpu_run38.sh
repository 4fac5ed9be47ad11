#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 0; do RZ_PDL=$pdl timeout 300 python scripts/small_kernel_probe.py > gpurun_out/r2_run38_small_kernels_pdl$pdl.log 2>&1; done
