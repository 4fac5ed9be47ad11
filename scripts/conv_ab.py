#!/usr/bin/env python
"""A/B timing of the trunk convolution's variants on ONE box, interleaved (boxes differ by several % in their
power-capped clocks): flags bit 1 = direct-store epilogue, bit 2 = no epilogue (probe).  Full-size layer
(8140 boards = 110 tiles per CTA pair), with and without a residual input; median of 5 rounds of 30 launches."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200 import _lib as L  # noqa: E402


def main():
    lib = L.load()
    torch.manual_seed(0)
    B = 74 * 110
    w = (torch.randn(9, 128, 128, device='cuda') * 0.03).to(torch.bfloat16).contiguous()
    b = torch.zeros(128, device='cuda')
    x = (torch.randn(B * 256, 128, device='cuda') * 0.5).to(torch.bfloat16).contiguous()
    y = torch.empty_like(x)
    r = torch.empty_like(x).copy_(x)
    variants = [(f, res) for f in [int(a) for a in sys.argv[1:]] or [0, 2, 4] for res in (False, True)]
    times = {v: [] for v in variants}
    for rnd in range(5):
        for v in variants:
            fl, res = v

            def run():
                L.check(lib.rz_net_conv3x3_tc2(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(r) if res else None, L.ptr(y), B, 15, 15,
                                               128, 1, 2, fl, 0, L.stream_ptr()))
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(30):
                run()
            e1.record()
            torch.cuda.synchronize()
            times[v].append(e0.elapsed_time(e1) / 30 * 1e3)
    clk = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader'],
                         stdout=subprocess.PIPE).stdout.decode().strip()
    flops = 2.0 * B * 225 * 128 * 128 * 9
    for v in variants:
        med = float(np.median(times[v]))
        print(json.dumps({'flags': v[0], 'residual': v[1], 'us_median': med, 'us_all': [round(t, 1) for t in times[v]],
                          'us_per_tile': med / 110, 'TFLOPs_algorithmic': flops / med / 1e6}))
    print(json.dumps({'nvidia_smi_after': clk}))


if __name__ == '__main__':
    main()
