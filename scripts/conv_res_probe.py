#!/usr/bin/env python
"""Why do the layers with a residual input run at 75 % instead of 93 % tensor-pipe activity?  One full-size
launch per variant (to be read under `ncu --metrics sm__cycles_elapsed.avg,...`, cycles are clock-independent):
  0 no residual   1 residual = a separate 537 MB tensor (DRAM stream)   2 residual = the input tensor itself
  (the same rows the TMA loads fetch: L2 hits)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200 import _lib as L  # noqa: E402

lib = L.load()
torch.manual_seed(0)
B = 74 * 110
w = (torch.randn(9, 128, 128, device='cuda') * 0.03).to(torch.bfloat16).contiguous()
b = torch.zeros(128, device='cuda')
x = (torch.randn(B * 256, 128, device='cuda') * 0.5).to(torch.bfloat16).contiguous()
y = torch.empty_like(x)
r = torch.empty_like(x).copy_(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for rep in range(2):
    for res in (None, r, x):
        flush.zero_()
        L.check(lib.rz_net_conv3x3_tc2(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(res) if res is not None else None, L.ptr(y), B,
                                       15, 15, 128, 1, 2, 2, 0, L.stream_ptr()))
torch.cuda.synchronize()
