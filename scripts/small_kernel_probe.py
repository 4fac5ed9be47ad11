#!/usr/bin/env python
"""Warm (graph-replayed) durations of the kernels of a single-board ResNet-10 evaluation: each kernel captured 50 times in
one CUDA graph, replayed, timed with CUDA events."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200 import _lib as L  # noqa: E402
from rlzero_b200.games.gomoku.policy_value_net import NativeForward, ResNetPolicyValueNet  # noqa: E402

torch.manual_seed(0)
net = ResNetPolicyValueNet(15, n_blocks=10).cuda().eval()
for n in (1, 64):
    nf = NativeForward(net, max_batch=n)
    x = (torch.rand(n, 4, 15, 15, device='cuda') < 0.3).float()
    nf.forward_planes(x)
    lib, ts = nf.lib, nf.trunk_small
    w1 = nf.w1x1_host.ctypes.data_as(C.c_void_p)
    b1 = nf.b1x1_host.ctypes.data_as(C.c_void_p)

    def trunk():
        L.check(lib.rz_net_trunk_small(L.ptr(nf.bufs[0]), L.ptr(ts['w']), L.ptr(ts['b']), ts['n'], ts['relu_mask'],
                                       ts['res_mask'], n, 15, 15, w1, b1, L.ptr(nf.feat), L.stream_ptr()), 'trunk')

    def heads():
        L.check(lib.rz_net_heads_tc(C.byref(nf.hdesc), L.ptr(nf.feat), L.ptr(nf.logp), L.ptr(nf.value), n,
                                    L.stream_ptr()), 'heads')

    def layer():
        l = nf.layers[1]
        L.check(lib.rz_net_conv3x3_tc2(L.ptr(nf.bufs[0]), L.ptr(l['w']), L.ptr(l['b']), None, L.ptr(nf.bufs[1]), n, 15, 15,
                                       128, 1, 2, 514, 0, L.stream_ptr()), 'conv')

    def full():
        nf.forward_planes(x)

    if n == 1:
        for name, fn, per, what in (
                ('trunk_small', trunk, ts['n'], 'per layer: inputs ready, MMAs issued, accumulator complete, TMEM read, '
                 'tile stored, proxy fence, warp sync, arrived'),
                ('heads_tc', heads, 24, 'per chunk: top, A free, converted, proxy fence, block sync, weights landed, '
                 'MMAs issued')):
            probe = torch.zeros(8 * per, dtype=torch.int64, device='cuda')
            lib.rz_debug_set_probe(L.ptr(probe))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            lib.rz_debug_set_probe(None)
            t = probe.cpu().numpy().reshape(-1, 8)
            t = t - t[0, 0]
            print(json.dumps({'kernel': name, 'probe_ns': what, 'stamps': t[:4].tolist() + t[-2:].tolist()}), flush=True)
    for name, fn in (('trunk_small', trunk), ('heads_tc', heads), ('conv_tc2_layer', layer), ('forward_planes', full)):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(50):
                    fn()
            for _ in range(3):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(s)
            for _ in range(10):
                g.replay()
            e1.record(s)
            torch.cuda.synchronize()
        print(json.dumps({'boards': n, 'kernel': name, 'us': e0.elapsed_time(e1) * 1000 / 500,
                          'pdl': os.environ.get('RZ_PDL', '1')}), flush=True)
