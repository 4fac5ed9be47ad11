#!/usr/bin/env python
"""Probe the conv v2 variants on a B200: correctness against torch and timing.

    python scripts/conv2_probe.py one <cta_group> <flags>     # one variant (may hang -> run under timeout)
    python scripts/conv2_probe.py all                          # every variant in a subprocess with a timeout
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def to_tile(x):
    import torch
    n, c, h, w = x.shape
    t = torch.zeros(n, 16, 16, c, dtype=torch.bfloat16, device=x.device)
    t[:, :h, :w, :] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return t.reshape(n, 256, c).contiguous()


def run_one(cg, flags):
    import torch
    from rlzero_b200 import _lib as L
    lib = L.load()
    dev = 'cuda'
    out_lines = []
    for (n, h, cin, relu, res) in [(1, 15, 128, 1, 0), (3, 15, 128, 1, 1), (5, 9, 64, 0, 0), (300, 15, 128, 1, 1)]:
        torch.manual_seed(n * 100 + h)
        x = (torch.randn(n, cin, h, h, device=dev) * 0.5).to(torch.bfloat16).float()
        w = (torch.randn(128, cin, 3, 3, device=dev) / (3.0 * cin ** 0.5)).to(torch.bfloat16).float()
        b = torch.randn(128, device=dev) * 0.1
        r = (torch.randn(n, 128, h, h, device=dev) * 0.5).to(torch.bfloat16).float() if res else None
        ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
        if res:
            ref = ref + r.double()
        if relu:
            ref = torch.relu(ref)
        xt = to_tile(x)
        wt = w.permute(2, 3, 0, 1).reshape(9, 128, cin).to(torch.bfloat16).contiguous()
        out = torch.full((n, 256, 128), 7.0, dtype=torch.bfloat16, device=dev)
        rt = None
        if res:
            out.copy_(to_tile(r))
            rt = out
        if cg == 0:
            L.check(lib.rz_net_conv3x3_tc(L.ptr(xt), L.ptr(wt), L.ptr(b), L.ptr(rt), L.ptr(out), n, h, cin, relu,
                                          0, L.stream_ptr()), 'conv v1')
        else:
            L.check(lib.rz_net_conv3x3_tc2(L.ptr(xt), L.ptr(wt), L.ptr(b), L.ptr(rt), L.ptr(out), n, h, h, cin, relu,
                                           cg, flags, 0, L.stream_ptr()), 'conv v2')
        torch.cuda.synchronize()
        got = out.reshape(n, 16, 16, 128)[:, :h, :h, :].permute(0, 3, 1, 2).double()
        err = (got - ref).abs().max().item()
        pad = max(out.reshape(n, 16, 16, 128)[:, h:].abs().max().item(),
                  out.reshape(n, 16, 16, 128)[:, :, h:].abs().max().item())
        out_lines.append(dict(case=[n, h, cin, relu, res], max_err=err, scale=ref.abs().max().item(), pad=pad))
    # timing at the bench size
    G = 8192
    x = torch.randn(G, 256, 128, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    w = (torch.randn(9, 128, 128, device=dev) * 0.03).to(torch.bfloat16)
    b = torch.zeros(128, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    for i in range(3 + reps):
        if i == 3:
            e0.record()
        if cg == 0:
            lib.rz_net_conv3x3_tc(L.ptr(x), L.ptr(w), L.ptr(b), None, L.ptr(y), G, 15, 128, 1, 0, L.stream_ptr())
        else:
            lib.rz_net_conv3x3_tc2(L.ptr(x), L.ptr(w), L.ptr(b), None, L.ptr(y), G, 15, 15, 128, 1, cg, flags, 0,
                                   L.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps(dict(cta_group=cg, flags=flags, cases=out_lines, ms=ms,
                          tflops_alg=2.0 * G * 225 * 128 * 1152 / ms / 1e9,
                          tflops_issued=2.0 * G * 256 * 128 * 1152 / ms / 1e9)))


if __name__ == '__main__':
    if sys.argv[1] == 'one':
        run_one(int(sys.argv[2]), int(sys.argv[3]))
    else:
        for cg, fl in [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1)]:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), 'one', str(cg), str(fl)],
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=150)
                print('variant cg=%d flags=%d rc=%d\n%s' % (cg, fl, r.returncode, r.stdout.decode()[-3000:]), flush=True)
            except subprocess.TimeoutExpired:
                print('variant cg=%d flags=%d TIMEOUT' % (cg, fl), flush=True)
