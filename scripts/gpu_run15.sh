#!/bin/bash
# Heads FC on the tensor cores (rz_net_heads_tc): parity, then timing against the CUDA-core heads kernel.
set -x
mkdir -p gpurun_out
python -m oracle.build_oracle
timeout 300 python -m pytest tests/test_gpu_net.py -m gpu -q -k "heads_on_the_tensor" > gpurun_out/r1_run33_pytest_heads_tc.log 2>&1
tail -15 gpurun_out/r1_run33_pytest_heads_tc.log
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_connect4.py tests/test_gpu_selfplay.py tests/test_gpu_api.py tests/test_gpu_go_search.py -m gpu -x -q > gpurun_out/r1_run33_pytest_net.log 2>&1
tail -5 gpurun_out/r1_run33_pytest_net.log
timeout 300 python - <<'P' > gpurun_out/r1_run33_heads_microbench.log 2>&1
import torch, numpy as np, json
from rlzero_b200 import _lib as L
import ctypes as C
from rlzero_b200.games.gomoku.policy_value_net import NativeForward, ResNetPolicyValueNet
for (h, w, a, planes, n) in [(15, 15, None, 4, 8192), (19, 19, 362, 17, 8192), (6, 7, 7, 4, 4096)]:
    kw = dict(board_width=w, in_planes=planes)
    if a: kw['n_actions'] = a
    net = ResNetPolicyValueNet(h, n_blocks=1, **kw).cuda().eval()
    nf = NativeForward(net, max_batch=n)
    nf.feat.uniform_(0, 1)
    lib = L.load()
    s = L.stream_ptr()
    def run_tc():
        L.check(lib.rz_net_heads_tc(C.byref(nf.hdesc), L.ptr(nf.feat), L.ptr(nf.logp), L.ptr(nf.value), n, s))
    def run_cc():
        L.check(lib.rz_net_heads(C.byref(nf.hdesc), L.ptr(nf.feat), 2, L.ptr(nf.logp), L.ptr(nf.value), n, s))
    out = {}
    for name, fn in (('tc', run_tc), ('cuda_core', run_cc)):
        for _ in range(5): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        out[name + '_us'] = e0.elapsed_time(e1) / 50 * 1e3
    print(json.dumps({'board': [h, w], 'AS': nf.AS, 'n': n, **out}), flush=True)
P
cat gpurun_out/r1_run33_heads_microbench.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1_run33_bench.json 2> gpurun_out/r1_run33_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r1_run33_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
timeout 600 python scripts/bench_configs.py 2 4 > gpurun_out/r1_run33_bench_configs_2_4.log 2>&1
cat gpurun_out/r1_run33_bench_configs_2_4.log
