#!/usr/bin/env python
"""Timeline of the kernels of a few graph-replayed waves (CUPTI through torch.profiler): start / duration of every kernel and
the gap to its predecessor, at the headline size after a sustained warm-up.  RZ_G / RZ_WARM override games / warm-up waves."""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet  # noqa: E402
from rlzero_b200.selfplay import BatchedSelfPlay  # noqa: E402

G = int(os.environ.get('RZ_G', '8192'))
torch.manual_seed(0)
if os.environ.get('RZ_CFG') == 'c4':      # BASELINE config 2: Connect Four 6x7, ResNet-6, 4096 games
    from rlzero_b200 import _lib as L
    G = int(os.environ.get('RZ_G', '4096'))
    net = ResNetPolicyValueNet(6, n_blocks=6, board_width=7, n_actions=7).cuda().eval()
    sp = BatchedSelfPlay(G, 6, 4, net=net, n_playout=200, add_noise=True, seed=2, board_width=7, game_type=L.GAME_CONNECT4)
else:
    net = ResNetPolicyValueNet(15, n_blocks=10).cuda().eval()
    sp = BatchedSelfPlay(G, 15, 5, net=net, n_playout=800, add_noise=True, seed=1)
sp.set_random_start_positions(max_random_moves=3) if os.environ.get('RZ_CFG') == 'c4' else sp.set_random_start_positions()
sp.warm_up()
for _ in range(int(os.environ.get('RZ_WARM', '1500'))):
    sp.step_wave()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(6):
        sp.step_wave()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and 'Memcpy' not in e.name
      and 'Memset' not in e.name]
ev.sort(key=lambda e: e.time_range.start)
rows = []
prev_end = None
for e in ev:
    st, en = e.time_range.start, e.time_range.end
    rows.append({'name': e.name[:48], 'start_us': st, 'dur_us': en - st, 'gap_us': None if prev_end is None else st - prev_end})
    prev_end = en
# the middle waves only
n = len(rows)
per = n // 6 if n >= 6 else n
mid = rows[2 * per:4 * per]
tot = mid[-1]['start_us'] + mid[-1]['dur_us'] - mid[0]['start_us'] if mid else 0
print(json.dumps({'games': G, 'kernels_recorded': n, 'kernels_per_wave': per, 'two_waves_us': tot,
                  'sum_dur_us': sum(r['dur_us'] for r in mid), 'sum_gap_us': sum((r['gap_us'] or 0) for r in mid[1:])}))
for r in mid[:per]:
    print('%-50s dur %8.1f gap %6.1f' % (r['name'], r['dur_us'], r['gap_us'] or 0))
