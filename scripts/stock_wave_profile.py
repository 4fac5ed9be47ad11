#!/usr/bin/env python
"""A few waves of the reference's own PolicyValueNet on its default path (mode tc32) at config-3 size, for an ncu launch
list: ncu --metrics gpu__time_duration.sum --clock-control none -s <warm-up> -c 60 --csv --log-file out.csv python
scripts/stock_wave_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet  # noqa: E402
from rlzero_b200.selfplay import BatchedSelfPlay  # noqa: E402

torch.manual_seed(0)
sp = BatchedSelfPlay(8192, 15, 5, net=PolicyValueNet(15).cuda().eval(), n_playout=800, add_noise=True, seed=1)
sp.set_random_start_positions()
for _ in range(12):
    sp._wave()            # eager waves (no graph) so that every launch is visible to the profiler
torch.cuda.synchronize()
print('mode', sp.evaluator.mode)
