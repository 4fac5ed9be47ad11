#!/usr/bin/env python
"""One ResNet-10 / batch 4096 training step for an ncu launch list (profiles/r2_*train_step_launches.csv):
    ncu --metrics gpu__time_duration.sum --clock-control none -s <warm-up launches> -c 600 --csv --log-file out.csv \\
        python scripts/train_step_profile.py
Also prints CUDA-event timings of trainer.learn alone (no inference-weight re-pack) and of agent.learn."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent  # noqa: E402
from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet, ResNetPolicyValueNet  # noqa: E402

STOCK = os.environ.get('RZ_TRAIN_NET', 'resnet') == 'stock'      # the reference's own network instead of ResNet-10
B = int(os.environ.get('RZ_TRAIN_B', 512 if STOCK else 4096))
steps = int(os.environ.get('RZ_TRAIN_STEPS', 3))
rs = np.random.RandomState(0)
x = torch.from_numpy((rs.rand(B, 4, 15, 15) < 0.2).astype(np.float32)).cuda()
pi = torch.from_numpy(rs.dirichlet(0.3 * np.ones(225), size=B).astype(np.float32)).cuda()
z = torch.from_numpy(rs.choice([-1.0, 0.0, 1.0], size=B).astype(np.float32)).cuda()
torch.manual_seed(0)
agent = AlphaZeroAgent(15, net=PolicyValueNet(15) if STOCK else ResNetPolicyValueNet(15, n_blocks=10))
agent.learn(x, pi, z)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(steps):
    agent.trainer.learn(x, pi, z)
e[1].record()
for _ in range(steps):
    agent.learn(x, pi, z)
e[2].record()
torch.cuda.synchronize()
print('trainer.learn %.2f ms/step, agent.learn (with refresh_weights) %.2f ms/step, batch %d' % (
    e[0].elapsed_time(e[1]) / steps, e[1].elapsed_time(e[2]) / steps, B))
