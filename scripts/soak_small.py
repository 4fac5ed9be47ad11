#!/usr/bin/env python
"""Soak of the small-batch path (one-launch trunk, cluster heads kernel, fused backup + selection, programmatic dependent
launch, two captured graphs per search): many whole games of ONE game at a time through the reference API, and long
self-play of 1 / 37 / 64 / 148 concurrent games; fault words checked, wall clock bounded by the caller's timeout."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200.games.gomoku import GomokuEnv  # noqa: E402
from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent  # noqa: E402
from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet  # noqa: E402
from rlzero_b200.mcts import AlphaZeroPlayer  # noqa: E402
from rlzero_b200.selfplay import BatchedSelfPlay  # noqa: E402

torch.manual_seed(0)
net = ResNetPolicyValueNet(9, n_blocks=4).cuda().eval()
agent = AlphaZeroAgent(9, net=net)
t0 = time.time()
plies = games = 0
player = AlphaZeroPlayer(agent.policy_value_fn, n_playout=200, c_puct=5, is_selfplay=True)
while time.time() - t0 < 60:
    env = GomokuEnv(9, 5)
    env.reset()
    while True:
        move = player.get_action(env, temperature=1.0)
        env.step(move)
        plies += 1
        end, _ = env.game_end_winner()
        if end:
            break
    player.reset_player()
    games += 1
torch.cuda.synchronize()
print(json.dumps({'soak': 'one game at a time through AlphaZeroPlayer (9x9, ResNet-4, 200 playouts/move)', 'games': games,
                  'plies': plies, 'playouts': plies * 200, 'playouts_per_s': plies * 200 / (time.time() - t0)}), flush=True)
net15 = ResNetPolicyValueNet(15, n_blocks=10).cuda().eval()
for G in (1, 37, 64, 148):
    sp = BatchedSelfPlay(G, 15, 5, net=net15, n_playout=100, add_noise=True, seed=G)
    sp.set_random_start_positions()
    t0 = time.time()
    moves = 0
    while time.time() - t0 < 25:
        sp.play(5)
        moves += 5
        torch.cuda.synchronize()
        sp.forest.raise_faults()
    st = sp.stats()
    print(json.dumps({'soak': 'self-play, ResNet-10 15x15, 100 playouts/move', 'games': G, 'moves': moves,
                      'games_done': st['games_done'], 'plies_done': st['plies_done'],
                      'sims_per_s': G * moves * 100 / (time.time() - t0)}), flush=True)
