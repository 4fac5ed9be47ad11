#!/usr/bin/env python
"""Single-game search through the reference API (AlphaZeroMCTS.simulate on ONE env): the sequential parity mode
(one leaf per wave) against the opt-in leaf-parallel mode (leaves_per_wave = K, virtual loss).  One JSON line per
setting: playouts/s of one tree."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200.games.gomoku import GomokuEnv  # noqa: E402
from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent  # noqa: E402
from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet  # noqa: E402
from rlzero_b200.mcts import AlphaZeroMCTS  # noqa: E402


def main():
    torch.manual_seed(0)
    for name, size, k, net, n_playout in (('config1 TicTacToe 3x3, stock PolicyValueNet fp32', 3, 3, None, 25 * 40),
                                          ('Gomoku 15x15, stock PolicyValueNet fp32', 15, 5, None, 1600),
                                          ('Gomoku 15x15, ResNet-10 bf16', 15, 5, 10, 1600)):
        agent = AlphaZeroAgent(size, net=ResNetPolicyValueNet(size, n_blocks=net) if net else None)
        agent.policy_value_net.eval()
        env = GomokuEnv(size, k)
        env.reset()
        for K in (1, 8, 32, 128):
            mcts = AlphaZeroMCTS(agent.policy_value_fn, n_playout=n_playout, c_puct=5, leaves_per_wave=K)
            mcts.simulate(env, 1.0)          # warm-up: pools, graph capture
            mcts.update_with_move(-1)
            torch.cuda.synchronize()
            t0 = time.time()
            mcts.simulate(env, 1.0)
            torch.cuda.synchronize()
            dt = time.time() - t0
            print(json.dumps({'case': name, 'leaves_per_wave': K, 'n_playout': n_playout, 'seconds': dt,
                              'playouts_per_s': n_playout / dt}), flush=True)


if __name__ == '__main__':
    main()
