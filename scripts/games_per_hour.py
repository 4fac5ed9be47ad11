#!/usr/bin/env python
"""Measured self-play games/hour of config 3 (the second half of BASELINE.json's metric): play whole
games (search + sampled move + re-root + refill) for a fixed number of moves and count finished
episodes.  Long (minutes); bench.py only estimates this figure from simulations/s.

    python scripts/games_per_hour.py [moves] [games]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet  # noqa: E402
from rlzero_b200.selfplay import BatchedSelfPlay  # noqa: E402


def main():
    moves = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    torch.manual_seed(0)
    net = ResNetPolicyValueNet(15, n_blocks=10).cuda().eval()
    sp = BatchedSelfPlay(G, 15, 5, net=net, n_playout=800, c_puct=5.0, temperature=1.0, add_noise=True, seed=1234)
    sp.forest.reset_games()          # every game from the empty board: unbiased episode lengths
    sp.warm_up()
    torch.cuda.synchronize()
    t0 = time.time()
    log = []
    for m in range(moves):
        n = sp.n_playout - sp.waves_in_move
        for _ in range(n):
            sp.step_wave()
        if (m + 1) % 10 == 0 or m == moves - 1:
            torch.cuda.synchronize()
            st = sp.stats()
            dt = time.time() - t0
            log.append(dict(move=m + 1, wall_s=dt, games_done=st['games_done'], plies_done=st['plies_done']))
            print(json.dumps(log[-1]), flush=True)
    sp.forest.raise_faults()
    st = sp.stats()
    dt = time.time() - t0
    mean_len = st['plies_done'] / max(1, st['games_done'])
    sims = float(G) * 800 * moves
    out = dict(metric='selfplay_games_per_hour', games=G, moves_played=moves, wall_s=dt,
               games_done=st['games_done'], mean_episode_plies=mean_len,
               games_per_hour_counted=st['games_done'] / dt * 3600.0,
               simulations_per_s=sims / dt,
               games_per_hour_steady_state=(sims / dt) / (800.0 * mean_len) * 3600.0 if st['games_done'] else None,
               note='counted = episodes finished within the run / wall time (first episodes only finish after '
                    '~mean_episode_plies moves, so it under-counts); steady_state = simulations/s / (800 x mean '
                    'finished-episode length)')
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
