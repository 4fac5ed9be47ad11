#!/usr/bin/env python
"""Full-width soak runs of the widened rows: many committed moves for 8192 concurrent games, fault words checked.
  go      19x19 Go self-play (captures / ko / passes / move cap, re-rooting, trajectories, refills)
  dm      DeepMindMCTS flavour with the device rollout evaluator on 8192 Gomoku positions
  muzero  MuZero self-play on 8192 Gomoku boards"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200 import _lib as L  # noqa: E402
from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet  # noqa: E402


def soak_go(G=8192, moves=135, n_playout=48):
    from rlzero_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    net = ResNetPolicyValueNet(19, n_blocks=2, n_actions=362, in_planes=17).cuda().eval()
    sp = BatchedSelfPlay(G, 19, 1, net=net, n_playout=n_playout, add_noise=True, seed=5, game_type=L.GAME_GO,
                         komi=7.5, max_moves=120, ring_capacity=G * 64)
    sp.set_random_start_positions(max_random_moves=31)
    t0 = time.time()
    for m in range(moves):
        sp.play(1)
        if m % 10 == 9:
            torch.cuda.synchronize()
            sp.forest.raise_faults()
    torch.cuda.synchronize()
    sp.forest.raise_faults()
    st = sp.stats()
    meta = sp.forest.boards()[1]
    rows = sp.forest.boards()[0]
    stones = np.array([bin(int(x)).count('1') for x in rows.reshape(-1)]).reshape(G, -1).sum(1)
    out = sp.forest.drain_trajectories()
    z = out['info'][:, 2]
    print(json.dumps({'soak': 'go', 'games': G, 'moves': moves, 'n_playout': n_playout, 'games_done': st['games_done'],
                      'plies_done': st['plies_done'], 'ring_dropped': out['dropped'], 'mean_stones_on_board': float(stones.mean()),
                      'passes_seen': int((out['info'][:, 1] == 361).sum()), 'black_win_share': float((z[out['info'][:, 0] == 0] == 1).mean()) if len(z) else None,
                      'ko_active_now': int((meta[:, L.META_KO] >= 0).sum()), 'wall_s': time.time() - t0}), flush=True)


def soak_dm(G=8192, sims=200):
    from rlzero_b200.engine import SearchForest
    from rlzero_b200.mcts import RandomRolloutEvaluator
    f = SearchForest(G, 9, 5, n_playout=sims, c_puct=2.0, rule=L.RULE_PUCT, flavour=L.FLAVOUR_DEEPMIND, solve=True,
                     returns_mode=L.RETURNS_ZERO_SUM, max_carry=0, noise_root_only=True)
    rs = np.random.RandomState(0)
    f.set_positions([[int(x) for x in rs.permutation(81)[:rs.randint(0, 40)]] for _ in range(G)])
    live = (f.root_meta[:, L.META_STATUS] == L.ACTIVE).cpu().numpy()    # a random position may already be won
    ev = RandomRolloutEvaluator(n_rollouts=4, seed=1)
    t0 = time.time()
    f.run_waves(sims, ev, noise_eps=0.25, noise_alpha=0.25, seed=2)
    torch.cuda.synchronize()
    f.raise_faults()
    best, oc = f.best_child()
    print(json.dumps({'soak': 'dm', 'games': G, 'sims': sims, 'proven_roots': int((oc != 0).sum()),
                      'live_roots': int(live.sum()), 'mean_root_N': float(f.root_N.float().mean()),
                      'best_valid': bool((best[live] >= 0).all()),
                      'rollout_playouts_per_s': G * sims * 4 / (time.time() - t0), 'wall_s': time.time() - t0}), flush=True)


def soak_muzero(G=8192, moves=120):
    from rlzero_b200.muzero import BatchedMuZeroSelfPlay, MuZeroConfig, MuZeroNet
    torch.manual_seed(0)
    net = MuZeroNet(15, repr_blocks=1, dyn_blocks=1).cuda().eval()
    sp = BatchedMuZeroSelfPlay(G, 15, 5, net=net, config=MuZeroConfig(num_simulations=20), seed=3)
    t0 = time.time()
    for _ in range(moves):
        sp.play_move()
    torch.cuda.synchronize()
    sp.search.raise_faults()
    bad = int((sp.meta[:, L.META_FAULT] != 0).sum())
    print(json.dumps({'soak': 'muzero', 'games': G, 'moves': moves, 'games_done': sp.games_done, 'env_faults': bad,
                      'wall_s': time.time() - t0}), flush=True)
    assert bad == 0


def soak_line(name, G, size, k, moves, n_playout, blocks=2, **kw):
    """Full-width self-play of a line game (Gomoku / Connect Four) through many committed moves: re-rooting, refills,
    trajectories; every game must finish and drain cleanly, no fault bits."""
    from rlzero_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    net = ResNetPolicyValueNet(size, n_blocks=blocks, **{k_: v for k_, v in kw.items() if k_ in ('board_width', 'n_actions')}).cuda().eval()
    spkw = {k_: v for k_, v in kw.items() if k_ not in ('n_actions',)}
    sp = BatchedSelfPlay(G, size, k, net=net, n_playout=n_playout, add_noise=True, seed=9, ring_capacity=G * 96, **spkw)
    sp.set_random_start_positions(max_random_moves=5)
    t0 = time.time()
    for m in range(moves):
        sp.play(1)
        if m % 10 == 9:
            torch.cuda.synchronize()
            sp.forest.raise_faults()
    torch.cuda.synchronize()
    sp.forest.raise_faults()
    st = sp.stats()
    states, pis, zs, info = sp.drain()
    assert np.allclose(pis.sum(1), 1.0, atol=1e-4) and set(np.unique(zs)).issubset({-1.0, 0.0, 1.0})
    print(json.dumps({'soak': name, 'games': G, 'moves': moves, 'n_playout': n_playout, 'leaves_per_tree': sp.K,
                      'games_done': st['games_done'], 'plies_done': st['plies_done'], 'records_drained': int(len(zs)),
                      'mean_episode_plies': float(st['plies_done']) / max(1, st['games_done']),
                      'first_player_win_share': float((zs[info[:, 0] == 0] == 1).mean()) if len(zs) else None,
                      'wall_s': time.time() - t0}), flush=True)
    assert st['games_done'] >= G


if __name__ == '__main__':
    which = sys.argv[1:] or ['go', 'dm', 'muzero']
    if 'gomoku' in which:
        soak_line('gomoku 15x15', 8192, 15, 5, 160, 32)
    if 'c4' in which:
        soak_line('connect four 6x7 (8-stride layout)', 4096, 6, 4, 60, 32, board_width=7, n_actions=7,
                  game_type=L.GAME_CONNECT4)
    if 'leafpar' in which:
        soak_line('gomoku 9x9, leaf-parallel K=4', 2048, 9, 5, 100, 33, leaves_per_tree=4)
    if 'go' in which:
        soak_go()
    if 'dm' in which:
        soak_dm()
    if 'muzero' in which:
        soak_muzero()
