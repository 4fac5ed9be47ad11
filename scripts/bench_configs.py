#!/usr/bin/env python
"""Throughput of the parity-case configurations of BASELINE.json (NOT the bench line, which is
config 3 in bench.py): config 1 TicTacToe / 25 sims / stock net, config 2 Connect Four / 200 sims /
4096 games / ResNet-6, config 4 Go 19x19 / 800 sims / ResNet-20, config 5 MuZero / 50 latent sims.  One JSON line per configuration."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200 import _lib as L  # noqa: E402
from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet, ResNetPolicyValueNet  # noqa: E402
from rlzero_b200.selfplay import BatchedSelfPlay  # noqa: E402


def run(name, sp, waves, warm, random_moves=3):
    sp.set_random_start_positions(max_random_moves=random_moves)
    sp.warm_up()
    for _ in range(warm):
        sp.step_wave()
    torch.cuda.synchronize()
    g0 = sp.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(waves):
        sp.step_wave()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    g1 = sp.stats()
    sp.forest.raise_faults()
    games = g1['games_done'] - g0['games_done']
    print(json.dumps({'config': name, 'games': sp.G, 'n_playout': sp.n_playout, 'waves': waves,
                      'ms_per_wave': ms / waves, 'simulations_per_s': sp.G * waves / ms * 1e3,
                      'games_finished': games, 'games_per_hour': games / (ms / 1e3) * 3600.0,
                      'wall_s': time.time() - t0}), flush=True)


def main():
    only = set(sys.argv[1:])           # e.g. `bench_configs.py 4` runs config 4 alone
    torch.manual_seed(0)
    if not only or '1' in only:
        # config 1: TicTacToe = GomokuEnv(3, 3), 25 simulations/move, stock PolicyValueNet (fp32 path)
        net1 = PolicyValueNet(3).cuda().eval()
        for G in (1, 8192):
            sp = BatchedSelfPlay(G, 3, 3, net=net1, n_playout=25, add_noise=True, seed=1)
            run('config1 TicTacToe 3x3 k=3, 25 sims/move, stock PolicyValueNet fp32, %d game(s)' % G, sp,
                25 * 40, 25)
    if not only or '2' in only:
        # config 2: Connect Four 6x7, 200 simulations/move, 4096 games, ResNet-6 bf16
        net2 = ResNetPolicyValueNet(6, n_blocks=6, board_width=7, n_actions=7).cuda().eval()
        sp = BatchedSelfPlay(4096, 6, 4, net=net2, n_playout=200, add_noise=True, seed=2, board_width=7,
                             game_type=L.GAME_CONNECT4)
        run('config2 Connect Four 6x7, 200 sims/move, ResNet-6 bf16, 4096 games', sp, 200 * 12, 200)
        del sp
        torch.cuda.empty_cache()
    if not only or '4' in only:
        # config 4: Go 19x19 (GoEnv rules: captures, ko, suicide, pass, Tromp-Taylor, komi 7.5), 800
        # simulations/move, ResNet-20 bf16 on the tensor cores, 17-plane observation, 362 actions;
        # games start after up to 30 random legal moves; engine-side cap of 722 moves per game
        net4 = ResNetPolicyValueNet(19, n_blocks=20, n_actions=362, in_planes=17).cuda().eval()
        sp = BatchedSelfPlay(8192, 19, 1, net=net4, n_playout=800, add_noise=True, seed=3, game_type=L.GAME_GO,
                             komi=7.5, max_moves=722)
        run('config4 Go 19x19 (komi 7.5), 800 sims/move, ResNet-20 bf16, 8192 games', sp, 160, 8, random_moves=31)
        del sp
        torch.cuda.empty_cache()
    if not only or '5' in only:
        # config 5: MuZero on Gomoku 15x15, 50 simulations/move in latent space, 8192 games: h = stem + 10
        # blocks on the real position, g = conv + 5 blocks on (hidden state, action plane), f = the heads
        from rlzero_b200.muzero import BatchedMuZeroSelfPlay, MuZeroConfig, MuZeroNet
        net5 = MuZeroNet(15, repr_blocks=10, dyn_blocks=5).cuda().eval()
        G = 8192
        sp = BatchedMuZeroSelfPlay(G, 15, 5, net=net5, config=MuZeroConfig(num_simulations=50), seed=4)
        for _ in range(3):
            sp.play_move()
        torch.cuda.synchronize()
        g0, n_moves = sp.games_done, 12
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(n_moves):
            sp.play_move()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        sp.search.raise_faults()
        sims = G * 50 * n_moves
        flops = G * n_moves * (net5.flops('initial') + 50 * net5.flops('recurrent'))
        print(json.dumps({'config': 'config5 MuZero Gomoku 15x15, 50 latent sims/move, h=ResNet-10 g=ResNet-5 bf16, '
                                    '%d games' % G, 'games': G, 'n_playout': 50, 'moves': n_moves,
                          'ms_per_move': ms / n_moves, 'simulations_per_s': sims / ms * 1e3,
                          'moves_per_s': G * n_moves / ms * 1e3, 'net_tflops': flops / ms / 1e9,
                          'hbm_gb': sp.search.hbm_bytes() / 1e9, 'kernels_per_move': sp.search.kernels_per_move(),
                          'games_finished': sp.games_done - g0, 'wall_s': time.time() - t0}), flush=True)
        del sp
        torch.cuda.empty_cache()
    if 'stock15' in only:
        # the reference's OWN network (PolicyValueNet 4 -> 32 -> 64 -> 128) at 15x15 / 800 sims / 8192 games: the
        # fp32 CUDA-core path (1e-5 parity, the default for this net) against the same weights zero-padded onto the
        # bf16 tensor-core trunk (mode 'tc', 1e-3)
        net = PolicyValueNet(15).cuda().eval()
        for mode, waves in (('f32', 24), ('tc', 400)):
            sp = BatchedSelfPlay(8192, 15, 5, net=net, n_playout=800, add_noise=True, seed=1, net_mode=mode)
            run('stock PolicyValueNet 15x15, 800 sims/move, 8192 games, mode %s' % mode, sp, waves, 4)
            del sp
            torch.cuda.empty_cache()
    if 'stock3' in only:
        # config 1's network and board (TicTacToe, stock PolicyValueNet) with 8192 games on the tensor-core path
        # (8-stride layout, layers zero-padded to 128 channels) against the fp32 default
        net = PolicyValueNet(3).cuda().eval()
        for mode in ('f32', 'tc'):
            sp = BatchedSelfPlay(8192, 3, 3, net=net, n_playout=25, add_noise=True, seed=1, net_mode=mode)
            run('config1 TicTacToe, stock PolicyValueNet, 8192 games, mode %s' % mode, sp, 25 * 40, 25)
            del sp
            torch.cuda.empty_cache()
    if '2vl' in only:
        # config 2 again with leaf-parallel waves (opt-in, not the parity mode): K playouts per game and wave make
        # the launches K times longer, which amortises the fixed cost of a short convolution launch
        net2 = ResNetPolicyValueNet(6, n_blocks=6, board_width=7, n_actions=7).cuda().eval()
        for K in (1, 2, 4):
            sp = BatchedSelfPlay(4096, 6, 4, net=net2, n_playout=200, add_noise=True, seed=2, board_width=7,
                                 game_type=L.GAME_CONNECT4, leaves_per_tree=K)
            sp.set_random_start_positions(max_random_moves=3)
            sp.play(2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_moves = 10
            e0.record()
            for _ in range(n_moves * sp.waves_per_move):
                sp.step_wave()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            sp.forest.raise_faults()
            print(json.dumps({'config': 'config2 Connect Four, 4096 games, ResNet-6, leaves_per_tree=%d' % K,
                              'waves_per_move': sp.waves_per_move, 'ms_per_move': ms / n_moves,
                              'simulations_per_s': 4096 * 200 * n_moves / ms * 1e3}), flush=True)
            del sp
            torch.cuda.empty_cache()
    if '4g' in only:
        # the same board and trunk with five-in-a-row rules (the pre-Go stand-in of earlier runs)
        net4 = ResNetPolicyValueNet(19, n_blocks=20).cuda().eval()
        sp = BatchedSelfPlay(8192, 19, 5, net=net4, n_playout=800, add_noise=True, seed=3)
        run('config4-board 19x19 five-in-a-row, 800 sims/move, ResNet-20 bf16, 8192 games', sp, 160, 8)


if __name__ == '__main__':
    main()
