#!/usr/bin/env python
"""bench.py -- MCTS simulations/s of batched self-play, Gomoku 15x15 (BASELINE.json config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One *step* = one wave = one playout for every game on the GPU (G simulations): select ->
encode -> ResNet-10 forward (21 tcgen05 convolutions + heads) -> expand+backup, plus, every
``n_playout`` waves, the move commit (root policy, play, re-root, trajectory, refill).  The
default K = 800 waves is exactly one full move of 800 simulations for all G games.

N > 1: one process per GPU under torchrun, games sharded by global id, NO collective on the data
path (weak scaling: G games per GPU); the timed region is bracketed by a barrier and
torch.cuda.synchronize(), timed with CUDA events, MAX over ranks.

Besides the headline (``value`` / ``e2e`` / ``roofline`` / ``cpu_baseline``) the JSON line carries, all measured
OUTSIDE the timed region:

* ``tree_roofline``   -- rz_select_kernel / rz_expand_backup_kernel timed inside a captured CUDA graph;
* ``games_per_hour``  -- finished self-play games / device time over ``--gph-moves`` whole moves from staggered
                         start positions (a count, not an estimate), plus the sustained simulations/s of that run;
* ``configs``         -- short runs of BASELINE.json's other configurations and of the reference's stock network
                         (N = 1 only), each with its own roofline fraction and, where the oracle plays the game,
                         a bounded CPU sample;
* ``exchange``        -- N > 1: the off-path NCCL trajectory all-gather (compact device records) and weight
                         broadcast + re-pack, and ``shard_hash_ok``: every rank's visit counts for its global game
                         ids equal rank 0's single-GPU search of the same ids.

--impl reference: the reference's CPU algorithm (oracle port: one Python search per host core,
same ResNet-10 weights evaluated by PyTorch on the CPU), same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BOARD, K_ROW, N_PLAYOUT, BLOCKS, C_PUCT = 15, 5, 800, 10, 5.0
DEFAULT_GAMES = 8192
STAGGER = 96         # games/hour run: game g starts after (1000+g) mod STAGGER random moves (~94-ply games)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=N_PLAYOUT)
    ap.add_argument('--warmup', type=int, default=8)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--games', type=int, default=int(os.environ.get('RZ_BENCH_GAMES', DEFAULT_GAMES)))
    ap.add_argument('--playouts', type=int, default=N_PLAYOUT)
    ap.add_argument('--blocks', type=int, default=BLOCKS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-configs', action='store_true')
    ap.add_argument('--no-exchange', action='store_true')
    ap.add_argument('--gph-moves', type=int, default=int(os.environ.get('RZ_BENCH_GPH_MOVES', 10)),
                    help='whole moves of the games/hour run (0 = skip)')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                      '--format=csv,noheader,nounits'], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for s in self.samples:
            for name, val in zip(names, s[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------- CPU baseline
# spec of a CPU search: (game, H, W, k, net, blocks, n_playout).  net 'stock' = the reference's own PolicyValueNet
CPU_SPECS = {
    'config3': dict(game='gomoku', H=BOARD, W=BOARD, k=K_ROW, net='resnet', blocks=BLOCKS, n_playout=N_PLAYOUT),
    'stock15': dict(game='gomoku', H=BOARD, W=BOARD, k=K_ROW, net='stock', blocks=0, n_playout=N_PLAYOUT),
    'config1': dict(game='gomoku', H=3, W=3, k=3, net='stock', blocks=0, n_playout=25),
    'config2': dict(game='connect4', H=6, W=7, k=4, net='resnet', blocks=6, n_playout=200),
    'config4': dict(game='go', H=19, W=19, k=1, net='resnet', blocks=20, n_playout=800),
}


class _GoSearchEnv(object):
    """oracle.go_oracle.GoEnvOracle behind the env duck-type the AlphaZero search drives."""

    def __init__(self, n):
        from oracle import go_oracle
        self.e = go_oracle.GoEnvOracle(n, 7.5)
        self.e.reset()
        self.n = n

    def step(self, a):
        self.e.step(int(a))

    def leagel_actions(self):
        return [int(a) for a in self.e.legal_actions()]

    def current_player(self):
        return self.e.current_player()

    def game_end_winner(self):
        if not self.e.is_terminal():
            return False, -1
        return True, (0 if self.e.returns()[0] == 1 else 1)

    def current_state(self):
        import numpy as np
        obs = self.e.observe(self.e.agent_selection)['observation']
        return np.ascontiguousarray(obs.transpose(2, 0, 1)).astype(np.float32)


def _cpu_worker(args):
    """One host core: the oracle port of AlphaZeroMCTS with the network evaluated on the CPU."""
    idx, spec, seconds = args
    import copy
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import pyoracle
    from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet, ResNetPolicyValueNet
    torch.manual_seed(0)
    H, W, game = spec['H'], spec['W'], spec['game']
    if spec['net'] == 'stock':
        net = PolicyValueNet(H).eval()
        planes = 4
    elif game == 'connect4':
        net = ResNetPolicyValueNet(H, n_blocks=spec['blocks'], board_width=W, n_actions=W).eval()
        planes = 4
    elif game == 'go':
        net = ResNetPolicyValueNet(H, n_blocks=spec['blocks'], n_actions=H * W + 1, in_planes=17).eval()
        planes = 17
    else:
        net = ResNetPolicyValueNet(H, n_blocks=spec['blocks']).eval()
        planes = 4

    def pvf(env):  # alphazero_agent.py:31-46 on the CPU
        legal = env.leagel_actions()
        x = torch.from_numpy(np.ascontiguousarray(env.current_state().reshape(-1, planes, H, W))).float()
        with torch.no_grad():
            logp, v = net(x)
        probs = np.exp(logp.numpy().flatten())
        return zip(legal, probs[legal]), v.item()

    if game == 'connect4':
        board = pyoracle.ConnectFourBoard()
        board.reset()
    elif game == 'go':
        board = _GoSearchEnv(H)
    else:
        board = pyoracle.Board(H, spec['k'])
        board.reset()
        if H == BOARD:      # the bench start positions
            rs = np.random.RandomState(1000 + idx)
            for m in rs.permutation(H * W)[:(1000 + idx) % 31]:
                board.step(int(m))
                if board.game_end_winner()[0]:
                    board.reset()
                    break
    n_playout = spec['n_playout']
    pyoracle.Search(pvf, n_playout, C_PUCT, add_noise=True).playout(copy.deepcopy(board))  # warm-up (lazy imports)
    t0 = time.time()
    n = 0
    while time.time() - t0 < seconds:
        # one move's search of at most n_playout playouts (alphazero_mcts.py:83-85), then a fresh tree, until the
        # time sample is used up
        s = pyoracle.Search(pvf, n_playout, C_PUCT, add_noise=True)
        k = 0
        while k < n_playout and time.time() - t0 < seconds:
            s.playout(copy.deepcopy(board))
            k += 1
            n += 1
    return n, time.time() - t0


def cpu_baseline(which, seconds, blocks=None, n_playout=None, processes=None):
    import multiprocessing as mp
    spec = dict(CPU_SPECS[which])
    if blocks is not None and spec['net'] == 'resnet':
        spec['blocks'] = blocks
    if n_playout is not None:
        spec['n_playout'] = n_playout
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    if processes is not None:
        cores = max(1, min(cores, int(processes)))
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(i, spec, seconds) for i in range(cores)])
    sims = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    netname = 'the stock PolicyValueNet' if spec['net'] == 'stock' else 'ResNet-%d' % spec['blocks']
    rules = {'gomoku': 'rlzero/mcts + GomokuEnv', 'connect4': 'rlzero/mcts + a Connect Four env',
             'go': 'rlzero/mcts + GoEnv rules (MiniGo restatement)'}[spec['game']]
    return {'value': sims / wall, 'unit': 'simulations/s', 'cores': cores, 'kind': 'port',
            'sample': '%d host processes x up to %.0f s of AlphaZeroMCTS playouts (oracle port of %s, %s fp32 on '
                      'the CPU, batch 1, noise on), %dx%d; %d playouts in %.1f s' % (
                          cores, seconds, rules, netname, spec['H'], spec['W'], sims, wall)}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    base = cpu_baseline('config3', max(10.0, min(120.0, args.cpu_seconds * 2)), args.blocks, args.playouts)
    line = {'impl': 'reference', 'metric': 'mcts_simulations_per_sec', 'value': base['value'],
            'unit': 'simulations/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * base['cores'] / base['value'] if base['value'] else None,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': bench_config(args, base['cores']), 'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': 'simulations/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}
    line['config']['note'] = ('reference arm: one step = one playout on each of the %d host processes'
                              % base['cores'])
    print(json.dumps(line))


def bench_config(args, games_per_unit):
    return {'workload': 'Gomoku 15x15 (k=5) AlphaZero self-play, %d simulations/move, %d parallel games per '
                        'GPU, random-init ResNet-%d/128ch bf16, UCB1 rule of the reference, c_puct=5, '
                        'Dirichlet noise on, T=1, random start positions (SURVEY 8d), finished games '
                        'refilled' % (args.playouts, args.games, args.blocks),
            'games_per_gpu': args.games, 'simulations_per_move': args.playouts,
            'step': 'one wave = one simulation for every game (+ move commit every %d waves)' % args.playouts,
            'l2': 'working set per wave (activations 2x%.0f MB + node pools) far exceeds the 126 MB L2; '
                  'no explicit flush' % (args.games * 256 * 128 * 2 / 1e6),
            'parallelism': 'games sharded over %d GPU(s), no data-path collective' % args.gpus}


def conv_share_in_step(sp, n_waves=4):
    """Share of a wave spent in the trunk convolutions, measured INSIDE the step: CUPTI timeline (torch.profiler) of a few
    graph-replayed waves; a kernel's time is end - max(start, end of its predecessor), because with programmatic
    dependent launch a kernel is resident (and waiting) before its predecessor ends.  None if the profiler is
    unavailable."""
    try:
        import torch
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(n_waves):
                sp.step_wave()
            torch.cuda.synchronize()
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name
              and 'Memcpy' not in e.name and 'Memset' not in e.name]
        ev.sort(key=lambda e: e.time_range.start)
        if len(ev) < 8:
            return None
        conv = total = 0.0
        prev_end = ev[0].time_range.start
        for e in ev:
            st, en = e.time_range.start, e.time_range.end
            eff = max(0.0, en - max(st, prev_end))
            total += eff + max(0.0, st - prev_end)          # gaps count towards the wave
            if 'conv3x3' in e.name:
                conv += eff
            prev_end = max(prev_end, en)
        return conv / total if total > 0 else None
    except Exception:
        return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


# ------------------------------------------------------------------- tree kernels, graph-timed
def tree_roofline(sp, ms_per_step, peaks, n_rep=20):
    """rz_select_kernel and rz_expand_backup_kernel on mid-search trees, each timed as ``n_rep`` launches inside ONE
    captured CUDA graph (no host launch gap between them), CUDA events around the replay.  select only writes the
    wave scratch, so it repeats on the live trees; expand+backup mutates them, so it is timed as the difference
    between a graph of n_rep x (select, expand_backup) and the graph of n_rep x select (it re-uses the last
    network outputs: the kernel's work does not depend on their values)."""
    import torch
    f = sp.forest
    G = f.G
    hbm = peaks.get('hbm_gbs') or 6650.0
    src = 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650 (of fallback)'
    prior_is_log = bool(getattr(sp.evaluator, 'prior_is_log', False))

    def capture(body):
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        from rlzero_b200.engine import capture_graph
        with capture_graph(g):
            for _ in range(n_rep):
                body()
        return g

    def timed(g, reps=10):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            t0.record()
            g.replay()
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / n_rep
            best = ms if best is None else min(best, ms)
        return best

    g_sel = capture(f.select)
    sel_ms = timed(g_sel)
    depth = f.depth.clamp(min=0).double()
    levels = float(depth.sum().item())
    full_levels = float((depth - 1).clamp(min=0).sum().item())
    # algorithmic bytes: one int32 visit-count sweep (4*AS B) per level descended + the fp64 value sweep (8*AS B) on
    # levels whose children are all visited (every level but the leaf's parent: a lower bound) + root and leaf
    # boards + the path record
    sel_bytes = levels * 4 * f.AS + full_levels * 8 * f.AS + G * (2 * 2 * f.H * 4 + 2 * 32) + levels * 8
    out = {'bound': 'hbm', 'kernel': 'rz_select_kernel (%d trees, mean depth %.2f)' % (G, levels / G),
           'achieved': sel_bytes / sel_ms / 1e6, 'peak': hbm, 'unit': 'GB/s', 'frac': sel_bytes / sel_ms / 1e6 / hbm,
           'launch_ms': sel_ms, 'algorithmic_bytes_per_launch': sel_bytes, 'share_of_step': sel_ms / ms_per_step,
           'peak_source': src,
           'how': '%d launches captured in one CUDA graph, replayed, CUDA events around the replay, best of 10 '
                  '(compare profiles/*wave_launches.csv: ncu times the same kernel alone)' % n_rep}

    def pair():
        f.select()
        f.expand_backup(prior_is_log, sp.noise_eps, sp.noise_alpha, sp.seed)
    g_pair = capture(pair)
    pair_ms = timed(g_pair, reps=1)          # every replay grows the trees by n_rep nodes: one timed replay
    eb_ms = max(pair_ms - sel_ms, 1e-6)
    eb_bytes = G * (4 * f.AS + 4 * f.AS * (2 if f.store_priors else 1) + 2 * f.H * 4) + (levels + G) * 24
    out['expand_backup'] = {
        'kernel': 'rz_expand_backup_kernel (%d trees, Dirichlet noise %s)' % (G, 'on' if sp.noise_eps > 0 else 'off'),
        'achieved': eb_bytes / eb_ms / 1e6, 'peak': hbm, 'unit': 'GB/s', 'frac': eb_bytes / eb_ms / 1e6 / hbm,
        'launch_ms': eb_ms, 'algorithmic_bytes_per_launch': eb_bytes, 'share_of_step': eb_ms / ms_per_step,
        'how': 'graph of %d x (select, expand_backup) minus graph of %d x select, per launch' % (n_rep, n_rep),
        'note': 'latency- and ALU-bound (one warp per tree: %d Gamma draws per expansion), not bandwidth-bound' % f.A}
    sp.waves_in_move = 0
    return out


# ------------------------------------------------------------------- games per hour, measured
def games_per_hour(sp, args, world, dist):
    """Finished self-play games per hour, COUNTED: the games restart from staggered positions (game g after
    (1000+g) mod STAGGER uniformly random moves, so that episodes end from the first move on, as in the steady
    state of continuous refill), ``--gph-moves`` whole moves are played (n_playout waves + commit each), and the
    trajectory store's own counters give the finished games and their plies."""
    import torch
    sp.set_random_start_positions(max_random_moves=STAGGER)
    sp.forest.reset_trees()
    sp.warm_up()
    torch.cuda.synchronize()
    s0 = sp.stats()
    m0 = sp.moves_played
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    waves = 0
    e0.record()
    while sp.moves_played - m0 < args.gph_moves:
        sp.step_wave()
        waves += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    s1 = sp.stats()
    sp.forest.raise_faults()
    t = torch.tensor([ms, float(s1['games_done'] - s0['games_done']), float(s1['plies_done'] - s0['plies_done']),
                      float(waves)], dtype=torch.float64, device='cuda')
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(mx[0].item())
    games, plies, waves_all = float(t[1].item()), float(t[2].item()), float(t[3].item())
    return {'value': games / (ms / 1e3) * 3600.0, 'unit': 'finished games/hour (all %d GPU(s))' % world,
            'games_finished': int(games), 'moves_played': args.gph_moves, 'seconds': ms / 1e3,
            'mean_plies_searched_per_finished_game': plies / games if games else None,
            'start': 'game g starts after (1000+g) mod %d uniformly random moves (mean prefix %.1f plies, not '
                     'counted in mean_plies); finished games refilled from the empty board' % (STAGGER, (STAGGER - 1) / 2),
            'sustained_simulations_per_s': sp.G * waves_all / (ms / 1e3),
            'sustained_note': '%d waves incl. %d move commits per GPU, CUDA events, max over ranks' % (
                waves_all / world, args.gph_moves)}


# ------------------------------------------------------------------- the other configurations (N = 1)
def _timed_waves(sp, waves, warm):
    import torch
    sp.warm_up()
    sp.step_waves(max(warm, 8))        # (also captures the eight-wave graph small batches are replayed with)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sp.step_waves(waves)
    e1.record()
    torch.cuda.synchronize()
    sp.forest.raise_faults()
    return e0.elapsed_time(e1) / waves


def config_block(args, peaks):
    """Short runs of BASELINE.json's configs 1, 2, 4, 5 and of the reference's own network at config 3's size.  Per
    entry: ms_per_wave, simulations/s, the network FLOPs per simulation and roofline.frac = achieved network
    TFLOP/s / the measured SUSTAINED bf16 peak (these are kernels timed inside a step), and a bounded CPU sample of
    the same search where the oracle plays the game."""
    import torch
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet, ResNetPolicyValueNet
    from rlzero_b200.selfplay import BatchedSelfPlay
    sustained = peaks.get('bf16_tflops_sustained') or 1400.0
    src = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else 'fallback 1400 (of fallback)'
    cpu_s = 0.0 if args.no_cpu_baseline else min(5.0, args.cpu_seconds)
    out = []

    def stock_flops(H):
        hw = H * H
        return 2 * hw * 9 * (4 * 32 + 32 * 64 + 64 * 128) + 2 * hw * 128 * 6 + 2 * 4 * hw * hw + 2 * 2 * hw * 64 + 128

    def entry(name, sp, flops, waves, warm, random_moves, cpu=None, note=None, dtype='bf16'):
        sp.set_random_start_positions(max_random_moves=random_moves)
        ms = _timed_waves(sp, waves, warm)
        sims = sp.G / ms * 1e3            # parity mode: one playout per game and wave
        tf = flops * sims / 1e12
        e = {'config': name, 'games': sp.G, 'simulations_per_move': sp.n_playout, 'waves_timed': waves,
             'ms_per_wave': ms, 'simulations_per_s': sims, 'net_flops_per_simulation': flops, 'dtype': dtype,
             'kernels_per_wave': sp.kernels_per_wave(),
             'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': sustained, 'unit': 'TFLOP/s',
                          'frac': tf / sustained, 'peak_source': src}}
        if note:
            e['note'] = note
        if cpu and cpu_s > 0:
            e['cpu_baseline'] = cpu_baseline(cpu, cpu_s)
        out.append(e)
        return e

    torch.manual_seed(0)
    # config 1: TicTacToe = GomokuEnv(3, 3), 25 simulations/move, the stock PolicyValueNet on its default
    # path (mode 'tc32': tensor cores at float32-level accuracy, 16-stride layout)
    net1 = PolicyValueNet(3).cuda().eval()
    for G in (1, 8192):
        sp = BatchedSelfPlay(G, 3, 3, net=net1, n_playout=25, add_noise=True, seed=1)
        entry('config1: TicTacToe 3x3 k=3, 25 sims/move, stock PolicyValueNet (mode %s), %d game(s)' % (
            sp.evaluator.mode, G), sp, stock_flops(3), 25 * 20, 25, 3, cpu='config1' if G == 1 else None, dtype='bf16x3',
              note='one game = one board per launch: launch-latency bound, the tensor roofline does not apply'
              if G == 1 else 'a 3x3 board occupies 9 of the 256 rows of its 16-stride tile: the issued MMAs are 28x the '
                             'algorithmic FLOPs counted here')
        del sp
    # config 2: Connect Four 6x7, 200 simulations/move, 4096 games, ResNet-6 bf16
    net2 = ResNetPolicyValueNet(6, n_blocks=6, board_width=7, n_actions=7).cuda().eval()
    sp = BatchedSelfPlay(4096, 6, 4, net=net2, n_playout=200, add_noise=True, seed=2, board_width=7,
                         game_type=L.GAME_CONNECT4)
    entry('config2: Connect Four 6x7, 200 sims/move, 4096 games, ResNet-6 bf16', sp, net2.flops_per_eval(), 200 * 4,
          200, 3, cpu='config2')
    del sp
    torch.cuda.empty_cache()
    # config 4: Go 19x19 (GoEnv rules, komi 7.5), 800 simulations/move, ResNet-20 bf16, 17 planes, 362 actions
    net4 = ResNetPolicyValueNet(19, n_blocks=20, n_actions=362, in_planes=17).cuda().eval()
    sp = BatchedSelfPlay(8192, 19, 1, net=net4, n_playout=800, add_noise=True, seed=3, game_type=L.GAME_GO,
                         komi=7.5, max_moves=722)
    entry('config4: Go 19x19 (komi 7.5), 800 sims/move, 8192 games, ResNet-20 bf16', sp, net4.flops_per_eval(), 40, 4,
          31, cpu='config4')
    del sp, net4
    torch.cuda.empty_cache()
    # the reference's OWN network at config 3's size: fp32 default path and the tensor-core path
    net = PolicyValueNet(15).cuda().eval()
    for mode, waves in (('tc32', 200), ('f32', 16), ('tc', 200)):
        sp = BatchedSelfPlay(8192, 15, 5, net=net, n_playout=800, add_noise=True, seed=1, net_mode=mode)
        entry('stock PolicyValueNet 15x15, 800 sims/move, 8192 games, mode %s%s' % (
            mode, ' (the default: float32-level accuracy on the tensor cores)' if mode == 'tc32' else ''), sp,
              stock_flops(15), waves, 4, 31, dtype={'f32': 'f32', 'tc': 'bf16', 'tc32': 'bf16x3'}[mode],
              note='the CPU sample of this search is cpu_baseline.stock_net of the headline')
        del sp
        torch.cuda.empty_cache()
    # small batches of the headline configuration (the sequential single-game search the reference's training script
    # runs, alphazero_mcts.py:73-94, and 64 games side by side): stem, one-launch trunk, cluster heads per wave
    net3 = ResNetPolicyValueNet(15, n_blocks=10).cuda().eval()
    for G in (1, 64):
        sp = BatchedSelfPlay(G, 15, 5, net=net3, n_playout=800, add_noise=True, seed=1)
        entry('small batch: Gomoku 15x15, 800 sims/move, ResNet-10 bf16, %d game(s)' % G, sp, net3.flops_per_eval(), 800,
              50, 31, note='latency bound: one CTA pair per board keeps the activation in shared memory across the 20 '
                           'trunk layers (rz_net_trunk_small.cu); per-layer kernels took 250 us per wave for one game')
        del sp
    del net3
    torch.cuda.empty_cache()
    # config 5: MuZero on Gomoku 15x15, 50 latent simulations/move, 8192 games (no reference code: parity unpinned)
    from rlzero_b200.muzero import BatchedMuZeroSelfPlay, MuZeroConfig, MuZeroNet
    net5 = MuZeroNet(15, repr_blocks=10, dyn_blocks=5).cuda().eval()
    G = 8192
    mz = BatchedMuZeroSelfPlay(G, 15, 5, net=net5, config=MuZeroConfig(num_simulations=50), seed=4)
    for _ in range(2):
        mz.play_move()
    torch.cuda.synchronize()
    n_moves = 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_moves):
        mz.play_move()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    mz.search.raise_faults()
    flops = G * n_moves * (net5.flops('initial') + 50 * net5.flops('recurrent'))
    tf = flops / ms / 1e9
    out.append({'config': 'config5: MuZero Gomoku 15x15, 50 latent sims/move, 8192 games, h=ResNet-10 g=ResNet-5 bf16',
                'games': G, 'simulations_per_move': 50, 'moves_timed': n_moves, 'ms_per_move': ms / n_moves,
                'ms_per_wave': ms / n_moves / 50, 'simulations_per_s': G * 50 * n_moves / ms * 1e3,
                'net_flops_per_simulation': net5.flops('recurrent'), 'dtype': 'bf16',
                'kernels_per_move': mz.search.kernels_per_move(),
                'roofline': {'bound': 'tensor', 'achieved': tf, 'peak': sustained, 'unit': 'TFLOP/s',
                             'frac': tf / sustained, 'peak_source': src},
                'note': 'no reference MuZero exists (parity unpinned): no CPU sample'})
    del mz, net5
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------- the training step (N = 1)
def train_block(peaks):
    """AlphaZeroAgent.learn (alphazero_agent.py:59-86) on synthetic batches: the hand-written step (tensor-core trunk for
    ResNet-10, float32 CUDA-core kernels for the reference's own network) next to the same step through PyTorch autograd
    (cuDNN / cuBLAS, PyTorch's default TF32 convolutions) on the same GPU.  FLOPs per step = 3 x forward (forward, data
    gradient, weight gradient); frac = achieved / the measured sustained bf16 peak."""
    import numpy as np
    import torch
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet, ResNetPolicyValueNet
    sustained = peaks.get('bf16_tflops_sustained') or 1400.0
    out = []
    rs = np.random.RandomState(0)

    def stock_flops(H):
        hw = H * H
        return 2 * hw * 9 * (4 * 32 + 32 * 64 + 64 * 128) + 2 * hw * 128 * 6 + 2 * 4 * hw * hw + 2 * 2 * hw * 64 + 128

    for name, H, make, B, flops, steps in (
            ('ResNet-10/128ch 15x15, batch 4096', 15, lambda: ResNetPolicyValueNet(15, n_blocks=10), 4096, None, 6),
            ('stock PolicyValueNet 15x15, batch 512', 15, lambda: PolicyValueNet(15), 512, stock_flops(15), 10),
            ('stock PolicyValueNet 6x6, batch 32 (the sizes of tools/train_alphazero.py:19-33)', 6,
             lambda: PolicyValueNet(6), 32, stock_flops(6), 30)):
        x = torch.from_numpy((rs.rand(B, 4, H, H) < 0.2).astype(np.float32)).cuda()
        pi = torch.from_numpy(rs.dirichlet(0.3 * np.ones(H * H), size=B).astype(np.float32)).cuda()
        z = torch.from_numpy(rs.choice([-1.0, 0.0, 1.0], size=B).astype(np.float32)).cuda()
        entry = {'step': 'AlphaZeroAgent.learn, ' + name}
        kinds = ('native', 'native_tc', 'autograd', 'autograd_fp32_strict') if 'stock' in name else (
            'native', 'autograd', 'autograd_fp32_strict')
        for kind in kinds:
            torch.manual_seed(0)
            net = make()
            strict = kind == 'autograd_fp32_strict'
            torch.backends.cudnn.allow_tf32 = not strict
            torch.backends.cuda.matmul.allow_tf32 = False
            agent = AlphaZeroAgent(H, net=net, trainer=kind if kind.startswith('native') else 'autograd')
            if flops is None:
                flops = net.flops_per_eval()
            for _ in range(2):
                agent.learn(x, pi, z)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss, ent = agent.learn(x, pi, z)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            tf = 3.0 * flops * B / ms / 1e9
            entry[kind] = {'ms_per_step': ms, 'steps_per_s': 1e3 / ms, 'samples_per_s': B * 1e3 / ms, 'tflops': tf,
                           'frac_of_sustained_bf16_peak': tf / sustained, 'loss_after': loss,
                           'path': ('hand-written kernels (%s%s)' % (type(agent.trainer).__name__, ', whole trunk on the tensor '
                                    'cores: gradients within 1e-2 of the float64 oracle instead of 1e-5' if kind == 'native_tc'
                                    else '')) if agent.trainer is not None
                           else ('PyTorch autograd, cuDNN/cuBLAS, true fp32 convolutions (cudnn.allow_tf32 = False): the '
                                 'accuracy class of the float32 native step' if strict else
                                 'PyTorch autograd, cuDNN/cuBLAS, fp32 tensors with TF32 convolutions (PyTorch default)')}
            del agent, net
            torch.cuda.empty_cache()
        torch.backends.cudnn.allow_tf32 = True
        # the strongest stock-PyTorch variant of the same step: bf16 autocast, channels_last, fused Adam
        import torch.nn.functional as F
        torch.manual_seed(0)
        net = make().cuda().to(memory_format=torch.channels_last)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
        if flops is None:
            flops = net.flops_per_eval()
        xc = x.contiguous(memory_format=torch.channels_last)

        def ac_step():
            net.train()
            with torch.autocast('cuda', dtype=torch.bfloat16):
                lp, v = net(xc)
            loss = F.mse_loss(v.float().view(-1), z) - torch.mean(torch.sum(pi * lp.float(), dim=1))
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        for _ in range(2):
            ac_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = ac_step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        entry['autograd_bf16_autocast'] = {'ms_per_step': ms, 'steps_per_s': 1e3 / ms, 'tflops': 3.0 * flops * B / ms / 1e9,
                                           'loss_after': float(loss.item()),
                                           'path': 'PyTorch autograd, torch.autocast(bf16), channels_last, fused Adam; no '
                                                   'inference re-pack (not part of that path)'}
        del net, opt
        torch.cuda.empty_cache()
        entry['speedup_vs_autograd'] = entry['autograd']['ms_per_step'] / entry['native']['ms_per_step']
        entry['speedup_vs_autograd_bf16_autocast'] = entry['autograd_bf16_autocast']['ms_per_step'] / entry['native']['ms_per_step']
        entry['speedup_vs_autograd_fp32_strict'] = entry['autograd_fp32_strict']['ms_per_step'] / entry['native']['ms_per_step']
        if 'native_tc' in entry:
            entry['native_tc_speedup_vs_autograd'] = entry['autograd']['ms_per_step'] / entry['native_tc']['ms_per_step']
        entry['includes'] = 'forward, loss, backward, Adam, and the re-pack of the inference weights (refresh_weights)'
        out.append(entry)
    return out


# ------------------------------------------------------------------- off-path exchange (N > 1)
def exchange_block(sp, net, world, rank, dist):
    """The two collectives of a generation, off the search path, on NCCL: all-gather of compact trajectory records
    that never leave HBM, and the weight broadcast + re-pack (+ wave-graph re-capture).  And shard invariance
    across REAL ranks: every rank searches 64 small games with its own global ids; rank 0 also searches all
    world x 64 of them on its one GPU; the per-block hashes of the visit tensors must agree."""
    import torch
    from rlzero_b200 import parallel
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    from rlzero_b200.selfplay import BatchedSelfPlay
    f = sp.forest
    dev = torch.device('cuda', torch.cuda.current_device())
    # -- trajectory all-gather: 32768 records per rank in the ring's own format (contents: whatever the ring holds)
    n_rec = 32768
    cap = f.ring_capacity
    idx = torch.arange(n_rec, device=dev) % cap
    rows, info, pi = f.traj['ring_rows'][idx], f.traj['ring_info'][idx], f.traj['ring_pi'][idx]
    rec_bytes = rows[0].numel() * 4 + info[0].numel() * 4 + pi[0].numel() * 4
    parallel.gather_records_device(rows[:64], info[:64], pi[:64], global_offset=rank * f.G)      # warm NCCL
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ag_best = 1e30
    for _ in range(3):          # best of 3: the first exchange of a new size also pays NCCL's buffer set-up
        dist.barrier()
        e0.record()
        g_rows, g_info, g_pi, counts = parallel.gather_records_device(rows, info, pi, global_offset=rank * f.G)
        e1.record()
        torch.cuda.synchronize()
        ag_best = min(ag_best, e0.elapsed_time(e1))
    ok_gather = (g_info.shape[0] == n_rec * world and
                 bool(torch.equal(g_pi[rank * n_rec:(rank + 1) * n_rec], pi)))
    t = torch.tensor([ag_best], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ag_ms = float(t.item())
    recv_bytes = rec_bytes * n_rec * (world - 1)       # what every GPU receives over NVLink
    # the collective alone on a larger pre-packed payload (262144 records per rank): NVLink bandwidth, no pack / count
    big = parallel.pack_records(rows, info, pi).repeat(8, 1)
    recv = torch.empty(world * big.shape[0], big.shape[1], dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, big)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(5):
        dist.all_gather_into_tensor(recv, big)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    coll_ms = float(t.item())
    coll_bytes = big.numel() * 4 * (world - 1)
    del recv, big
    # -- weight broadcast + re-pack + graph re-capture
    dist.barrier()
    torch.cuda.synchronize()
    bc_ms = 1e30
    for _ in range(3):          # best of 3, as above
        dist.barrier()
        e0.record()
        moved = parallel.broadcast_weights(net, src=0)
        e1.record()
        torch.cuda.synchronize()
        bc_ms = min(bc_ms, e0.elapsed_time(e1))
    t1 = time.perf_counter()
    sp.evaluator.refresh_weights()
    sp.step_wave()                                       # re-captures the wave graph with the new weights
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    t = torch.tensor([bc_ms, (t2 - t1) * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # -- shard invariance across ranks
    g_small, n_po = 64, 64
    torch.manual_seed(0)
    small_net = ResNetPolicyValueNet(6, n_blocks=2).cuda().eval()

    def block_hashes(n_games, offset):
        s = BatchedSelfPlay(n_games, 6, 4, net=small_net, n_playout=n_po + 1, add_noise=True, global_offset=offset,
                            seed=77, temperature=1.0)      # n_po + 1: the move is not committed inside this run
        s.set_random_start_positions(max_random_moves=5)
        s.warm_up()
        for _ in range(n_po - 1):
            s.step_wave()
        s.forest.root_policy(1.0, seed=77)
        v = s.forest.visits.long()
        w = (torch.arange(v.shape[1], device=dev, dtype=torch.int64) * 2654435761 % 2147483647 + 1)[None, :]
        gw = (torch.arange(n_games, device=dev, dtype=torch.int64) + offset + 1)[:, None] * 40503 % 1000003 + 1
        h = ((v + 1) * w * gw % 2305843009213693951).reshape(n_games // g_small, -1).sum(dim=1) % 2305843009213693951
        s.forest.raise_faults()
        return h
    mine = block_hashes(g_small, rank * g_small)
    every = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(every, mine)
    shard_ok = None
    if rank == 0:
        single = block_hashes(g_small * world, 0)
        shard_ok = bool(torch.equal(single, every))
    return {'trajectory_all_gather': {'records_per_rank': n_rec, 'bytes_per_record': rec_bytes,
                                      'bytes_received_per_gpu': recv_bytes, 'ms': ag_ms,
                                      'gb_per_s_per_gpu': recv_bytes / ag_ms / 1e6,
                                      'vs_nvlink_770_gb_s_measured_peer_copy': recv_bytes / ag_ms / 1e6 / 770.0,
                                      'payload_ok': ok_gather,
                                      'collective_only': {'records_per_rank': 8 * n_rec, 'bytes_received_per_gpu': coll_bytes,
                                                          'ms': coll_ms, 'gb_per_s_per_gpu': coll_bytes / coll_ms / 1e6,
                                                          'vs_nvlink_770_gb_s_measured_peer_copy':
                                                              coll_bytes / coll_ms / 1e6 / 770.0},
                                      'how': 'pack + NCCL all_gather_into_tensor + unpack of device tensors, CUDA '
                                             'events, max over ranks (includes the 8-byte count collective)'},
            'weight_broadcast': {'bytes': moved, 'ms': float(t[0].item()),
                                 'gb_per_s': moved / float(t[0].item()) / 1e6,
                                 'repack_and_graph_recapture_ms': float(t[1].item())},
            'backend': dist.get_backend(), 'world': world}, shard_ok


# -------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    from rlzero_b200.selfplay import BatchedSelfPlay

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    G, P = args.games, args.playouts
    torch.manual_seed(0)
    net = ResNetPolicyValueNet(BOARD, n_blocks=args.blocks).cuda().eval()
    sp = BatchedSelfPlay(G, BOARD, K_ROW, net=net, n_playout=P, c_puct=C_PUCT, temperature=1.0,
                         add_noise=True, global_offset=rank * G, seed=1234)
    sp.set_random_start_positions()
    peaks = load_peaks()

    def time_conv(reps=20):
        """One 128->128 trunk layer (the dominant kernel) alone on the launching stream, CUDA events."""
        ev = sp.evaluator
        lib = L.load()
        layer = ev.layers[1]
        x, y = ev.bufs[0], ev.bufs[1]
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3 + reps):
            if i == 3:
                k0.record()
            L.check(lib.rz_net_conv3x3_tc2(L.ptr(x), L.ptr(layer['w']), L.ptr(layer['b']), None, L.ptr(y), G,
                                           BOARD, BOARD, 128, 1, 2, getattr(ev, 'conv_flags', 0), 0, L.stream_ptr()))
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / reps

    # kernel-alone (burst) figure: taken BEFORE the sustained run heats the part into its power cap
    conv_ms_cold = time_conv() if rank == 0 else None
    sp.warm_up()
    for _ in range(max(3, args.warmup) - 1):
        sp.step_wave()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    commits0 = sp.moves_played
    barrier()
    e0.record()
    for _ in range(args.steps):
        sp.step_wave()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if sampler:
        sampler.stop_flag = True
    commits = sp.moves_played - commits0
    sp.forest.raise_faults()
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sims = float(G) * args.steps * world
    value = sims / (ms / 1e3)

    # dominant kernel: one 128->128 conv layer of the trunk, timed alone on the same stream
    ev = sp.evaluator
    roof = None
    tree_roof = None
    if rank == 0:
        conv_ms_hot = time_conv()         # right after the sustained run (power-capped clocks)
        conv_ms = conv_ms_cold
        flops = 2.0 * G * BOARD * BOARD * 128 * 128 * 9      # algorithmic: 225 squares x 128 x 1152 MACs
        n_conv = len(ev.layers) - 1
        in_step_est = min(1.0, n_conv * conv_ms_hot / (ms / args.steps))
        in_step = conv_share_in_step(sp)
        share_how = ('CUPTI timeline of 4 graph-replayed waves right after the timed region: sum over the conv launches of '
                     'end - max(start, end of the previous kernel), divided by the wave time')
        if in_step is None:
            in_step, share_how = in_step_est, 'launches_per_step x launch_ms_after_sustained_run / ms_per_step (upper bound)'
        peak = peaks.get('bf16_tflops') or 1590.0          # burst figure: the kernel is timed alone
        sustained = peaks.get('bf16_tflops_sustained') or 1400.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'conv3x3_tc_traffic.json')))['dram_bytes_per_launch']
        except Exception:
            pass
        roof = {'bound': 'tensor', 'kernel': 'rz_conv3x3_tc2_kernel<2> (128->128, %d boards, flags %d)' % (G, getattr(ev, 'conv_flags', 0)),
                'achieved': flops / conv_ms / 1e9, 'peak': peak, 'unit': 'TFLOP/s',
                'frac': flops / conv_ms / 1e9 / peak,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst, of measured)' if peaks else 'fallback 1590 (of fallback)',
                'launch_ms': conv_ms, 'launch_ms_after_sustained_run': conv_ms_hot,
                'frac_after_sustained_run_of_sustained_peak': flops / conv_ms_hot / 1e9 / sustained,
                'launches_per_step': n_conv, 'share_of_step': in_step, 'share_of_step_how': share_how,
                'share_of_step_upper_bound': in_step_est,
                'traffic': traffic,
                'issued_tflops': flops * 256.0 / 225.0 / conv_ms / 1e9,
                'issued_frac_of_nominal_2250': flops * 256.0 / 225.0 / conv_ms / 1e9 / 2250.0,
                'note': ('algorithmic FLOPs count the 225 squares of a board; the kernel issues MMAs for the 256 rows '
                         'of its padded 16x16 tile.  frac may exceed 1: the MEASURED_PEAKS burst figure is cuBLAS '
                         'at its own power-limited clock (about 0.76 of the nominal 2250 TFLOP/s)'),
                'net_forward_tflops_in_step': net.flops_per_eval() * G / (ms / args.steps) / 1e9,
                'frac_in_step_of_sustained': net.flops_per_eval() * G / (ms / args.steps) / 1e9 / sustained}
        for _ in range(min(args.playouts // 2, 400)):   # mid-search trees (the timed region ended on a commit)
            sp.step_wave()
        tree_roof = tree_roofline(sp, ms / args.steps, peaks)

    # end to end through the public API with host buffers (per move: H2D positions, D2H pi/moves)
    e2e = None
    if not args.no_e2e:
        rows, meta = sp.forest.boards()
        meta = meta.copy()
        meta[:, L.META_STATUS] = L.ACTIVE
        n_moves = max(1, args.steps // P)
        sp.get_actions(rows, meta) if os.environ.get('RZ_E2E_WARM') else None
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_moves):
            sp.get_actions(rows, meta)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d, d2h = sp.api_bytes()
        e2e = {'value': float(G) * P * n_moves * world / float(tt.item()), 'unit': 'simulations/s',
               'h2d_bytes_per_step': h2d / P, 'd2h_bytes_per_step': d2h / P,
               'api': 'BatchedSelfPlay.get_actions(host positions) -> host (moves, pi, visits); '
                      '%d move(s) of %d playouts' % (n_moves, P)}

    exchange, shard_ok = None, None
    if world > 1 and not args.no_exchange:
        exchange, shard_ok = exchange_block(sp, net, world, rank, dist)

    gph = games_per_hour(sp, args, world, dist) if args.gph_moves > 0 else None

    if rank == 0:
        kpw = sp.kernels_per_wave()
        line = {'metric': 'mcts_simulations_per_sec', 'value': value, 'unit': 'simulations/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                'config': bench_config(args, G),
                'clocks': sampler.summary() if sampler else None,
                'gpu_launches': args.steps * kpw + commits * 2,
                'e2e': e2e, 'roofline': roof, 'tree_roofline': tree_roof, 'games_per_hour': gph}
        if exchange is not None:
            line['exchange'] = exchange
            line['shard_hash_ok'] = shard_ok
        hbm_main = sp.forest.hbm_bytes()
        del sp
        torch.cuda.empty_cache()
        if world == 1 and not args.no_configs:
            line['configs'] = config_block(args, peaks)
            line['train'] = train_block(peaks)
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline('config3', args.cpu_seconds, args.blocks, P)
            # the same CPU search with the reference's own stock network as evaluator, a shorter sample
            stock = cpu_baseline('stock15', min(8.0, args.cpu_seconds), n_playout=P)
            line['cpu_baseline']['stock_net'] = {'value': stock['value'], 'unit': stock['unit'], 'cores': stock['cores'],
                                                 'sample': stock['sample']}
            # ONE process, the way the reference's own training script runs the search
            one = cpu_baseline('stock15', min(4.0, args.cpu_seconds), n_playout=P, processes=1)
            line['cpu_baseline']['stock_net_single_process'] = {'value': one['value'], 'unit': one['unit'], 'cores': 1,
                                                                'sample': one['sample']}
        line['hbm_bytes_node_pools_and_scratch'] = hbm_main
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
