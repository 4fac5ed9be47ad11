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
ASSUMED_PLIES = 94   # SURVEY.md section 6 probe of the reference: 94-ply self-play game at 15x15


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
def _cpu_worker(args):
    """One host core: the oracle port of AlphaZeroMCTS with the ResNet evaluated on the CPU."""
    idx, blocks, n_playout, seconds, seed_moves, state = args
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import pyoracle
    from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet, ResNetPolicyValueNet
    torch.manual_seed(0)
    # blocks < 0: the reference's own stock PolicyValueNet (what tools/train_alphazero.py runs) instead of the ResNet
    net = (PolicyValueNet(BOARD) if blocks < 0 else ResNetPolicyValueNet(BOARD, n_blocks=blocks)).eval()
    if state is not None:
        net.load_state_dict(state)

    def pvf(env):  # alphazero_agent.py:31-46 on the CPU
        legal = env.leagel_actions()
        x = torch.from_numpy(np.ascontiguousarray(env.current_state().reshape(-1, 4, BOARD, BOARD))).float()
        with torch.no_grad():
            logp, v = net(x)
        probs = np.exp(logp.numpy().flatten())
        return zip(legal, probs[legal]), v.item()

    board = pyoracle.Board(BOARD, K_ROW)
    board.reset()
    rs = np.random.RandomState(1000 + idx)
    for m in rs.permutation(BOARD * BOARD)[:(1000 + idx) % 31]:
        board.step(int(m))
        if board.game_end_winner()[0]:
            board.reset()
            break
    import copy
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


def cpu_baseline(blocks, n_playout, seconds):
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(i, blocks, n_playout, seconds, None, None) for i in range(cores)])
    sims = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    netname = 'the stock PolicyValueNet' if blocks < 0 else 'ResNet-%d' % blocks
    return {'value': sims / wall, 'unit': 'simulations/s', 'cores': cores, 'kind': 'port',
            'sample': '%d host processes x up to %.0f s of AlphaZeroMCTS playouts (oracle port of '
                      'rlzero/mcts + GomokuEnv, %s fp32 on the CPU, batch 1, noise on), '
                      '15x15 from the bench start positions; %d playouts in %.1f s' % (
                          cores, seconds, netname, sims, wall)}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    base = cpu_baseline(args.blocks, args.playouts, max(10.0, min(120.0, args.cpu_seconds * 2)))
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
    lib = L.load()
    roof = None
    tree_roof = None
    if rank == 0:
        conv_ms_hot = time_conv()         # right after the sustained run (power-capped clocks)
        conv_ms = conv_ms_cold
        flops = 2.0 * G * BOARD * BOARD * 128 * 128 * 9      # algorithmic: 225 squares x 128 x 1152 MACs
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        n_conv = len(ev.layers) - 1
        in_step = min(1.0, n_conv * conv_ms_hot / (ms / args.steps))
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
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst, of measured)' if peaks else 'fallback 1590',
                'launch_ms': conv_ms, 'launch_ms_after_sustained_run': conv_ms_hot,
                'frac_after_sustained_run_of_sustained_peak': flops / conv_ms_hot / 1e9 / sustained,
                'launches_per_step': n_conv, 'share_of_step': in_step,
                'traffic': traffic,
                'issued_tflops': flops * 256.0 / 225.0 / conv_ms / 1e9,
                'issued_frac_of_nominal_2250': flops * 256.0 / 225.0 / conv_ms / 1e9 / 2250.0,
                'note': ('algorithmic FLOPs count the 225 squares of a board; the kernel issues MMAs for the 256 rows '
                         'of its padded 16x16 tile.  frac may exceed 1: the MEASURED_PEAKS burst figure is cuBLAS '
                         'at its own power-limited clock (about 0.76 of the nominal 2250 TFLOP/s)'),
                'net_forward_tflops_in_step': net.flops_per_eval() * G / (ms / args.steps) / 1e9,
                'frac_in_step_of_sustained': net.flops_per_eval() * G / (ms / args.steps) / 1e9 / sustained}

        # tree side (HBM-bound kernels): select is idempotent (it only writes the wave scratch), so it
        # can be timed back to back on the live trees; algorithmic bytes = one int32 visit-count
        # sweep (4*AS B) per level descended + the fp64 value sweep (8*AS B) on levels whose children
        # are all visited (counted as every level but the leaf's parent: a lower bound) + root and
        # leaf boards + the path record
        f = sp.forest
        for _ in range(min(args.playouts // 2, 400)):   # mid-search trees (the timed region ended on a commit)
            sp.step_wave()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f.select()
        torch.cuda.synchronize()
        t0.record()
        for _ in range(20):
            f.select()
        t1.record()
        torch.cuda.synchronize()
        sel_ms = t0.elapsed_time(t1) / 20
        depth = f.depth.clamp(min=0).double()
        levels = float(depth.sum().item())
        full_levels = float((depth - 1).clamp(min=0).sum().item())
        sel_bytes = levels * 4 * f.AS + full_levels * 8 * f.AS + G * (2 * 2 * BOARD * 4 + 2 * 32) + levels * 8
        hbm = peaks.get('hbm_gbs') or 6650.0
        tree_roof = {'bound': 'hbm', 'kernel': 'rz_select_kernel (%d trees, mean depth %.2f)' % (G, levels / G),
                     'achieved': sel_bytes / sel_ms / 1e6, 'peak': hbm, 'unit': 'GB/s',
                     'frac': sel_bytes / sel_ms / 1e6 / hbm, 'launch_ms': sel_ms,
                     'algorithmic_bytes_per_launch': sel_bytes,
                     'share_of_step': sel_ms / (ms / args.steps),
                     'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650'}
        # expand + backup mutates the trees, so it is timed inside live (un-graphed) waves, CUDA events around the
        # launch.  Algorithmic bytes per tree: the leaf's log-priors (4*AS read), the new edge block's visit
        # counts (4*AS written; + 4*AS priors when they are stored), the leaf occupancy rows, and a read-modify-write
        # of (N int32, W fp64) per edge of the path plus the root
        eb_ms, n_eb = 0.0, 10
        for _ in range(n_eb):
            f.select()
            sp.evaluator(f)
            t0.record()
            f.expand_backup(bool(getattr(sp.evaluator, 'prior_is_log', False)), sp.noise_eps, sp.noise_alpha, sp.seed)
            t1.record()
            torch.cuda.synchronize()
            eb_ms += t0.elapsed_time(t1) / n_eb
        sp.waves_in_move += n_eb
        eb_bytes = G * (4 * f.AS + 4 * f.AS * (2 if f.store_priors else 1) + 2 * BOARD * 4) + (levels + G) * 24
        tree_roof['expand_backup'] = {
            'kernel': 'rz_expand_backup_kernel (%d trees, Dirichlet noise %s)' % (G, 'on' if sp.noise_eps > 0 else 'off'),
            'achieved': eb_bytes / eb_ms / 1e6, 'peak': hbm, 'unit': 'GB/s', 'frac': eb_bytes / eb_ms / 1e6 / hbm,
            'launch_ms': eb_ms, 'algorithmic_bytes_per_launch': eb_bytes, 'share_of_step': eb_ms / (ms / args.steps),
            'note': 'latency- and ALU-bound (one warp per tree: %d Gamma draws per expansion), not bandwidth-bound' % f.A}

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

    if rank == 0:
        kpw = sp.kernels_per_wave()
        line = {'metric': 'mcts_simulations_per_sec', 'value': value, 'unit': 'simulations/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                'config': bench_config(args, G),
                'games_per_hour_est': value / (P * ASSUMED_PLIES) * 3600.0,
                'games_per_hour_note': 'simulations/s / (%d sims/move x %d plies/game, the reference '
                                       'probe in SURVEY 6)' % (P, ASSUMED_PLIES),
                'clocks': sampler.summary() if sampler else None,
                'gpu_launches': args.steps * kpw + commits * 2,
                'e2e': e2e, 'roofline': roof, 'tree_roofline': tree_roof}
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(args.blocks, P, args.cpu_seconds)
            # the same CPU search with the reference's own stock network as evaluator, a shorter sample
            stock = cpu_baseline(-1, P, min(8.0, args.cpu_seconds))
            line['cpu_baseline']['stock_net'] = {'value': stock['value'], 'unit': stock['unit'], 'cores': stock['cores'],
                                                 'sample': stock['sample']}
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
