#!/usr/bin/env python
"""Generate tests/golden/*.json from the LIVE reference (authoring container only).

Run:  python scripts/make_golden.py
Needs /root/reference (read-only).  The fixtures it writes travel with the repo
so that the oracle restatement and the CUDA path can be checked on the GPU box,
where the reference does not exist.  Every number below is produced by the
unmodified reference classes (oracle/ref_loader.py) -- nothing is computed by
code from this repository except the closed-form evaluators that are fed to it.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.evaluators import (EVAL_HASH, EVAL_KAT, EVAL_ZERO,  # noqa: E402
                               make_policy_value_fn)

OUT = os.path.join(ROOT, 'tests', 'golden')


def fhex(x):
    return float(x).hex()


def dump_root(root, n_actions):
    visits = [0] * n_actions
    w = [fhex(0.0)] * n_actions
    prior = [fhex(0.0)] * n_actions
    legal = [0] * n_actions
    for a, ch in root._children.items():
        visits[a] = int(ch.explore_count)
        w[a] = fhex(ch.total_reward)
        prior[a] = fhex(ch.prior)
        legal[a] = 1
    return {'root_N': int(root.explore_count), 'root_W': fhex(root.total_reward),
            'visits': visits, 'W': w, 'prior': prior, 'has_child': legal}


def count_nodes(root):
    total, expanded, depth = 0, 0, 0
    stack = [(root, 0)]
    while stack:
        node, d = stack.pop()
        total += 1
        depth = max(depth, d)
        if node._children:
            expanded += 1
            for ch in node._children.values():
                stack.append((ch, d + 1))
    return {'nodes': total, 'expanded': expanded, 'depth': depth}


MCTS_CASES = [
    # name, size, k, playouts, c, eval, rule, pre_moves, chain
    ('A_3x3_zero', 3, 3, 25, 5, EVAL_ZERO, 'uct', [], []),
    ('B_3x3_kat', 3, 3, 25, 5, EVAL_KAT, 'uct', [], []),
    ('C_3x3_kat_reuse', 3, 3, 25, 5, EVAL_KAT, 'uct', [], [4, 0]),
    ('D_6x6_kat', 6, 4, 400, 5, EVAL_KAT, 'uct', [], []),
    ('E_15x15_kat', 15, 5, 800, 5, EVAL_KAT, 'uct', [], []),
    ('F_3x3_terminal', 3, 3, 60, 5, EVAL_KAT, 'uct', [0, 3, 1, 4], []),
    ('G_3x3_hash_full', 3, 3, 200, 5, EVAL_HASH, 'uct', [], [4, 0, 8, 2]),
    ('H_6x6_hash_reuse', 6, 4, 300, 5, EVAL_HASH, 'uct', [14, 15], [20, 21, 9]),
    ('I_8x8_hash_c1', 8, 5, 500, 1.25, EVAL_HASH, 'uct', [27, 28, 35], [36]),
    ('J_15x15_hash_mid', 15, 5, 800, 5, EVAL_HASH, 'uct',
     [112, 113, 97, 98, 127, 128, 82, 83], [68]),
    ('K_3x3_tie', 3, 3, 100, 5, EVAL_HASH, 'uct', [0, 1, 2, 4, 3, 5, 7, 6], []),
    ('L_19x19_hash', 19, 5, 400, 5, EVAL_HASH, 'uct', [180, 181], [200]),
    ('P_6x6_puct', 6, 4, 300, 5, EVAL_HASH, 'puct', [], [14, 15]),
    ('Q_3x3_puct', 3, 3, 120, 2.0, EVAL_HASH, 'puct', [4], [0, 8]),
    ('R_15x15_puct', 15, 5, 800, 5, EVAL_HASH, 'puct', [112, 113], [127]),
    ('S_8x8_puct_kat', 8, 5, 600, 3, EVAL_KAT, 'puct', [], [27]),
]


def run_mcts_case(ref, case):
    name, size, k, n_playout, c, eval_id, rule, pre_moves, chain = case
    env = ref.GomokuEnv(board_size=size, n_in_row=k)
    env.reset()
    for m in pre_moves:
        env.step(m)
    mcts = ref.AlphaZeroMCTS(make_policy_value_fn(eval_id), n_playout=n_playout,
                             c_puct=c, add_noise=False)
    stages = []

    def one():
        acts, probs = mcts.simulate(env, 1.0)
        st = dump_root(mcts._root, size * size)
        st['acts'] = [int(a) for a in acts]
        st['probs_T1'] = [fhex(p) for p in probs]
        st['tree'] = count_nodes(mcts._root)
        stages.append(st)

    def go():
        one()
        for m in chain:
            env.step(m)
            mcts.update_with_move(m)
            one()

    if rule == 'puct':
        with ref_loader.use_puct_rule(ref):
            go()
    else:
        go()
    return {'name': name, 'board_size': size, 'n_in_row': k, 'n_playout': n_playout,
            'c_puct': c, 'eval_id': eval_id, 'rule': rule, 'pre_moves': pre_moves,
            'chain': chain, 'stages': stages}


def run_env_case(ref, size, k, seed):
    rs = np.random.RandomState(seed)
    env = ref.GomokuEnv(board_size=size, n_in_row=k)
    obs0 = env.reset()
    plies = []
    assert obs0.dtype == np.float64
    while True:
        legal = list(env.leagel_actions())
        a = int(legal[rs.randint(len(legal))])
        obs, reward, win, _ = env.step(a)
        end, winner = env.game_end_winner()
        plies.append({'a': a, 'reward': int(reward), 'win': bool(win), 'end': bool(end),
                      'winner': int(winner), 'player_after': int(env.current_player()),
                      'obs_sha1': hashlib.sha1(
                          np.ascontiguousarray(obs.astype(np.float32)).tobytes()).hexdigest(),
                      'n_legal': len(env.leagel_actions())})
        if end:
            break
    return {'board_size': size, 'n_in_row': k, 'seed': seed, 'plies': plies,
            'returns': [int(x) for x in env.returns()]}


def run_selfplay_case(ref, size, k, n_playout, eval_id, seed, temperature):
    """Full start_self_play episode with the global numpy RNG seeded; noise is
    on (is_selfplay=True) so the RNG stream includes one dirichlet per expansion."""
    np.random.seed(seed)
    env = ref.GomokuEnv(board_size=size, n_in_row=k)
    game = ref.GameControl(env)
    player = ref.AlphaZeroPlayer(make_policy_value_fn(eval_id), n_playout=n_playout,
                                 c_puct=5, is_selfplay=True)
    winner, data = game.start_self_play(player, temperature=temperature)
    data = list(data)
    moves = []
    # the move played at ply t is the single new stone between state t and t+1
    stones_prev = set()
    recs = []
    for state, pi, z in data:
        recs.append({'state_sha1': hashlib.sha1(
            np.ascontiguousarray(state.astype(np.float32)).tobytes()).hexdigest(),
            'pi': [fhex(x) for x in pi], 'z': float(z)})
    # recover the move list from the final env
    order = list(env.states.keys())  # dict preserves play order
    moves = [int(m) for m in order]
    del stones_prev
    return {'board_size': size, 'n_in_row': k, 'n_playout': n_playout, 'eval_id': eval_id,
            'seed': seed, 'temperature': temperature, 'winner': int(winner),
            'moves': moves, 'records': recs}


def run_net_case(ref, size, seed, batch):
    import torch
    torch.manual_seed(seed)
    net = ref.PolicyValueNet(size)
    net.eval()
    rs = np.random.RandomState(seed)
    x = (rs.rand(batch, 4, size, size) < 0.3).astype(np.float32)
    with torch.no_grad():
        logp, v = net(torch.from_numpy(x))
    sd = {k: v_.numpy() for k, v_ in net.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, 'net_pvn_%d.npz' % size), x=x,
                        logp=logp.numpy(), v=v.numpy(), **{'w_' + k: a for k, a in sd.items()})


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    mcts = [run_mcts_case(ref, c) for c in MCTS_CASES]
    with open(os.path.join(OUT, 'mcts_kat.json'), 'w') as f:
        json.dump({'generator': 'scripts/make_golden.py', 'cases': mcts}, f)
    envs = []
    for size, k in ((3, 3), (6, 4), (8, 5), (15, 5), (19, 5)):
        for seed in range(6):
            envs.append(run_env_case(ref, size, k, 100 * size + seed))
    with open(os.path.join(OUT, 'env_games.json'), 'w') as f:
        json.dump({'generator': 'scripts/make_golden.py', 'games': envs}, f)
    sp = [run_selfplay_case(ref, 3, 3, 25, EVAL_KAT, 7, 1.0),
          run_selfplay_case(ref, 3, 3, 40, EVAL_HASH, 11, 1e-3),
          run_selfplay_case(ref, 6, 4, 60, EVAL_HASH, 3, 1.0)]
    with open(os.path.join(OUT, 'selfplay.json'), 'w') as f:
        json.dump({'generator': 'scripts/make_golden.py', 'episodes': sp}, f)
    run_net_case(ref, 3, 0, 5)
    run_net_case(ref, 6, 1, 4)
    for c in mcts:
        st = c['stages'][0]
        print(c['name'], st['root_N'], st['tree'],
              hashlib.sha1(np.array(st['visits'], dtype='<i4').tobytes()).hexdigest()[:12])


if __name__ == '__main__':
    main()
