#!/usr/bin/env python
"""Generate the fixtures of the "next" rows from the LIVE reference: pure-MCTS
(rollout) searches with deterministic rollout policies and TrainPipeline.get_equi_data.

Run in the authoring container:  python scripts/make_golden_next.py
Writes tests/golden/rollout.json and tests/golden/equi.json.
"""
import copy
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def rollout_case(ref, size, k, n_playout, pre, mode):
    from rlzero.mcts.rollout_mcts import RolloutMCTS
    env = ref.GomokuEnv(board_size=size, n_in_row=k)
    env.reset()
    for m in pre:
        env.step(m)
    s = RolloutMCTS(n_playout=n_playout, c_puct=5)

    def policy(game_env, mode=mode):      # deterministic stand-in for np.random.rand (:96-100)
        legal = game_env.leagel_actions()
        probs = -np.arange(len(legal), dtype=np.float64) if mode == 'first' else np.arange(len(legal), dtype=np.float64)
        return zip(legal, probs)
    s.rollout_policy = policy
    move = s.simulate(copy.deepcopy(env))
    visits = [0] * (size * size)
    w = [0.0] * (size * size)
    for a, ch in s._root._children.items():
        visits[a] = int(ch.explore_count)
        w[a] = float(ch.total_reward)
    return dict(size=size, k=k, n_playout=n_playout, pre=list(pre), mode=mode, move=int(move), visits=visits,
                W=w, root_N=int(s._root.explore_count), root_W=float(s._root.total_reward))


def equi_case(size, seed):
    spec = importlib.util.spec_from_file_location('ref_train', os.path.join(ref_loader.REFERENCE_ROOT, 'tools',
                                                                           'train_alphazero.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rs = np.random.RandomState(seed)
    n = 3
    states = rs.randint(0, 2, size=(n, 4, size, size)).astype(np.float64)
    pis = rs.rand(n, size * size)
    zs = rs.choice([-1.0, 0.0, 1.0], size=n)
    fake_self = types.SimpleNamespace(board_size=size)
    out = mod.TrainPipeline.get_equi_data(fake_self, list(zip(states, pis, zs)))
    return dict(size=size, seed=seed, n=n,
                states=[o[0].astype(np.int8).reshape(-1).tolist() for o in out],
                pis=[[float(x).hex() for x in o[1]] for o in out], zs=[float(o[2]) for o in out])


def connect4_case(ref, n_playout, pre, eval_id, c_puct=5, chain=()):
    """The LIVE reference's AlphaZeroMCTS (duck-typed env) searching the oracle's
    Connect Four board with a closed-form evaluator; optional chain of (move, playouts) with tree
    reuse through update_with_move."""
    from oracle import pyoracle
    from oracle.evaluators import make_policy_value_fn
    b = pyoracle.ConnectFourBoard()
    b.reset()
    for m in pre:
        b.step(m)
    s = ref.AlphaZeroMCTS(make_policy_value_fn(eval_id), n_playout=n_playout, c_puct=c_puct)
    stages = []

    def dump():
        visits, w = [0] * 7, [0.0] * 7
        for a, ch in s._root._children.items():
            visits[a] = int(ch.explore_count)
            w[a] = float(ch.total_reward)
        stages.append(dict(visits=visits, W=w, root_N=int(s._root.explore_count), root_W=float(s._root.total_reward)))
    acts, probs = s.simulate(b, 1.0)
    dump()
    stages[-1]['acts'] = [int(a) for a in acts]
    for move, n in chain:
        b.step(move)
        s.update_with_move(move)
        s.n_playout = n
        if b.game_end_winner()[0]:
            break
        s.simulate(b, 1.0)
        dump()
    return dict(n_playout=n_playout, pre=list(pre), eval_id=eval_id, c_puct=c_puct, chain=[list(c) for c in chain],
                stages=stages)


def connect4_games(n, seed):
    """Random legal games on the oracle board: per ply the action, legal list, winner/end flags and
    the observation planes -- the fixture for the device rules kernels."""
    from oracle import pyoracle
    rs = np.random.RandomState(seed)
    games = []
    for _ in range(n):
        b = pyoracle.ConnectFourBoard()
        b.reset()
        plies = []
        while True:
            a = int(b.legal[rs.randint(len(b.legal))])
            _, reward, win, _ = b.step(a)
            end, winner = b.game_end_winner()
            plies.append(dict(a=a, reward=int(reward), win=bool(win), end=bool(end), winner=int(winner),
                              legal=list(b.legal), last=int(b.last_move),
                              planes=b.current_state().astype(np.int8).reshape(-1).tolist()))
            if end:
                break
        games.append(plies)
    return games


def main():
    ref = ref_loader.load()
    from oracle.evaluators import EVAL_HASH, EVAL_KAT
    c4 = [connect4_case(ref, 200, [], EVAL_KAT), connect4_case(ref, 200, [3, 3, 2], EVAL_HASH),
          connect4_case(ref, 150, [0, 0, 0, 0, 0, 0, 3], EVAL_HASH, chain=[(3, 150), (4, 150)]),
          connect4_case(ref, 300, [3, 4, 3, 4, 3], EVAL_HASH, c_puct=2.5, chain=[(2, 100)])]
    json.dump({'generator': 'scripts/make_golden_next.py', 'cases': c4, 'games': connect4_games(8, 5)},
              open(os.path.join(OUT, 'connect4.json'), 'w'))
    for c in c4:
        print('connect4', c['pre'], [st['visits'] for st in c['stages']])
    cases = [rollout_case(ref, 3, 3, 60, [], 'first'), rollout_case(ref, 3, 3, 60, [4, 0], 'last'),
             rollout_case(ref, 5, 4, 120, [12, 6], 'first'), rollout_case(ref, 6, 4, 200, [14, 15, 20], 'last'),
             rollout_case(ref, 6, 4, 150, [], 'first'), rollout_case(ref, 8, 5, 100, [27, 28, 35, 36], 'last')]
    json.dump({'generator': 'scripts/make_golden_next.py', 'cases': cases}, open(os.path.join(OUT, 'rollout.json'), 'w'))
    eq = [equi_case(6, 0), equi_case(3, 1), equi_case(15, 2)]
    json.dump({'generator': 'scripts/make_golden_next.py', 'cases': eq}, open(os.path.join(OUT, 'equi.json'), 'w'))
    for c in cases:
        print('rollout', c['size'], c['mode'], c['move'], c['root_N'], c['root_W'])


if __name__ == '__main__':
    main()
