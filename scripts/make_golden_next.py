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


def main():
    ref = ref_loader.load()
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
