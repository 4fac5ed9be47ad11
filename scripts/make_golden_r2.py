#!/usr/bin/env python
"""Round-2 fixtures from the LIVE reference (authoring container only; needs /root/reference):

* tests/golden/mcts_noise.json      -- AlphaZeroMCTS with ``add_noise=True`` and a seeded global numpy stream
  (node.py:63-69), UCB1 and the PUCT rule (test-side formula patch, oracle/ref_loader.use_puct_rule): the
  root's visits / value sums / float64 priors, the number and a sha1 of the Dirichlet samples drawn (the legacy
  numpy stream is frozen, so a test re-draws them from the seed), and a tree-reuse chain.  Pins the seeded-noise parity of the host-callback path and of the ``noise64`` input of
  ``rz_tree_expand_backup_ex``.
* tests/golden/dm_mcts_shuffle.json -- the UNMODIFIED DeepMindMCTS with its private RandomState seeded: real child
  shuffle (deepmind_mcts.py:508) and root noise drawn from the same stream.

Every number is produced by the reference's own classes; this repository contributes the closed-form evaluators
only.
"""
import contextlib
import hashlib
import io
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dm_oracle, ref_loader  # noqa: E402
from oracle.evaluators import EVAL_HASH, EVAL_KAT, make_policy_value_fn  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def fhex(x):
    return float(x).hex()


NOISE_CASES = [
    # name, size, k, playouts, c, eval, rule, seed, pre_moves, chain
    ('N1_3x3_puct', 3, 3, 120, 2.0, EVAL_HASH, 'puct', 7, [4], [0]),
    ('N2_6x6_puct', 6, 4, 300, 5, EVAL_HASH, 'puct', 11, [14, 15], [20]),
    ('N3_6x6_uct', 6, 4, 200, 5, EVAL_HASH, 'uct', 3, [], [14]),
    ('N4_8x8_puct_kat', 8, 5, 400, 3, EVAL_KAT, 'puct', 5, [27, 28], []),
    ('N5_15x15_puct', 15, 5, 500, 5, EVAL_HASH, 'puct', 2, [112, 113, 97], [127]),
]


def dump_root(root, n_actions):
    visits, w, prior, legal = [0] * n_actions, [fhex(0.0)] * n_actions, [fhex(0.0)] * n_actions, [0] * n_actions
    for a, ch in root._children.items():
        visits[a], w[a], prior[a], legal[a] = int(ch.explore_count), fhex(ch.total_reward), fhex(ch.prior), 1
    return {'root_N': int(root.explore_count), 'root_W': fhex(root.total_reward), 'visits': visits, 'W': w,
            'prior': prior, 'has_child': legal}


def noise_case(ref, name, size, k, n_playout, c, eval_id, rule, seed, pre, chain):
    env = ref.GomokuEnv(board_size=size, n_in_row=k)
    env.reset()
    for m in pre:
        env.step(m)
    draws = []
    real = np.random.dirichlet

    def recording(alpha, *a, **kw):        # record what the reference drew, in order
        out = real(alpha, *a, **kw)
        draws.append([fhex(x) for x in out])
        return out
    cm = ref_loader.use_puct_rule(ref) if rule == 'puct' else contextlib.nullcontext()
    stages = []
    with cm:
        np.random.seed(seed)
        np.random.dirichlet = recording
        try:
            s = ref.AlphaZeroMCTS(make_policy_value_fn(eval_id), n_playout=n_playout, c_puct=c, add_noise=True)
            acts, probs = s.simulate(env, 1.0)
            st = dump_root(s._root, size * size)
            st['acts'] = [int(a) for a in acts]
            st['probs'] = [fhex(p) for p in probs]
            st['n_draws'] = len(draws)
            stages.append(st)
            for move in chain:
                env.step(move)
                s.update_with_move(move)
                s.simulate(env, 1.0)
                st = dump_root(s._root, size * size)
                st['n_draws'] = len(draws)
                stages.append(st)
        finally:
            np.random.dirichlet = real
    return dict(name=name, size=size, k=k, n_playout=n_playout, c_puct=c, eval_id=eval_id, rule=rule, seed=seed,
                pre=list(pre), chain=list(chain), stages=stages, n_draws=len(draws),
                draws_sha1=hashlib.sha1(json.dumps(draws).encode()).hexdigest(), first_draw=draws[0])


DM_CASES = [
    # size, k, moves, sims, method, solve, eval_id, add_noise, seed
    (3, 3, [], 60, 'puct', True, 2, False, 1),
    (3, 3, [4, 0], 120, 'uct', True, 2, False, 2),
    (3, 3, [0, 3, 1, 4], 80, 'uct', False, 1, False, 3),
    (4, 3, [], 120, 'puct', True, 2, True, 11),
    (5, 4, [12, 7, 13, 8, 11], 300, 'uct', True, 2, False, 4),
    (6, 4, [14, 15, 20, 21, 8], 400, 'puct', True, 2, True, 5),
    (6, 4, [14, 15, 20], 200, 'uct', False, 1, False, 6),
    (8, 5, [27, 28, 35, 36, 19], 500, 'puct', True, 1, False, 7),
    (9, 5, [40, 41, 31, 32, 49, 22], 300, 'puct', True, 2, True, 3),
    (15, 5, [112, 113, 97], 300, 'uct', True, 1, False, 8),
]


def dm_cases(ref):
    sys.path.insert(0, ref_loader.REFERENCE_ROOT + '/rlzero')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from rlzero.mcts import deepmind_mcts as dm

    class Adapter(ref.GomokuEnv):
        def legal_actions(self, player=None):
            return list(self.leagel_actions())

    out = []
    for size, k, moves, sims, method, solve, eval_id, noise, seed in DM_CASES:
        env = Adapter(size, k)
        env.reset()
        for a in moves:
            env.step(a)
        ev = dm_oracle.ClosedFormEvaluator(eval_id)
        bot = dm.DeepMindMCTS(env, uct_c=2, max_simulations=sims, evaluator=ev, child_selection_method=method,
                              add_exploration_noise=noise, dirichlet_noise_alpha=1.0, dirichlet_noise_epsilon=0.25,
                              solve=solve)
        bot._random_state = np.random.RandomState(seed)      # the reference's own attribute, seeded
        with contextlib.redirect_stdout(io.StringIO()):
            root = bot.mcts_search(env)
        out.append(dict(size=size, k=k, moves=moves, sims=sims, method=method, solve=solve, eval_id=eval_id,
                        noise=noise, seed=seed, root_n=root.explore_count, root_w=fhex(root.total_reward),
                        root_outcome=root.outcome, best=int(root.best_child().action),
                        children=[[int(c.action), int(c.explore_count), fhex(c.total_reward), c.outcome,
                                   fhex(c.prior)] for c in root.children]))
        print('dm', size, k, moves, method, 'seed', seed, '-> N', root.explore_count, 'best', out[-1]['best'],
              'order', [c[0] for c in out[-1]['children']][:6])
    return out


def main():
    ref = ref_loader.load()
    cases = [noise_case(ref, *c) for c in NOISE_CASES]
    for c in cases:
        print(c['name'], 'draws', c['n_draws'], 'visits', c['stages'][0]['visits'][:12])
    json.dump({'generator': 'scripts/make_golden_r2.py', 'cases': cases}, open(os.path.join(OUT, 'mcts_noise.json'), 'w'))
    json.dump({'generator': 'scripts/make_golden_r2.py', 'cases': dm_cases(ref)},
              open(os.path.join(OUT, 'dm_mcts_shuffle.json'), 'w'))


if __name__ == '__main__':
    main()
