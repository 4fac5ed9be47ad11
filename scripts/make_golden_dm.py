#!/usr/bin/env python
"""Golden vectors of the reference's DeepMindMCTS (rlzero/mcts/deepmind_mcts.py) -> tests/golden/dm_mcts.json.

Run in the authoring container (needs /root/reference).  The UNMODIFIED reference class searches an
adapter subclass of the reference GomokuEnv (legal_actions() gets a default argument,
deepmind_mcts.py:497 calls it bare), with the closed-form evaluators of oracle/evaluators.py and its
private RandomState replaced by oracle.dm_oracle.NoShuffle (no child shuffle; seeded root noise)."""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dm_oracle, ref_loader  # noqa: E402

CASES = [
    # size, k, moves, sims, method, solve, eval_id, noise_seed (None = no noise)
    (3, 3, [], 60, 'puct', True, 2, None),
    (3, 3, [0, 3, 1, 4], 80, 'puct', True, 2, None),
    (3, 3, [4, 0], 120, 'uct', True, 2, None),
    (3, 3, [0, 3, 1, 4], 80, 'uct', False, 1, None),
    (3, 3, [1, 0, 4, 3], 200, 'puct', True, 2, None),
    (4, 3, [5, 0, 6], 150, 'puct', True, 2, None),
    (4, 3, [], 120, 'puct', True, 2, 11),
    (5, 4, [12, 7, 13, 8, 11], 300, 'uct', True, 2, None),
    (6, 4, [14, 15, 20], 200, 'puct', False, 2, None),
    (6, 4, [14, 15, 20, 21, 8], 400, 'puct', True, 2, 5),
    (8, 5, [27, 28, 35, 36, 19], 500, 'puct', True, 2, None),
    (15, 5, [112, 113, 97], 300, 'puct', True, 2, None),
    # nearly full boards: every line ends quickly, so the solver proves the root and the search stops early
    (3, 3, [0, 1, 2, 4, 3, 5, 7], 50, 'puct', True, 2, None),
    (3, 3, [0, 1, 2, 4, 3, 5], 200, 'puct', True, 2, None),
    (3, 3, [4, 0, 8, 2], 400, 'uct', True, 2, None),
    (3, 3, [4, 0, 8, 2, 1], 400, 'puct', True, 2, None),
    (4, 3, [0, 4, 1, 5, 8, 2, 9, 7, 12, 10, 3], 600, 'puct', True, 2, None),
    # larger boards and budgets (added at the end of round 1)
    (9, 5, [40, 41, 31, 32, 49], 400, 'uct', False, 2, None),
    (9, 5, [40, 41, 31, 32, 49, 22], 600, 'puct', True, 2, 3),
    (15, 5, [112, 113, 97, 98, 127], 500, 'uct', True, 2, None),
    (15, 5, [], 400, 'puct', False, 1, None),
    (7, 4, [24, 25, 17, 18, 31], 500, 'puct', True, 2, None),
    (6, 4, [14, 15, 20, 21, 8, 9], 300, 'uct', True, 2, 7),
]


def main():
    ref = ref_loader.load()
    sys.path.insert(0, ref_loader.REFERENCE_ROOT + '/rlzero')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from rlzero.mcts import deepmind_mcts as dm

    class Adapter(ref.GomokuEnv):
        def legal_actions(self, player=None):
            return list(self.leagel_actions())

    out = []
    for size, k, moves, sims, method, solve, eval_id, noise_seed in CASES:
        env = Adapter(size, k)
        env.reset()
        for a in moves:
            env.step(a)
        ev = dm_oracle.ClosedFormEvaluator(eval_id)
        bot = dm.DeepMindMCTS(env, uct_c=2, max_simulations=sims, evaluator=ev, child_selection_method=method,
                              add_exploration_noise=noise_seed is not None, dirichlet_noise_alpha=1.0,
                              dirichlet_noise_epsilon=0.25, solve=solve)
        bot._random_state = dm_oracle.NoShuffle(None if noise_seed is None else np.random.RandomState(noise_seed))
        with contextlib.redirect_stdout(io.StringIO()):
            root = bot.mcts_search(env)
            policy, action = None, root.best_child().action
        out.append(dict(size=size, k=k, moves=moves, sims=sims, method=method, solve=solve, eval_id=eval_id,
                        noise_seed=noise_seed, root_n=root.explore_count, root_w=root.total_reward,
                        root_outcome=root.outcome, best=int(action),
                        children=[[int(c.action), int(c.explore_count), float(c.total_reward), c.outcome,
                                   float(c.prior)] for c in root.children]))
        print(size, k, moves, method, solve, '-> N', root.explore_count, 'outcome', root.outcome, 'best', action)
    with open(os.path.join(ROOT, 'tests', 'golden', 'dm_mcts.json'), 'w') as f:
        json.dump(out, f)


if __name__ == '__main__':
    main()
