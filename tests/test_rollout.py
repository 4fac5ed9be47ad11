"""Pure-MCTS (rollout) search: oracle restatement and device search against the fixtures the live
reference produced with deterministic rollout policies (scripts/make_golden_next.py)."""
import copy
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, 'golden', 'rollout.json')))['cases']


def _board(case):
    from oracle import pyoracle
    b = pyoracle.Board(case['size'], case['k'])
    b.reset()
    for m in case['pre']:
        b.step(m)
    return b


@pytest.mark.parametrize('case', CASES, ids=lambda c: '%dx%d_%s_%d' % (c['size'], c['size'], c['mode'], c['n_playout']))
def test_oracle_rollout_search_matches_golden(case):
    from oracle import pyoracle
    s = pyoracle.RolloutSearch(case['n_playout'], 5, rollout=case['mode'])
    move = s.simulate(_board(case))
    assert move == case['move']
    assert s.root_visits(case['size'] ** 2).tolist() == case['visits']
    assert s.root.n == case['root_N'] and s.root.w == case['root_W']


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES, ids=lambda c: '%dx%d_%s_%d' % (c['size'], c['size'], c['mode'], c['n_playout']))
def test_device_rollout_search_matches_golden(case):
    """RolloutMCTS.simulate on the device (select / rz_eval_rollout / expand+backup kernels) ==
    the reference's RolloutMCTS with the same deterministic rollout policy: visits, W, move."""
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.mcts import RolloutMCTS
    env = GomokuEnv(case['size'], case['k'])
    env.reset()
    for m in case['pre']:
        env.step(m)
    s = RolloutMCTS(n_playout=case['n_playout'], c_puct=5, rollout=case['mode'])
    move = s.simulate(env)
    visits, w, has, root_n, root_w = s._forest.root_stats()
    assert move == case['move']
    assert visits[0].tolist() == case['visits']
    assert w[0].tolist() == case['W']
    assert int(root_n[0]) == case['root_N'] and float(root_w[0]) == case['root_W']
    # the caller's env is untouched (alphazero_mcts.py:83 deep-copies)
    assert len(env.states) == len(case['pre'])


@pytest.mark.gpu
def test_device_random_rollouts_statistics_and_player():
    """Random playouts: seeded determinism, and the value distribution of the evaluator equals
    the oracle's (-1 for every decisive playout, 0 for ties: the reference's literal rule)."""
    import torch
    from oracle import pyoracle
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import RolloutEvaluator, SearchForest
    from rlzero_b200.games.gomoku import GameControl, GomokuEnv
    from rlzero_b200.mcts import RolloutPlayer
    G = 4096
    f = SearchForest(G, 3, 3, n_playout=1)
    ev = RolloutEvaluator(n_limit=1000, seed=5)
    f.select()
    ev(f)
    v = f.value.cpu().numpy()
    assert set(np.unique(v)).issubset({-1.0, 0.0})
    tie_rate = float((v == 0).mean())
    # oracle: random tic-tac-toe playouts from the empty board
    rng = np.random.RandomState(0)
    ties = 0
    n = 3000
    for _ in range(n):
        b = pyoracle.Board(3, 3)
        b.reset()
        s = pyoracle.RolloutSearch(1, 5, rng=rng)
        ties += s.evaluate(b) == 0
    assert abs(tie_rate - ties / n) < 0.03, (tie_rate, ties / n)      # ~12.7 % of random games are drawn
    f2 = SearchForest(G, 3, 3, n_playout=1)
    f2.select()
    RolloutEvaluator(n_limit=1000, seed=5)(f2)
    assert torch.equal(f.value, f2.value)
    # n_limit = 0: no playout, the position is not over => value 0
    RolloutEvaluator(n_limit=0, seed=5)(f2)
    assert float(f2.value.abs().max()) == 0.0
    # RolloutPlayer through GameControl.start_play (tools/train_alphazero.py:139-163 uses it like this)
    np.random.seed(3)
    env = GomokuEnv(5, 4)
    game = GameControl(env)
    p0 = RolloutPlayer(n_playout=30, seed=1)
    p1 = RolloutPlayer(n_playout=30, seed=2)
    winner = game.start_play(p0, p1, start_player=0, is_shown=0)
    assert winner in (-1, 0, 1)
