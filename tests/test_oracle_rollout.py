"""The oracle's RolloutMCTS restatement against the live reference (same global np.random
stream => same playouts => same visit counts and moves)."""
import copy

import numpy as np
import pytest

from oracle import pyoracle, ref_loader

pytestmark = pytest.mark.reference


@pytest.mark.parametrize('size,k,n_playout,seed', [(3, 3, 40, 0), (5, 4, 60, 1), (6, 4, 80, 2)])
def test_rollout_search_matches_reference(size, k, n_playout, seed):
    ref = ref_loader.load()
    from rlzero.mcts.rollout_mcts import RolloutMCTS, RolloutPlayer  # noqa: E402  (path set by the loader)
    env = ref.GomokuEnv(board_size=size, n_in_row=k)
    env.reset()
    board = pyoracle.Board(size, k)
    board.reset()
    rs = np.random.RandomState(seed)
    for m in rs.permutation(size * size)[:size]:
        env.step(int(m))
        board.step(int(m))
        if env.game_end_winner()[0]:
            return
    np.random.seed(100 + seed)
    a = RolloutMCTS(n_playout=n_playout, c_puct=5)
    move_ref = a.simulate(copy.deepcopy(env))
    visits_ref = np.zeros(size * size, dtype=np.int32)
    for act, node in a._root._children.items():
        visits_ref[act] = node.explore_count
    np.random.seed(100 + seed)
    b = pyoracle.RolloutSearch(n_playout=n_playout, c_puct=5)
    move = b.simulate(copy.deepcopy(board))
    assert move == move_ref
    assert np.array_equal(b.root_visits(size * size), visits_ref)
    # players: same move sequence in a short match against themselves
    np.random.seed(7)
    p_ref = RolloutPlayer(n_playout=20)
    np.random.seed(7)
    p = pyoracle.RolloutSearchPlayer(n_playout=20)
    e2, b2 = copy.deepcopy(env), copy.deepcopy(board)
    np.random.seed(9)
    seq_ref = []
    for _ in range(4):
        if e2.game_end_winner()[0]:
            break
        m = p_ref.get_action(e2)
        seq_ref.append(m)
        e2.step(m)
    np.random.seed(9)
    seq = []
    for _ in range(4):
        if b2.game_end_winner()[0]:
            break
        m = p.get_action(b2)
        seq.append(m)
        b2.step(m)
    assert seq == seq_ref
