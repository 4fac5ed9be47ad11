"""CPU tests (authoring container only): restatement vs the LIVE reference on
randomised inputs.  Skipped where /root/reference does not exist."""
import numpy as np
import pytest

from oracle import pyoracle, ref_loader
from oracle.evaluators import EVAL_HASH, EVAL_KAT, make_policy_value_fn

pytestmark = pytest.mark.reference


@pytest.fixture(scope='module')
def ref():
    return ref_loader.load()


@pytest.mark.parametrize('size,k,n_playout,seed', [
    (3, 3, 40, 0), (3, 3, 90, 1), (5, 4, 150, 2), (6, 4, 220, 3), (8, 5, 300, 4),
    (9, 5, 250, 5), (15, 5, 500, 6)])
@pytest.mark.parametrize('rule', ['uct', 'puct'])
def test_random_positions(ref, size, k, n_playout, seed, rule):
    rs = np.random.RandomState(seed)
    env = ref.GomokuEnv(size, k)
    env.reset()
    board = pyoracle.Board(size, k)
    board.reset()
    for _ in range(rs.randint(0, max(1, size * size // 3))):
        legal = env.leagel_actions()
        a = int(legal[rs.randint(len(legal))])
        env.step(a)
        if env.game_end_winner()[0]:
            return
        board.step(a)
    fn = make_policy_value_fn(EVAL_HASH if seed % 2 else EVAL_KAT)
    c = float(rs.choice([0.5, 1.25, 5.0]))
    m_ref = ref.AlphaZeroMCTS(fn, n_playout=n_playout, c_puct=c)
    m_or = pyoracle.Search(fn, n_playout, c, rule=pyoracle.RULE_PUCT if rule == 'puct' else 0)

    def both():
        a1, p1 = m_ref.simulate(env, 1.0)
        a2, p2 = m_or.simulate(board, 1.0)
        assert tuple(a1) == tuple(a2)
        assert [float(x).hex() for x in p1] == [float(x).hex() for x in p2]
        for a in a1:
            r, o = m_ref._root._children[a], m_or.root.children[a]
            assert (r.explore_count, float(r.total_reward).hex()) == (o.n, float(o.w).hex())

    def run():
        both()
        for _ in range(2):
            legal = env.leagel_actions()
            a = int(legal[rs.randint(len(legal))])
            env.step(a)
            board.step(a)
            if env.game_end_winner()[0]:
                return
            m_ref.update_with_move(a)
            m_or.update_with_move(a)
            both()

    if rule == 'puct':
        with ref_loader.use_puct_rule(ref):
            run()
    else:
        run()


def test_env_observation_and_quirks(ref):
    rs = np.random.RandomState(5)
    for size, k in ((3, 3), (7, 4), (15, 5)):
        env, b = ref.GomokuEnv(size, k), pyoracle.Board(size, k)
        assert np.array_equal(env.reset(), b.reset())
        while True:
            legal = env.leagel_actions()
            a = int(legal[rs.randint(len(legal))])
            o1, r1, w1, _ = env.step(a)
            o2, r2, w2, _ = b.step(a)
            assert np.array_equal(o1, o2) and (r1, w1) == (r2, w2)
            assert env.returns() == b.returns()
            assert env.game_end_winner() == b.game_end_winner()
            if env.game_end_winner()[0]:
                break
    with pytest.raises(AssertionError):
        env.step(a)
    with pytest.raises(AssertionError):
        b.step(a)


# ----------------------------------------------------------------- DeepMindMCTS (second search driver)
def _dm_reference():
    import contextlib
    import io
    import sys
    ref = ref_loader.load()
    if ref_loader.REFERENCE_ROOT + '/rlzero' not in sys.path:
        sys.path.insert(0, ref_loader.REFERENCE_ROOT + '/rlzero')   # deepmind_mcts.py:9 `from games.base_env ...`
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from rlzero.mcts import deepmind_mcts as dm

    class Adapter(ref.GomokuEnv):                 # deepmind_mcts.py:497 calls legal_actions() bare
        def legal_actions(self, player=None):
            return list(self.leagel_actions())

    def run(env, **kw):
        m = dm.DeepMindMCTS(env, **kw)
        return m

    return dm, Adapter, contextlib, io


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
@pytest.mark.parametrize('size,k,moves,sims,method,solve,eval_id', [
    (3, 3, [], 60, 'puct', True, 2), (3, 3, [0, 3, 1, 4], 80, 'puct', True, 2), (3, 3, [4, 0], 120, 'uct', True, 2),
    (3, 3, [0, 3, 1, 4], 80, 'uct', False, 1), (4, 3, [5, 0, 6], 150, 'puct', True, 2), (6, 4, [14, 15, 20], 200, 'puct', False, 2),
    (5, 4, [12, 7, 13, 8, 11], 300, 'uct', True, 2)])
def test_dm_restatement_matches_live_reference(size, k, moves, sims, method, solve, eval_id):
    from oracle import dm_oracle
    dm, Adapter, contextlib, io = _dm_reference()
    env = Adapter(size, k)
    env.reset()
    mine = pyoracle.DMBoard(size, k)
    mine.reset()
    for a in moves:
        env.step(a)
        mine.step(a)
    ev = dm_oracle.ClosedFormEvaluator(eval_id)
    bot = dm.DeepMindMCTS(env, uct_c=2, max_simulations=sims, evaluator=ev, child_selection_method=method,
                          solve=solve)
    bot._random_state = dm_oracle.NoShuffle()
    with contextlib.redirect_stdout(io.StringIO()):
        root = bot.mcts_search(env)
        best = root.best_child().action
    s = dm_oracle.DMSearch(ev, sims, 2, method, solve=solve)
    got = s.search(mine)
    assert got.n == root.explore_count and got.w == root.total_reward
    assert got.outcome == root.outcome
    assert [(c.action, c.n, c.w, c.outcome) for c in got.children] == \
        [(c.action, c.explore_count, c.total_reward, c.outcome) for c in root.children]
    assert got.best_child().action == best


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_dm_root_noise_matches_live_reference():
    from oracle import dm_oracle
    dm, Adapter, contextlib, io = _dm_reference()
    env = Adapter(4, 3)
    env.reset()
    mine = pyoracle.DMBoard(4, 3)
    mine.reset()
    ev = dm_oracle.ClosedFormEvaluator(2)
    bot = dm.DeepMindMCTS(env, uct_c=2, max_simulations=120, evaluator=ev, child_selection_method='puct',
                          add_exploration_noise=True, dirichlet_noise_alpha=1.0, dirichlet_noise_epsilon=0.25,
                          solve=True)
    bot._random_state = dm_oracle.NoShuffle(np.random.RandomState(11))
    with contextlib.redirect_stdout(io.StringIO()):
        root = bot.mcts_search(env)
    rs = np.random.RandomState(11)
    s = dm_oracle.DMSearch(ev, 120, 2, 'puct', add_exploration_noise=True, dirichlet_noise_epsilon=0.25,
                           solve=True, noise_fn=lambda n: rs.dirichlet([0.25] * n))
    got = s.search(mine)
    assert [(c.action, c.n, c.w, c.prior) for c in got.children] == \
        [(c.action, c.explore_count, c.total_reward, c.prior) for c in root.children]
