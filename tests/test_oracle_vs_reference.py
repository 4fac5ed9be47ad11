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
