"""GPU parity of the Go rule kernels (csrc/rz_go.cu, rz_go.cuh) and the GoEnv mirror against the
Go oracle (oracle/go_oracle.py; parity UNPINNED against the reference, see its header): boards,
ko point, legal-move masks, scores, results, rewards and the 17-plane observation must be
bit-identical at every ply of seeded random games."""
import numpy as np
import pytest
import torch

from oracle.go_oracle import GoEnvOracle

pytestmark = pytest.mark.gpu


def _go():
    from rlzero_b200.games.go import GoBoards, GoEnv
    return GoBoards, GoEnv


def _check_all(gb, envs, n, tag):
    from rlzero_b200 import _lib as L
    boards = gb.boards()
    meta = gb.meta.cpu().numpy()
    legal = gb.legal_mask().cpu().numpy()
    obs = gb.observe().cpu().numpy()
    score, result = (x.cpu().numpy() for x in gb.score())
    for i, env in enumerate(envs):
        pos = env._go
        assert np.array_equal(boards[i], pos.board), (tag, i)
        ko = -1 if pos.ko is None else pos.ko[0] * n + pos.ko[1]
        assert meta[i, L.META_KO] == ko, (tag, i)
        assert meta[i, L.META_PLAYER] == env.current_player(), (tag, i)
        assert meta[i, L.META_STONES] == pos.n, (tag, i)
        want = np.zeros(n * n + 1, dtype=np.uint8)
        want[np.asarray(env.legal_actions(), dtype=np.int64)] = 1
        assert np.array_equal(legal[i], want), (tag, i, np.nonzero(legal[i] != want))
        assert score[i] == pos.score() and result[i] == pos.result(), (tag, i)
        o = env.observe(env.agent_selection)['observation']
        assert np.array_equal(obs[i].transpose(1, 2, 0) != 0, o), (tag, i)
        assert (meta[i, L.META_STATUS] != L.ACTIVE) == env.is_terminal(), (tag, i)


@pytest.mark.parametrize('n,games,plies,komi', [(3, 24, 60, 0.5), (5, 24, 150, 2.5), (9, 8, 220, 7.5),
                                               (19, 2, 140, 7.5)])
def test_random_games_match_the_oracle(n, games, plies, komi):
    GoBoards, _ = _go()
    rs = np.random.RandomState(100 + n)
    gb = GoBoards(games, n, komi)
    envs = [GoEnvOracle(n, komi) for _ in range(games)]
    for e in envs:
        e.reset()
    _check_all(gb, envs, n, 'start')
    for t in range(plies):
        acts = []
        for e in envs:
            if e.is_terminal():
                acts.append(-1)
                continue
            legal = list(e.legal_actions())
            # pass rarely while the board is open, often when it is nearly full (so games end)
            if len(legal) > 1 and rs.rand() > (0.02 if len(legal) > n else 0.3):
                legal = legal[:-1]
                a = int(legal[rs.randint(len(legal))])
            else:
                a = n * n
            acts.append(a)
        reward, done = gb.step(acts)
        reward, done = reward.cpu().numpy(), done.cpu().numpy()
        for i, (e, a) in enumerate(zip(envs, acts)):
            if a < 0:
                continue
            e.step(a)
            assert bool(done[i]) == e.is_terminal()
            assert list(reward[i]) == [int(x) for x in e.returns()]
        assert not gb.faults().any()
        _check_all(gb, envs, n, 'ply %d' % t)
        if all(e.is_terminal() for e in envs):
            break
    assert any(e.is_terminal() for e in envs) or n == 19


def test_illegal_moves_fault():
    from rlzero_b200 import _lib as L
    GoBoards, _ = _go()
    gb = GoBoards(3, 5, 7.5)
    gb.step([0, 0, 0])
    gb.step([0, 26, 1])          # occupied, out of range, legal
    f = gb.faults()
    assert f[0] & L.FAULT_ILLEGAL_MOVE and f[1] & L.FAULT_ILLEGAL_MOVE and not f[2]


def test_goenv_mirror_matches_the_oracle_step_by_step():
    _, GoEnv = _go()
    rs = np.random.RandomState(5)
    env, ref = GoEnv(board_size=5, komi=7.5), GoEnvOracle(5, 7.5)
    env.reset()
    ref.reset()
    assert env.current_player() == 0 and list(env.legal_actions()) == list(ref.legal_actions())
    for t in range(200):
        legal = list(ref.legal_actions())
        a = int(legal[rs.randint(len(legal))]) if rs.rand() > 0.05 else 25
        got, want = env.step(a), ref.step(a)
        assert np.array_equal(got.obs['observation'], want[0]['observation'])
        assert np.array_equal(got.obs['action_mask'], want[0]['action_mask'])
        assert got.reward == want[1] and isinstance(got.reward, float)
        assert got.done == want[2] and isinstance(got.done, bool)
        assert env.current_player() == ref.current_player() and env.returns() == ref.returns()
        assert list(env.legal_actions()) == list(ref.legal_actions())
        if got.done:
            break
    assert env.is_terminal() and env.max_utility() == 1
    with pytest.raises(RuntimeError):
        env.step(25)
    c = env.clone()
    assert c.returns() == env.returns() and c is not env


def test_goenv_reference_smoke_test():
    """The reference's own env test (rlzero/games/go/test_go_env.py:9-39): random play on 9x9,
    obs dict / done bool / reward float."""
    _, GoEnv = _go()
    env = GoEnv(board_size=9, komi=7.5)
    env.seed(0)
    env.reset()
    for i in range(100):
        obs, reward, done, info = env.step(env.random_action())
        assert isinstance(obs, dict) and isinstance(done, bool) and isinstance(reward, float)
        if done:
            break
        illegal = np.nonzero(obs['action_mask'] == 0)[0]
        if len(illegal):
            with pytest.raises(ValueError):
                env.step(int(illegal[0]))


def test_hand_derived_positions_on_the_device():
    """The known-answer sequences of tests/test_go_oracle.py (hand-derived: two groups captured by one move, corner
    and chain captures, suicide of a chain, ko, Tromp-Taylor with dame) replayed through rz_go_step: boards, ko, legal
    masks and scores on the device equal the hand-derived answers, not just the oracle's."""
    from rlzero_b200 import _lib as L
    GoBoards, _ = _go()
    n = 5
    P = n * n   # the pass
    col_walls = [a for r in range(n) for a in (r * n + 1, r * n + 3)]
    seqs = [
        [1, 6, 11, 8, 5, 24, 3, 23, 13, 22, 9, 21, 7],              # (1,2) captures (1,1) and (1,3)
        [24, 0, 1, 20, 5],                                            # corner capture
        [0, 1, 6, 2, 7, 24, 3],                                       # two-stone chain captured on the edge
        [0, 5, 24, 6, 23, 2],                                         # black (0,1) would be suicide
        [1, 2, 5, 8, 11, 12, 24, 6, 7],                               # ko: black retakes at (1,2), ko at (1,1)
        col_walls,                                                    # walls on columns 1 and 3
        col_walls + [14, P],                                          # + a black stone inside white's column
    ]
    gb = GoBoards(len(seqs), n, 0.5)
    for t in range(max(len(s) for s in seqs)):
        gb.step([s[t] if t < len(s) else -1 for s in seqs])
    assert not gb.faults().any()
    boards = gb.boards()
    meta = gb.meta.cpu().numpy()
    legal = gb.legal_mask().cpu().numpy()
    score, result = (x.cpu().numpy() for x in gb.score())
    # 0: both white stones gone, no ko, (1,1) / (1,3) are suicide for white
    assert boards[0][1, 1] == 0 and boards[0][1, 3] == 0 and boards[0][1, 2] != 0 and meta[0, L.META_KO] == -1
    assert legal[0][6] == 0 and legal[0][8] == 0
    # 1: the corner stone is gone, no ko
    assert boards[1][0, 0] == 0 and meta[1, L.META_KO] == -1
    # 2: the chain (0,1),(0,2) is gone
    assert boards[2][0, 1] == 0 and boards[2][0, 2] == 0 and meta[2, L.META_KO] == -1
    # 3: black to move, (0,1) is illegal (suicide of the chain), the pass is legal
    assert meta[3, L.META_PLAYER] == 0 and legal[3][1] == 0 and legal[3][P] == 1
    # 4: white stone at (1,1) captured, ko at (1,1): white may not retake
    assert boards[4][1, 1] == 0 and meta[4, L.META_KO] == 6 and legal[4][6] == 0
    # 5 / 6: Tromp-Taylor with komi 0.5: 10 - 10 - 0.5; 11 - 5 - 0.5
    assert score[5] == -0.5 and result[5] == -1
    assert score[6] == 5.5 and result[6] == 1
