"""GPU parity of the DeepMindMCTS flavour (rlzero/mcts/deepmind_mcts.py:384-646; RZ_FLAVOUR_DEEPMIND)
against vectors produced by the LIVE reference class (tests/golden/dm_mcts.json) and against the
restatement (oracle.dm_oracle): root children's visit counts, value sums and outcomes, the root's own
N / W / outcome (MCTS-Solver, early stop on a proven root) and best_child must match bit-exactly."""
import json
import os

import numpy as np
import pytest

from oracle import dm_oracle, pyoracle

pytestmark = pytest.mark.gpu


def _decode(code):
    return None if code == 0 else [(code & 3) - 1, ((code >> 2) & 3) - 1]


def _device_search(size, k, moves_list, sims, method, solve, eval_id, returns_mode=0, game_type=None, **kw):
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    f = SearchForest(len(moves_list), size, k, n_playout=sims, c_puct=2.0,
                     rule=L.RULE_PUCT if method == 'puct' else L.RULE_UCT, flavour=L.FLAVOUR_DEEPMIND,
                     solve=solve, returns_mode=returns_mode, max_carry=0,
                     game_type=L.GAME_GOMOKU if game_type is None else game_type, **kw)
    f.set_positions(moves_list)
    f.search(ClosedFormEvaluator(eval_id))
    f.raise_faults()
    return f


def _compare(f, g, want_children, root_n, root_w, root_outcome, best):
    d = f.dump_tree(g)
    got = []
    for a in range(f.A):
        n = int(d['N'][0][a]) if d['n_nodes'] > 0 else -1
        if n < 0:
            continue
        got.append([a, n, float(d['W'][0][a]) if n > 0 else 0.0, _decode(int(d['O'][0][a]))])
    assert got == [c[:4] for c in want_children]
    assert d['root_N'] == root_n and d['root_W'] == root_w and _decode(d['root_O']) == root_outcome
    b, _ = f.best_child()
    assert int(b[g]) == best


def test_matches_live_reference_fixtures():
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'dm_mcts.json')) as fh:
        cases = [c for c in json.load(fh) if c['noise_seed'] is None]
    assert any(c['root_outcome'] is not None for c in cases)
    for c in cases:
        f = _device_search(c['size'], c['k'], [c['moves']], c['sims'], c['method'], c['solve'], c['eval_id'])
        _compare(f, 0, c['children'], c['root_n'], c['root_w'], c['root_outcome'], c['best'])


@pytest.mark.parametrize('size,k,sims,method,solve,returns_mode', [
    (3, 3, 150, 'puct', True, 1), (3, 3, 300, 'uct', True, 1), (4, 3, 250, 'puct', True, 0), (5, 4, 300, 'uct', True, 1),
    (6, 4, 200, 'puct', False, 1), (4, 4, 500, 'puct', True, 1)])
def test_batched_random_positions_match_the_restatement(size, k, sims, method, solve, returns_mode):
    """Many positions at once, both returns conventions (the reference's and the zero-sum one)."""
    rs = np.random.RandomState(size * 31 + sims)
    G = 12
    boards, lists = [], []
    for g in range(G):
        while True:
            b = pyoracle.DMBoard(size, k, zero_sum=bool(returns_mode))
            b.reset()
            mv = [int(x) for x in rs.permutation(size * size)[:rs.randint(0, size * size - 2)]]
            ok = True
            for a in mv:
                b.step(a)
                if b.game_end_winner()[0]:
                    ok = False
                    break
            if ok:
                break
        boards.append(b)
        lists.append(mv)
    f = _device_search(size, k, lists, sims, method, solve, 2, returns_mode)
    for g, b in enumerate(boards):
        s = dm_oracle.DMSearch(dm_oracle.ClosedFormEvaluator(2), sims, 2, method, solve=solve)
        root = s.search(b)
        want = [[ch.action, ch.n, ch.w, ch.outcome] for ch in root.children]
        _compare(f, g, want, root.n, root.w, root.outcome, root.best_child().action if root.children else -1)


def test_go_positions_with_the_deepmind_driver():
    """The reference's Go path is DeepMindMCTS on GoEnv (rlzero/games/go/test_mcts_bot.py): same
    search on the device Go rules vs the restatement over the Go oracle (returns [1,-1] / [-1,1])."""
    from rlzero_b200 import _lib as L
    from oracle.go_oracle import GoSearchBoard

    class GoDM(GoSearchBoard):
        def legal_actions(self, player=None):
            return self.leagel_actions()

        def is_terminal(self):
            return self.game_end_winner()[0]

        def returns(self):
            end, w = self.game_end_winner()
            return [0, 0] if not end else ([1, -1] if w == 0 else [-1, 1])

    n, sims = 3, 400
    rs = np.random.RandomState(2)
    boards, lists = [], []
    for g in range(8):
        b = GoDM(n, 0.5)
        mv = []
        for _ in range(rs.randint(0, 8)):
            legal = b.leagel_actions()
            a = legal[rs.randint(len(legal))]
            b.step(a)
            mv.append(a)
            if b.game_end_winner()[0]:
                b.reset()
                mv = []
        boards.append(b)
        lists.append(mv)
    f = _device_search(n, 1, lists, sims, 'puct', True, 2, game_type=L.GAME_GO, komi=0.5)
    proven = 0
    for g, b in enumerate(boards):
        s = dm_oracle.DMSearch(dm_oracle.ClosedFormEvaluator(2), sims, 2, 'puct', solve=True)
        root = s.search(b)
        want = [[ch.action, ch.n, ch.w, ch.outcome] for ch in root.children]
        _compare(f, g, want, root.n, root.w, root.outcome, root.best_child().action)
        proven += root.outcome is not None
    assert proven >= 0


def test_root_only_noise():
    """add_exploration_noise perturbs the priors of the root's children only (deepmind_mcts.py:484-485)."""
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    f = SearchForest(2, 5, 4, n_playout=80, c_puct=2.0, rule=L.RULE_PUCT, flavour=L.FLAVOUR_DEEPMIND, solve=True,
                     max_carry=0, noise_root_only=True)
    f.run_waves(80, ClosedFormEvaluator(L.EVAL_ZERO), noise_eps=0.25, noise_alpha=0.25, seed=3)
    f.raise_faults()
    d = f.dump_tree(0)
    uni = np.float32(1.0) / np.float32(25)
    assert not np.allclose(d['P'][0], uni)                    # root: noisy
    assert abs(float(d['P'][0].sum()) - 1.0) < 1e-5
    for node in range(1, d['n_nodes']):
        legal = d['N'][node] >= 0
        assert np.all(d['P'][node][legal] == np.float32(1.0) / np.float32(legal.sum()))   # below: untouched


# ------------------------------------------------------------------ the reference-facing classes
def test_deepmind_mcts_class_with_a_python_evaluator_matches_fixtures():
    """rlzero_b200.mcts.DeepMindMCTS(env, evaluator=<user Evaluator>) on the GPU GomokuEnv: the root
    SearchNode view carries the numbers the live reference produced."""
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.mcts import DeepMindMCTS
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'dm_mcts.json')) as fh:
        cases = [c for c in json.load(fh) if c['noise_seed'] is None and c['size'] <= 6][:8]
    for c in cases:
        env = GomokuEnv(c['size'], c['k'])
        env.reset()
        for a in c['moves']:
            env.step(a)
        bot = DeepMindMCTS(env, uct_c=2, max_simulations=c['sims'], evaluator=dm_oracle.ClosedFormEvaluator(c['eval_id']),
                           child_selection_method=c['method'], solve=c['solve'], child_shuffle=False)
        root = bot.mcts_search(env)
        assert root.explore_count == c['root_n'] and root.total_reward == c['root_w'] and root.outcome == c['root_outcome']
        got = [[ch.action, ch.explore_count, ch.total_reward, ch.outcome] for ch in root.children]
        assert got == [x[:4] for x in c['children']]
        assert root.best_child().action == c['best']
        policy, action = bot.step_with_policy(env)
        assert action == c['best'] and [a for a, _ in policy] == list(env.legal_actions())
        assert sum(p for _, p in policy) == 1.0


def test_random_rollout_evaluator_on_the_device():
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.mcts import DeepMindMCTS, RandomRolloutEvaluator
    env = GomokuEnv(3, 3)
    env.reset()
    for a in [0, 3, 1, 4]:          # player 0 to move, wins at 2; player 1 threatens 5
        env.step(a)
    bot = DeepMindMCTS(env, uct_c=2, max_simulations=300, evaluator=RandomRolloutEvaluator(n_rollouts=8, seed=3),
                       solve=True, returns_mode=L.RETURNS_ZERO_SUM)
    root = bot.mcts_search(env)
    kids = {c.action: c for c in root.children}
    assert sorted(kids) == [2, 5, 6, 7, 8]
    assert kids[2].outcome == [1, -1] and root.outcome == [1, -1]        # proven win, search stopped early
    assert root.explore_count < 300
    assert bot.step(env) == 2
    for c in root.children:
        assert abs(c.total_reward) <= c.explore_count


def test_mcts_bot_plays_go_like_the_reference_script():
    """rlzero/games/go/test_mcts_bot.py:9-43: DeepMindMCTS picks moves for black on a Go board,
    white answers with a random legal move; here on 5x5 to the end of the game."""
    from rlzero_b200.games.go import GoEnv
    from rlzero_b200.mcts import DeepMindMCTS, RandomRolloutEvaluator
    env = GoEnv(board_size=5, komi=0.5)
    env.seed(1)
    env.reset()
    bot = DeepMindMCTS(env, max_simulations=40, evaluator=RandomRolloutEvaluator(n_rollouts=2, n_limit=60, seed=1),
                       solve=True)
    for ply in range(60):
        if env.is_terminal():
            break
        if ply % 2 == 0:
            policy, action = bot.step_with_policy(env)
            assert action in list(env.legal_actions()) and len(policy) == len(env.legal_actions())
        else:
            action = env.random_action()
        env.step(int(action))
    assert ply > 4
