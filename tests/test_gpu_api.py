"""GPU parity tests of the reference-facing API shim (AlphaZeroMCTS / AlphaZeroPlayer /
GomokuEnv / GameControl) against the golden fixtures of the live reference."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _cases(name, key):
    with open(os.path.join(HERE, 'golden', name)) as f:
        return json.load(f)[key]


SMALL = [c for c in _cases('mcts_kat.json', 'cases') if c['n_playout'] <= 300]


@pytest.mark.parametrize('case', SMALL, ids=lambda c: c['name'])
def test_alphazero_mcts_api_with_python_evaluator(case):
    """The drop-in class with a user-supplied Python policy_value_fn (slow path): same acts,
    same act_probs (bit-exact, host numpy softmax), same TreeNode statistics."""
    from oracle.evaluators import make_policy_value_fn
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.mcts import AlphaZeroMCTS
    size, k = case['board_size'], case['n_in_row']
    env = GomokuEnv(size, k)
    env.reset()
    for m in case['pre_moves']:
        env.step(m)
    mcts = AlphaZeroMCTS(make_policy_value_fn(case['eval_id']), n_playout=case['n_playout'],
                         c_puct=case['c_puct'], add_noise=False,
                         rule=L.RULE_PUCT if case['rule'] == 'puct' else L.RULE_UCT)

    def check(st):
        acts, probs = mcts.simulate(env, 1.0)
        assert list(acts) == st['acts']
        assert [float(p).hex() for p in probs] == st['probs_T1']
        root = mcts._root
        assert root.explore_count == st['root_N']
        assert float(root.total_reward).hex() == st['root_W']
        assert [root._children[a].explore_count for a in acts] == [st['visits'][a] for a in acts]
        assert [float(root._children[a].total_reward).hex() for a in acts] == [st['W'][a] for a in acts]
        assert not root.is_leaf() and root.is_root()

    check(case['stages'][0])
    for m, st in zip(case['chain'], case['stages'][1:]):
        env.step(m)
        mcts.update_with_move(m)
        check(st)
    # the caller's env is never mutated by the search (deepcopy in the reference, :83)
    assert len(env.states) == len(case['pre_moves']) + len(case['chain'])


@pytest.mark.parametrize('ep', _cases('selfplay.json', 'episodes'),
                         ids=lambda e: '%dx%d_n%d' % (e['board_size'], e['board_size'], e['n_playout']))
def test_start_self_play_episode_matches_reference(ep):
    """GameControl.start_self_play + AlphaZeroPlayer(is_selfplay=True) reproduce the reference
    episode move for move under the same global numpy seed (noise on: one dirichlet draw per
    expansion keeps the RNG stream aligned; visits do not depend on priors in UCT mode)."""
    from oracle.evaluators import make_policy_value_fn
    from rlzero_b200.games.gomoku import GameControl, GomokuEnv
    from rlzero_b200.mcts import AlphaZeroPlayer
    np.random.seed(ep['seed'])
    env = GomokuEnv(ep['board_size'], ep['n_in_row'])
    game = GameControl(env)
    player = AlphaZeroPlayer(make_policy_value_fn(ep['eval_id']), n_playout=ep['n_playout'],
                             c_puct=5, is_selfplay=True)
    winner, data = game.start_self_play(player, temperature=ep['temperature'])
    data = list(data)
    assert winner == ep['winner']
    assert [int(m) for m in env.states.keys()] == ep['moves']
    assert len(data) == len(ep['records'])
    for (state, pi, z), rec in zip(data, ep['records']):
        sha = hashlib.sha1(np.ascontiguousarray(state.astype(np.float32)).tobytes()).hexdigest()
        assert sha == rec['state_sha1']
        assert [float(x).hex() for x in pi] == rec['pi']
        assert float(z) == rec['z']


def test_start_play_and_player_shell():
    from oracle.evaluators import EVAL_HASH, EVAL_KAT, make_policy_value_fn
    from rlzero_b200.games.gomoku import GameControl, GomokuEnv
    from rlzero_b200.mcts import AlphaZeroPlayer
    np.random.seed(3)
    env = GomokuEnv(4, 3)
    game = GameControl(env)
    p1 = AlphaZeroPlayer(make_policy_value_fn(EVAL_HASH), n_playout=30, player_name='a')
    p2 = AlphaZeroPlayer(make_policy_value_fn(EVAL_KAT), n_playout=30, player_name='b')
    winner = game.start_play(p1, p2, start_player=1, is_shown=False)
    assert winner in (-1, 0, 1)
    assert (p1.get_player_id(), p2.get_player_id()) == (0, 1)
    assert env.game_end_winner()[0]
    assert 'AlphaZeroPlayer' in str(p1)


def test_terminal_root_and_illegal_inputs():
    from oracle.evaluators import EVAL_KAT, make_policy_value_fn
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.games.gomoku.gomoku_env import Error
    from rlzero_b200.mcts import AlphaZeroMCTS
    env = GomokuEnv(3, 3)
    env.reset()
    for m in (0, 3, 1, 4, 2):
        env.step(m)
    assert env.game_end_winner() == (True, 0)
    mcts = AlphaZeroMCTS(make_policy_value_fn(EVAL_KAT), n_playout=5)
    with pytest.raises(ValueError):
        mcts.simulate(env, 1.0)   # the reference fails to unpack an empty child list (:90)
    with pytest.raises(Error):
        GomokuEnv(3, 5).reset()
    with pytest.raises(Error):
        GomokuEnv(3, 3).reset(start_player_idx=2)
    with pytest.raises(AttributeError):
        GomokuEnv(3, 3).step(0)    # reset() must come first, as in the reference
