"""CPU tests: the oracle restatement against the fixtures the live reference
generated (tests/golden, scripts/make_golden.py).  Runs everywhere."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import pyoracle
from oracle.evaluators import make_policy_value_fn


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def _cases(name, key):
    here = os.path.join(os.path.dirname(__file__), 'golden', name)
    with open(here) as f:
        return json.load(f)[key]


def check_stage(search, st, n_actions):
    root = search.root
    assert root.n == st['root_N']
    assert float(root.w).hex() == st['root_W']
    visits = search.root_visits(n_actions)
    assert visits.tolist() == st['visits']
    w = search.root_values(n_actions)
    assert [float(x).hex() for x in w] == st['W']
    assert [int(a) for a in root.children.keys()] == st['acts']


@pytest.mark.parametrize('case', _cases('mcts_kat.json', 'cases'), ids=lambda c: c['name'])
def test_mcts_kat(case):
    size, k = case['board_size'], case['n_in_row']
    if case['n_playout'] * size * size > 400 * 361 and os.environ.get('RZ_FAST'):
        pytest.skip('fast mode')
    board = pyoracle.Board(size, k)
    board.reset()
    for m in case['pre_moves']:
        board.step(m)
    rule = pyoracle.RULE_PUCT if case['rule'] == 'puct' else pyoracle.RULE_UCT
    s = pyoracle.Search(make_policy_value_fn(case['eval_id']), case['n_playout'],
                        case['c_puct'], add_noise=False, rule=rule)
    acts, probs = s.simulate(board, 1.0)
    check_stage(s, case['stages'][0], size * size)
    assert [float(p).hex() for p in probs] == case['stages'][0]['probs_T1']
    for m, st in zip(case['chain'], case['stages'][1:]):
        board.step(m)
        s.update_with_move(m)
        acts, probs = s.simulate(board, 1.0)
        check_stage(s, st, size * size)


def test_survey_kat_hashes():
    """SURVEY.md section 4 KAT D/E sha1 of the int32-LE visit vector."""
    by_name = {c['name']: c for c in _cases('mcts_kat.json', 'cases')}
    for name, sha in (('D_6x6_kat', '764a8eb831feca489a6a38dce961189132c1c0b4'),
                      ('E_15x15_kat', '3de9b3fe633c1ae88953e2ad545a54b3a70e97b9')):
        v = np.array(by_name[name]['stages'][0]['visits'], dtype='<i4')
        assert hashlib.sha1(v.tobytes()).hexdigest() == sha
    assert by_name['B_3x3_kat']['stages'][0]['visits'] == [3, 3, 3, 2, 3, 2, 3, 3, 2]
    assert by_name['F_3x3_terminal']['stages'][0]['visits'] == [0, 0, 20, 0, 0, 11, 10, 8, 10]


@pytest.mark.parametrize('game', _cases('env_games.json', 'games'),
                         ids=lambda g: '%dx%d_s%d' % (g['board_size'], g['board_size'], g['seed']))
def test_env_games(game):
    b = pyoracle.Board(game['board_size'], game['n_in_row'])
    b.reset()
    for ply in game['plies']:
        obs, reward, win, _ = b.step(ply['a'])
        end, winner = b.game_end_winner()
        assert (int(reward), bool(win), bool(end), int(winner)) == (
            ply['reward'], ply['win'], ply['end'], ply['winner'])
        assert b.current_player() == ply['player_after']
        assert len(b.leagel_actions()) == ply['n_legal']
        sha = hashlib.sha1(np.ascontiguousarray(obs.astype(np.float32)).tobytes()).hexdigest()
        assert sha == ply['obs_sha1']
    assert b.returns() == game['returns']


@pytest.mark.parametrize('ep', _cases('selfplay.json', 'episodes'),
                         ids=lambda e: '%dx%d_n%d' % (e['board_size'], e['board_size'], e['n_playout']))
def test_selfplay_episode(ep):
    """Whole start_self_play episode incl. the global-RNG stream (one dirichlet
    per expansion, one choice per move) -- game.py:96-134."""
    np.random.seed(ep['seed'])
    board = pyoracle.Board(ep['board_size'], ep['n_in_row'])
    player = pyoracle.SearchPlayer(make_policy_value_fn(ep['eval_id']), ep['n_playout'],
                                   c_puct=5, is_selfplay=True)
    winner, data = pyoracle.self_play_episode(board, player, ep['temperature'])
    assert winner == ep['winner']
    assert [int(m) for m in board.states.keys()] == ep['moves']
    assert len(data) == len(ep['records'])
    for (state, pi, z), rec in zip(data, ep['records']):
        sha = hashlib.sha1(np.ascontiguousarray(state.astype(np.float32)).tobytes()).hexdigest()
        assert sha == rec['state_sha1']
        assert [float(x).hex() for x in pi] == rec['pi']
        assert float(z) == rec['z']


# ----------------------------------------------------------------- DeepMindMCTS fixtures
def _dm_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'dm_mcts.json')) as f:
        return json.load(f)


def test_dm_restatement_matches_reference_fixtures():
    """oracle.dm_oracle.DMSearch against vectors the live DeepMindMCTS produced
    (scripts/make_golden_dm.py): visit counts, value sums, priors, outcomes, best child, early stop."""
    import numpy as np
    from oracle import dm_oracle, pyoracle
    cases = _dm_cases()
    assert len(cases) >= 15 and any(c['root_outcome'] is not None for c in cases)
    for c in cases:
        b = pyoracle.DMBoard(c['size'], c['k'])
        b.reset()
        for a in c['moves']:
            b.step(a)
        kw = {}
        if c['noise_seed'] is not None:
            rs = np.random.RandomState(c['noise_seed'])
            kw = dict(add_exploration_noise=True, dirichlet_noise_epsilon=0.25,
                      noise_fn=lambda n, rs=rs: rs.dirichlet([0.25] * n))
        s = dm_oracle.DMSearch(dm_oracle.ClosedFormEvaluator(c['eval_id']), c['sims'], 2, c['method'],
                               solve=c['solve'], **kw)
        root = s.search(b)
        assert root.n == c['root_n'] and root.w == c['root_w'] and root.outcome == c['root_outcome'], c['moves']
        got = [[ch.action, ch.n, ch.w, ch.outcome, ch.prior] for ch in root.children]
        assert got == c['children'], c['moves']
        assert root.best_child().action == c['best']
