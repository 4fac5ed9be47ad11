"""The plain-C oracle (oracle/c/rz_oracle.c, built by oracle/build_oracle.py) against the golden vectors the
LIVE reference produced (tests/golden/mcts_kat.json: KAT A-F of SURVEY.md section 4 and more, UCT and PUCT, tree
reuse, terminal leaves, ties) and against the Python restatement on random positions.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import build_oracle, pyoracle
from oracle.evaluators import make_policy_value_fn


def test_c_oracle_matches_live_reference_fixtures(golden_dir):
    with open(os.path.join(golden_dir, 'mcts_kat.json')) as f:
        cases = json.load(f)['cases']
    assert len(cases) >= 16
    for c in cases:
        rule = 1 if c['rule'] == 'puct' else 0
        visits, w, rn, rw = build_oracle.search_game(c['board_size'], c['n_in_row'], c['pre_moves'], c['n_playout'],
                                                     c['c_puct'], rule, c['eval_id'], follow=c['chain'])
        assert len(c['stages']) == len(c['chain']) + 1
        for j, st in enumerate(c['stages']):
            assert visits[j].tolist() == st['visits'], (c['name'], j)
            assert [float(x).hex() for x in w[j]] == [float.fromhex(x).hex() for x in st['W']], (c['name'], j)
            assert int(rn[j]) == st['root_N'] and float(rw[j]).hex() == float.fromhex(st['root_W']).hex()


@pytest.mark.parametrize('size,k,n_playout,rule,eval_id,cpuct', [(3, 3, 150, 0, 2, 5.0), (4, 3, 200, 1, 2, 2.0),
                                                                 (6, 4, 250, 0, 1, 5.0), (8, 5, 300, 1, 2, 5.0),
                                                                 (5, 4, 120, 0, 0, 5.0)])
def test_c_oracle_matches_python_restatement_on_random_positions(size, k, n_playout, rule, eval_id, cpuct):
    rs = np.random.RandomState(size * 7 + n_playout)
    lists, boards = [], []
    while len(lists) < 10:
        b = pyoracle.Board(size, k)
        b.reset()
        mv = [int(x) for x in rs.permutation(size * size)[:rs.randint(0, size * size - 1)]]
        ok = True
        for a in mv:
            b.step(a)
            if b.game_end_winner()[0]:
                ok = False
                break
        if ok:
            lists.append(mv)
            boards.append(b)
    visits, w, rn, rw = build_oracle.search_batch(size, k, lists, n_playout, cpuct, rule, eval_id)
    for g, b in enumerate(boards):
        s = pyoracle.Search(make_policy_value_fn(eval_id), n_playout, cpuct, rule=rule)
        s.simulate(b, 1.0)
        assert visits[g].tolist() == s.root_visits(size * size).tolist(), g
        assert [float(x).hex() for x in w[g]] == [float(x).hex() for x in s.root_values(size * size)], g
        assert int(rn[g]) == s.root.n and float(rw[g]) == float(s.root.w)


@pytest.mark.parametrize('size,k,n_playout,rule,K,vl', [(3, 3, 90, 0, 4, 1.0), (6, 4, 200, 0, 8, 1.0),
                                                        (8, 5, 250, 1, 16, 1.0), (5, 4, 300, 0, 32, 0.5)])
def test_c_oracle_leaf_parallel_wave_matches_python_restatement(size, k, n_playout, rule, K, vl):
    """The leaf-parallel wave (the product's opt-in mode, parity unpinned) restated twice -- oracle/c/rz_oracle.c and
    pyoracle.Search.wave -- must agree bit for bit; the C one checks the kernels at full size."""
    rs = np.random.RandomState(size * 11 + K)
    lists, boards = [], []
    while len(lists) < 8:
        b = pyoracle.Board(size, k)
        b.reset()
        mv = [int(x) for x in rs.permutation(size * size)[:rs.randint(0, size * size - 1)]]
        ok = True
        for a in mv:
            b.step(a)
            if b.game_end_winner()[0]:
                ok = False
                break
        if ok:
            lists.append(mv)
            boards.append(b)
    visits, w, rn, rw = build_oracle.search_batch_vl(size, k, lists, n_playout, 2.5, rule, 2, K, vl)
    for g, b in enumerate(boards):
        s = pyoracle.Search(make_policy_value_fn(2), n_playout, 2.5, rule=rule, leaves_per_wave=K, virtual_loss=vl)
        s.simulate(b, 1.0)
        assert visits[g].tolist() == s.root_visits(size * size).tolist(), g
        assert [float(x).hex() for x in w[g]] == [float(x).hex() for x in s.root_values(size * size)], g
        assert int(rn[g]) == s.root.n and float(rw[g]) == float(s.root.w)


@pytest.mark.parametrize('n_playout,rule,K', [(200, 0, 1), (300, 1, 1), (250, 0, 8)])
def test_c_oracle_connect_four_matches_python_restatement(n_playout, rule, K):
    """Connect Four in the C oracle (gravity, 6x7, actions = columns) against pyoracle.ConnectFourBoard -- which the
    live reference MCTS pins through tests/golden/connect4.json -- for the sequential search and the leaf-parallel wave."""
    rs = np.random.RandomState(n_playout + K)
    lists, boards = [], []
    while len(lists) < 10:
        b = pyoracle.ConnectFourBoard()
        b.reset()
        mv = [int(x) for x in rs.permutation(np.repeat(np.arange(7), 6))[:rs.randint(0, 30)]]
        ok = True
        for a in mv:
            b.step(a)
            if b.game_end_winner()[0]:
                ok = False
                break
        if ok:
            lists.append(mv)
            boards.append(b)
    visits, w, rn, rw = build_oracle.search_batch_c4(lists, n_playout, 5.0, rule, 2, leaves_per_wave=K)
    for g, b in enumerate(boards):
        s = pyoracle.Search(make_policy_value_fn(2), n_playout, 5.0, rule=rule, leaves_per_wave=K)
        s.simulate(b, 1.0)
        assert visits[g].tolist() == s.root_visits(7).tolist(), g
        assert [float(x).hex() for x in w[g]] == [float(x).hex() for x in s.root_values(7)], g
        assert int(rn[g]) == s.root.n and float(rw[g]) == float(s.root.w)


def _enc(o):
    return 0 if o is None else (0x100 | (int(o[0]) + 1) | ((int(o[1]) + 1) << 2))


def test_c_oracle_deepmind_mcts_matches_live_reference_fixtures(golden_dir):
    """The C restatement of DeepMindMCTS against the vectors the LIVE reference class produced (all fixtures without
    root noise): root children's N / W / outcome, the root's N / W / outcome (solver, early stop), best_child."""
    with open(os.path.join(golden_dir, 'dm_mcts.json')) as f:
        cases = [c for c in json.load(f) if c['noise_seed'] is None]
    assert len(cases) >= 18
    for c in cases:
        r = build_oracle.dm_search_batch(c['size'], c['k'], [c['moves']], c['sims'], 2.0, c['method'], c['solve'], 0,
                                         c['eval_id'])
        want = {int(ch[0]): ch for ch in c['children']}
        for a in range(c['size'] ** 2):
            if a in want:
                _, n, w, o, _ = want[a]
                assert int(r['visits'][0, a]) == n and float(r['w'][0, a]).hex() == float(w).hex(), (c['moves'], a)
                assert int(r['outcome'][0, a]) == _enc(o)
            else:
                assert int(r['visits'][0, a]) == -1
        assert int(r['root_n'][0]) == c['root_n'] and float(r['root_w'][0]).hex() == float(c['root_w']).hex()
        assert int(r['root_outcome'][0]) == _enc(c['root_outcome']) and int(r['best'][0]) == c['best']


@pytest.mark.parametrize('size,k,sims,method,solve,mode', [(3, 3, 200, 'puct', True, 1), (4, 3, 300, 'uct', True, 0),
                                                           (5, 4, 300, 'puct', True, 1), (6, 4, 250, 'uct', False, 1)])
def test_c_oracle_deepmind_mcts_matches_python_restatement(size, k, sims, method, solve, mode):
    from oracle import dm_oracle
    rs = np.random.RandomState(size * 13 + sims)
    lists, boards = [], []
    while len(lists) < 10:
        b = pyoracle.DMBoard(size, k, zero_sum=bool(mode))
        b.reset()
        mv = [int(x) for x in rs.permutation(size * size)[:rs.randint(0, size * size - 2)]]
        ok = True
        for a in mv:
            b.step(a)
            if b.game_end_winner()[0]:
                ok = False
                break
        if ok:
            lists.append(mv)
            boards.append(b)
    r = build_oracle.dm_search_batch(size, k, lists, sims, 2.0, method, solve, mode, 2)
    for g, b in enumerate(boards):
        s = dm_oracle.DMSearch(dm_oracle.ClosedFormEvaluator(2), sims, 2, method, solve=solve)
        root = s.search(b)
        for ch in root.children:
            assert int(r['visits'][g, ch.action]) == ch.n and float(r['w'][g, ch.action]).hex() == float(ch.w).hex()
            assert int(r['outcome'][g, ch.action]) == _enc(ch.outcome)
        assert int(r['root_n'][g]) == root.n and float(r['root_w'][g]).hex() == float(root.w).hex()
        assert int(r['root_outcome'][g]) == _enc(root.outcome)
        assert int(r['best'][g]) == root.best_child().action


@pytest.mark.parametrize('size,k,n_playout,rule', [(6, 4, 150, 0), (8, 5, 200, 1)])
def test_c_oracle_tree_reuse_batch_matches_python_restatement(size, k, n_playout, rule):
    rs = np.random.RandomState(size + n_playout)
    lists, boards = [], []
    while len(lists) < 8:
        b = pyoracle.Board(size, k)
        b.reset()
        mv = [int(x) for x in rs.permutation(size * size)[:rs.randint(0, size * size // 2)]]
        ok = True
        for a in mv:
            b.step(a)
            if b.game_end_winner()[0]:
                ok = False
                break
        if ok:
            lists.append(mv)
            boards.append(b)
    move, visits, w, rn, rw = build_oracle.search_batch_reuse(size, k, lists, n_playout, 5.0, rule, 2)
    for g, b in enumerate(boards):
        s = pyoracle.Search(make_policy_value_fn(2), n_playout, 5.0, rule=rule)
        s.simulate(b, 1.0)
        m = int(np.argmax(s.root_visits(size * size)))
        assert int(move[g]) == m
        b.step(m)
        if b.game_end_winner()[0]:
            assert int(rn[g]) == 0
            continue
        s.update_with_move(m)
        s.simulate(b, 1.0)
        assert visits[g].tolist() == s.root_visits(size * size).tolist(), g
        assert [float(x).hex() for x in w[g]] == [float(x).hex() for x in s.root_values(size * size)], g
        assert int(rn[g]) == s.root.n and float(rw[g]) == float(s.root.w)


@pytest.mark.parametrize('size,k', [(3, 3), (6, 4), (15, 5), (19, 5)])
def test_c_oracle_game_replay_matches_python_board(size, k):
    """GomokuEnv.step / game_end_winner in the C oracle against pyoracle.Board (pinned to the live reference's env games
    by tests/test_oracle_golden.py) on random full games: same ending ply, same winner, ties included."""
    rs = np.random.RandomState(size)
    G = 40
    moves = np.stack([rs.permutation(size * size) for _ in range(G)]).astype(np.int32)
    end_ply, winner, ended = build_oracle.replay_games(size, k, moves)
    for g in range(G):
        b = pyoracle.Board(size, k)
        b.reset()
        t, w, e = 0, -1, False
        for a in moves[g]:
            b.step(int(a))
            t += 1
            e, w = b.game_end_winner()
            if e:
                break
        assert (int(end_ply[g]), int(winner[g]), bool(ended[g])) == (t, int(w), bool(e)), g
    assert ended.all()
