"""GPU parity tests of the round-2 boundary work: seeded Dirichlet noise (host-callback path and the `noise64`
input of rz_tree_expand_backup_ex), the seeded DeepMindMCTS child shuffle, large kept subtrees, graph re-capture
after a buffer re-allocation, the per-search device seed, outcome codes across a re-root."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _cases(name):
    with open(os.path.join(HERE, 'golden', name)) as f:
        return json.load(f)['cases']


def _hexes(xs):
    return [float(x).hex() for x in xs]


# ------------------------------------------------------------------ seeded noise, reference API
@pytest.mark.parametrize('case', _cases('mcts_noise.json'), ids=lambda c: c['name'])
def test_seeded_noise_through_the_reference_api(case):
    """AlphaZeroMCTS(add_noise=True) with a Python policy_value_fn under np.random.seed(s): the noise comes from the
    global numpy stream in the reference's order, the priors are kept in float64 (PUCT), so visits, value sums and
    priors equal the live reference's bit for bit (tests/golden/mcts_noise.json), tree reuse included."""
    from oracle.evaluators import make_policy_value_fn
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.mcts import AlphaZeroMCTS
    size, k = case['size'], case['k']
    env = GomokuEnv(size, k)
    env.reset()
    for m in case['pre']:
        env.step(m)
    puct = case['rule'] == 'puct'
    np.random.seed(case['seed'])
    mcts = AlphaZeroMCTS(make_policy_value_fn(case['eval_id']), n_playout=case['n_playout'], c_puct=case['c_puct'],
                         add_noise=True, rule=L.RULE_PUCT if puct else L.RULE_UCT)

    def check(st):
        acts, probs = mcts.simulate(env, 1.0)
        root = mcts._root
        assert root.explore_count == st['root_N'] and float(root.total_reward).hex() == st['root_W']
        assert [root._children[a].explore_count for a in acts] == [st['visits'][a] for a in acts]
        assert _hexes(root._children[a].total_reward for a in acts) == [st['W'][a] for a in acts]
        if puct:
            assert _hexes(root._children[a].prior for a in acts) == [st['prior'][a] for a in acts]
        else:       # UCB1 never reads a prior: the pool keeps float32
            assert [float(np.float32(root._children[a].prior)) for a in acts] == \
                [float(np.float32(float.fromhex(st['prior'][a]))) for a in acts]
        return acts, probs

    acts, probs = check(case['stages'][0])
    assert list(acts) == case['stages'][0]['acts'] and _hexes(probs) == case['stages'][0]['probs']
    for m, st in zip(case['chain'], case['stages'][1:]):
        env.step(m)
        mcts.update_with_move(m)
        check(st)


@pytest.mark.parametrize('case', [c for c in _cases('mcts_noise.json') if c['eval_id'] == 2 and not c['chain'][1:]][:3],
                         ids=lambda c: c['name'])
def test_noise64_input_of_the_c_abi(case):
    """rz_tree_expand_backup_ex(noise64=...): the evaluator runs on the device (closed form, float32 priors), the
    host supplies the Dirichlet samples it drew from np.random.seed(s) in the reference's order; the kernel mixes
    them with numpy's arithmetic into the float64 prior pool.  First stage of the fixtures, bit for bit."""
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import SearchForest
    size, k = case['size'], case['k']
    rule = L.RULE_PUCT if case['rule'] == 'puct' else L.RULE_UCT
    f = SearchForest(1, size, k, n_playout=case['n_playout'], c_puct=case['c_puct'], rule=rule, prior_f64=True)
    f.set_positions([case['pre']])
    np.random.seed(case['seed'])
    noise = torch.zeros(1, f.AS, dtype=torch.float64, device='cuda')
    for _ in range(case['n_playout']):
        f.select()
        rows, meta, depth = f.leaf_boards()
        host = np.zeros(f.AS)
        if meta[0][L.META_STATUS] == L.ACTIVE:
            occ = rows[0][0] | rows[0][1]
            legal = [r * size + c for r in range(size) for c in range(size) if not (int(occ[r]) >> c) & 1]
            host[legal] = np.random.dirichlet(0.3 * np.ones(len(legal)))
        noise.copy_(torch.from_numpy(host)[None])
        f.eval_closed_form(case['eval_id'])
        L.check(f.lib.rz_tree_expand_backup_ex(C.byref(f.desc), L.ptr(f.prior), 0, L.ptr(f.value), None, 0.25, 0.3,
                                               0, L.ptr(noise), None, L.stream_ptr()), 'rz_tree_expand_backup_ex')
    f.raise_faults()
    st = case['stages'][0]
    visits, w, has, root_n, root_w = f.root_stats()
    assert int(root_n[0]) == st['root_N'] and float(root_w[0]).hex() == st['root_W']
    assert visits[0].tolist() == st['visits'] and _hexes(w[0]) == st['W']
    pri = f.dump_tree(0)['P'][0]
    assert _hexes(np.where(has[0], pri, 0.0)) == st['prior']


def test_expand_backup_ex_rejects_bad_arguments():
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import SearchForest
    f = SearchForest(2, 3, 3, n_playout=8, prior_f64=True)
    f.select()
    z = torch.zeros(2, f.AS, dtype=torch.float64, device='cuda')
    rc = f.lib.rz_tree_expand_backup_ex(C.byref(f.desc), L.ptr(f.prior), 0, L.ptr(f.value), None, 0.25, 0.3, 0,
                                        L.ptr(z), L.ptr(z), L.stream_ptr())
    assert rc != 0 and b'not both' in f.lib.rz_last_error()
    rc = f.lib.rz_tree_expand_backup_ex(C.byref(f.desc), L.ptr(f.prior), 0, L.ptr(f.value), None, 0.0, 0.3, 0,
                                        L.ptr(z), None, L.stream_ptr())
    assert rc != 0 and b'noise_eps' in f.lib.rz_last_error()


# ------------------------------------------------------------------ DeepMindMCTS child shuffle
@pytest.mark.parametrize('case', _cases('dm_mcts_shuffle.json'),
                         ids=lambda c: '%dx%d_%s_seed%d' % (c['size'], c['size'], c['method'], c['seed']))
def test_deepmind_mcts_seeded_shuffle_matches_the_live_reference(case):
    """DeepMindMCTS(random_state=RandomState(seed)): child shuffle (deepmind_mcts.py:508) and root noise from the
    reference's own stream -- children ORDER, statistics, float64 priors and the chosen move as the live reference
    produced them (tests/golden/dm_mcts_shuffle.json)."""
    from oracle import dm_oracle
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.mcts import DeepMindMCTS
    env = GomokuEnv(case['size'], case['k'])
    env.reset()
    for a in case['moves']:
        env.step(a)
    bot = DeepMindMCTS(env, uct_c=2, max_simulations=case['sims'], evaluator=dm_oracle.ClosedFormEvaluator(case['eval_id']),
                       child_selection_method=case['method'], add_exploration_noise=case['noise'], solve=case['solve'],
                       random_state=np.random.RandomState(case['seed']))
    root = bot.mcts_search(env)
    assert root.explore_count == case['root_n'] and float(root.total_reward).hex() == case['root_w']
    assert root.outcome == case['root_outcome']
    got = [[c.action, c.explore_count, float(c.total_reward).hex(), c.outcome, float(c.prior).hex()]
           for c in root.children]
    assert got == case['children']
    assert root.best_child().action == case['best']
    best, _ = bot._forest.best_child()          # the device's best_child breaks ties in list order too
    assert int(best[0]) == case['best']


def test_device_child_shuffle_is_a_random_tie_break():
    """child_shuffle='random' (the graph-capturable default of DeepMindMCTS): with a constant evaluator every score
    ties, so the first moves follow the random child order -- it differs between seeds and between trees, is
    reproducible for a seed, and the search stays a valid one."""
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest

    def first_children(seed):
        f = SearchForest(64, 5, 4, n_playout=3, c_puct=2.0, rule=L.RULE_PUCT, max_carry=0, flavour=L.FLAVOUR_DEEPMIND,
                         child_shuffle='random', global_offset=100)
        f.run_waves(3, ClosedFormEvaluator(L.EVAL_ZERO), seed=seed, use_graph=False)
        f.raise_faults()
        visits = f.root_stats()[0]
        assert (visits.sum(axis=1) == 2).all()
        return [tuple(np.nonzero(v)[0]) for v in visits]
    a, b, c = first_children(1), first_children(1), first_children(2)
    assert a == b and a != c
    assert len(set(a)) > 32                       # trees differ from each other
    assert any(x != (0, 1) for x in a)            # and from the unshuffled lowest-action order


def test_outcome_codes_survive_a_reroot():
    """rz_tree_advance compacts edge_O with the other pools: a proven child keeps its outcome."""
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    f = SearchForest(1, 3, 3, n_playout=300, c_puct=2.0, rule=L.RULE_PUCT, max_carry=300, flavour=L.FLAVOUR_DEEPMIND,
                     solve=True, returns_mode=L.RETURNS_ZERO_SUM)
    f.set_positions([[4, 0]])
    f.run_waves(300, ClosedFormEvaluator(L.EVAL_HASH), use_graph=False)
    before = f.dump_tree(0)
    # a root move whose own children carry proven outcomes
    moves = [a for a in range(9) if before['N'][0][a] > 0 and before['child'][0][a] > 0
             and (before['O'][int(before['child'][0][a])] != 0).any()]
    assert moves, 'the solver proved nothing two plies down'
    move = moves[0]
    child = int(before['child'][0][move])
    old_block = before['O'][child].copy()
    f.advance([move], keep_subtree=True)
    after = f.dump_tree(0)
    assert (after['O'][0] == old_block).all() and (after['N'][0] == before['N'][child]).all()
    assert after['root_O'] == int(before['O'][0][move])


# ------------------------------------------------------------------ kept subtrees larger than 64 nodes
def test_self_play_with_a_large_kept_subtree_matches_the_restatement():
    """on a small board the subtree update_with_move keeps outgrows 64 nodes (4x4 / 400 playouts: 70 at
    ply 10).  The single-game classes size the pool for it; the episode equals the oracle's move for move."""
    from oracle import pyoracle
    from oracle.evaluators import EVAL_HASH, make_policy_value_fn
    from rlzero_b200.games.gomoku import GameControl, GomokuEnv
    from rlzero_b200.mcts import AlphaZeroPlayer
    np.random.seed(5)
    env = GomokuEnv(4, 4)
    player = AlphaZeroPlayer(make_policy_value_fn(EVAL_HASH), n_playout=400, c_puct=5, is_selfplay=True)
    kept = []
    orig = player.mcts.update_with_move

    def spy(move):
        orig(move)
        if move is not None and move >= 0:
            kept.append(int(player.mcts._forest.n_nodes[0]))
    player.mcts.update_with_move = spy
    winner, data = GameControl(env).start_self_play(player, temperature=1.0)
    data = list(data)
    assert max(kept) > 64 and player.mcts._forest.carry_dropped == 0
    np.random.seed(5)
    board = pyoracle.Board(4, 4)
    ref_winner, ref_data = pyoracle.self_play_episode(
        board, pyoracle.SearchPlayer(make_policy_value_fn(EVAL_HASH), n_playout=400, c_puct=5, is_selfplay=True), 1.0)
    assert winner == ref_winner and len(data) == len(ref_data)
    for (s, pi, z), (rs, rpi, rz) in zip(data, ref_data):
        assert np.array_equal(s, rs) and _hexes(pi) == _hexes(rpi) and z == rz


def test_carry_drop_is_counted_not_raised():
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    f = SearchForest(4, 4, 4, n_playout=300, max_carry=8)
    f.run_waves(300, ClosedFormEvaluator(L.EVAL_HASH), use_graph=False)
    visits = f.root_stats()[0]
    f.advance(visits.argmax(axis=1).astype(np.int32), keep_subtree=True)
    f.raise_faults()                              # no exception: the trees restarted from a fresh root
    assert f.carry_dropped == 4 and int(f.n_nodes.sum()) == 0


# ------------------------------------------------------------------ captured graphs
def test_graph_is_recaptured_when_the_evaluator_buffers_grow():
    """NativeForward._alloc frees the buffers a captured wave graph points at.  A larger batch through
    the same evaluator between two searches must not change the second search."""
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    from rlzero_b200.mcts import AlphaZeroMCTS
    torch.manual_seed(0)
    agent = AlphaZeroAgent(6, net=ResNetPolicyValueNet(6, n_blocks=1))
    agent.policy_value_net.eval()
    env = GomokuEnv(6, 4)
    env.reset()
    env.step(14)
    mcts = AlphaZeroMCTS(agent.policy_value_fn, n_playout=48)
    acts, probs = mcts.simulate(env, 1.0)
    version = agent.native.weights_version
    junk = [torch.empty(1 << 20, device='cuda') for _ in range(8)]            # churn the caching allocator
    agent.policy_value(np.zeros((64, 4, 6, 6), dtype=np.float32))            # grows the evaluator's buffers
    assert agent.native.weights_version == version + 1
    del junk
    mcts.update_with_move(-1)
    acts2, probs2 = mcts.simulate(env, 1.0)
    assert tuple(acts) == tuple(acts2) and np.array_equal(probs, probs2)


def test_one_graph_serves_every_seed():
    """The per-search seed reaches the captured kernels through rz_tree_desc.seed_dev: successive noisy searches
    draw different noise without a new capture, and the same seed reproduces the same tree."""
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    f = SearchForest(8, 6, 4, n_playout=40, rule=L.RULE_PUCT, max_carry=0)
    ev = ClosedFormEvaluator(L.EVAL_HASH)

    def run(seed):
        f.reset_trees()
        f.run_waves(40, ev, noise_eps=0.25, noise_alpha=0.3, seed=seed)
        return f.dump_tree(3)['P'][0].copy(), f.root_stats()[0].copy()
    p1, v1 = run(1)
    p2, v2 = run(2)
    p1b, v1b = run(1)
    assert len(f._graphs) == 1
    assert not np.array_equal(p1, p2)
    assert np.array_equal(p1, p1b) and np.array_equal(v1, v1b)


def test_rollout_streams_differ_between_moves():
    """RolloutPlayer used the same random numbers for playout #v of every move and game.  The same
    position searched repeatedly by one player must not replay the same playouts: on 3x3 a random playout ends in a
    tie (0) or a win (-1 by the reference's literal rule), so the value sums tell the streams apart."""
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.mcts import RolloutPlayer
    env = GomokuEnv(3, 3)
    env.reset()
    p = RolloutPlayer(n_playout=40, seed=9)
    seen = set()
    for _ in range(6):
        p.mcts.simulate(env)
        seen.add(tuple(p.mcts._forest.root_stats()[1][0].tolist()))
        p.reset_player()
    assert len(seen) > 1


# ------------------------------------------------------------------ the reference's own training script
@pytest.mark.reference
def test_unmodified_reference_training_script_runs_on_this_package(tmp_path, monkeypatch):
    """tools/train_alphazero.py, unmodified, with rlzero.* aliased to rlzero_b200.* (needs a GPU AND the reference
    checkout, i.e. a maintainer's box): two self-play iterations incl. a policy update and an evaluation."""
    import importlib.util
    import sys
    import types
    from oracle import ref_loader
    import rlzero_b200
    import rlzero_b200.games.gomoku as gm
    import rlzero_b200.games.gomoku.alphazero_agent as ag
    import rlzero_b200.mcts.alphazero_mcts as am
    import rlzero_b200.mcts.rollout_mcts as rm
    alias = {'rlzero': rlzero_b200, 'rlzero.games': rlzero_b200.games, 'rlzero.games.gomoku': gm,
             'rlzero.games.gomoku.alphazero_agent': ag, 'rlzero.mcts': rlzero_b200.mcts,
             'rlzero.mcts.alphazero_mcts': am, 'rlzero.mcts.rollout_mcts': rm}
    for k, v in alias.items():
        monkeypatch.setitem(sys.modules, k, v)
    monkeypatch.chdir(tmp_path)
    spec = importlib.util.spec_from_file_location('ref_train_script',
                                                  os.path.join(ref_loader.REFERENCE_ROOT, 'tools', 'train_alphazero.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    np.random.seed(0)
    tp = mod.TrainPipeline()
    tp.n_playout, tp.game_batch_num, tp.check_freq, tp.pure_mcts_playout_num = 40, 2, 2, 20
    tp.mcts_player.mcts.n_playout = 40
    tp.batch_size = 16
    tp.run()
    assert len(tp.data_buffer) > tp.batch_size
    assert os.path.exists(tmp_path / 'current_policy.model')
