"""CPU tests of the round-2 boundary pieces: the seeded-noise and child-shuffle fixtures of the live reference
against the oracle restatements, and the host ``TreeNode`` against the reference's class."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from oracle import dm_oracle, pyoracle
from oracle.evaluators import make_policy_value_fn

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases(name):
    with open(os.path.join(HERE, 'golden', name)) as f:
        return json.load(f)['cases']


@pytest.mark.parametrize('case', _cases('mcts_noise.json'), ids=lambda c: c['name'])
def test_seeded_noise_fixtures_vs_the_restatement(case):
    """np.random.seed(s) + add_noise=True: the restatement draws the same Dirichlet samples in the same order as the
    live reference did (node.py:63-69) and ends with the same visits, value sums and float64 priors."""
    size, k = case['size'], case['k']
    board = pyoracle.Board(size, k)
    board.reset()
    for m in case['pre']:
        board.step(m)
    rule = pyoracle.RULE_PUCT if case['rule'] == 'puct' else pyoracle.RULE_UCT
    np.random.seed(case['seed'])
    s = pyoracle.Search(make_policy_value_fn(case['eval_id']), case['n_playout'], case['c_puct'], add_noise=True,
                        rule=rule)

    def check(st):
        assert s.root.n == st['root_N'] and float(s.root.w).hex() == st['root_W']
        assert s.root_visits(size * size).tolist() == st['visits']
        assert [float(x).hex() for x in s.root_values(size * size)] == st['W']
        pri = [float(0.0).hex()] * (size * size)
        for a, ch in s.root.children.items():
            pri[a] = float(ch.prior).hex()
        assert pri == st['prior']

    s.simulate(board, 1.0)
    check(case['stages'][0])
    for m, st in zip(case['chain'], case['stages'][1:]):
        board.step(m)
        s.update_with_move(m)
        s.simulate(board, 1.0)
        check(st)


def test_the_noise_stream_is_reproducible_from_the_seed():
    """The fixtures keep a sha1 of the draws instead of the draws: the legacy numpy stream is frozen."""
    c = _cases('mcts_noise.json')[0]
    np.random.seed(c['seed'])
    n_legal = c['size'] * c['size'] - len(c['pre'])
    first = np.random.dirichlet(0.3 * np.ones(n_legal))
    assert [float(x).hex() for x in first] == c['first_draw']


@pytest.mark.parametrize('case', _cases('dm_mcts_shuffle.json'),
                         ids=lambda c: '%dx%d_%s_seed%d' % (c['size'], c['size'], c['method'], c['seed']))
def test_dm_child_shuffle_fixtures_vs_the_restatement(case):
    """DeepMindMCTS with its RandomState seeded (real child shuffle, deepmind_mcts.py:508, and root noise from the
    same stream): the restatement reproduces the children order, their statistics and the move."""
    env = pyoracle.DMBoard(case['size'], case['k'])
    env.reset()
    for a in case['moves']:
        env.step(a)
    rs = np.random.RandomState(case['seed'])
    s = dm_oracle.DMSearch(dm_oracle.ClosedFormEvaluator(case['eval_id']), case['sims'], 2, case['method'],
                           add_exploration_noise=case['noise'], dirichlet_noise_epsilon=0.25, solve=case['solve'],
                           noise_fn=lambda n: rs.dirichlet([0.25] * n), shuffle_fn=rs.shuffle)
    root = s.search(env)
    assert root.n == case['root_n'] and float(root.w).hex() == case['root_w'] and root.outcome == case['root_outcome']
    got = [[ch.action, ch.n, float(ch.w).hex(), ch.outcome, float(ch.prior).hex()] for ch in root.children]
    assert got == case['children']
    assert root.best_child().action == case['best']


# ------------------------------------------------------------------ TreeNode (rlzero/mcts/node.py:7-184)
def _random_tree_ops(NodeCls, seed):
    """Drive a node class through a random sequence of the reference's own operations; return a digest."""
    rnd = random.Random(seed)
    np.random.seed(seed)
    root = NodeCls(None, 1.0)
    log = []
    for step in range(200):
        node = root
        path = []
        while not node.is_leaf():
            a, node = node.select(5.0 if step % 2 else 1.25)
            path.append(a)
        acts = sorted(rnd.sample(range(30), rnd.randint(1, 6)))
        if len(path) < 6:
            node.expand([(a, rnd.random()) for a in acts], add_noise=bool(step % 3 == 0))
        node.update_recursive(rnd.uniform(-1, 1))
        log.append((tuple(path), node.explore_count, float(node.total_reward).hex()))
    kids = [(a, c.explore_count, float(c.total_reward).hex(), float(c.prior).hex(), c.is_leaf(), c.is_root())
            for a, c in root._children.items()]
    scores = [float(c.uct_value(2.0)).hex() for c in root.children.values()]
    scores += [float(c.ucb_value(2.0)).hex() for c in root.children.values()]
    scores += [float(c.puct_value(2.0)).hex() for c in root.children.values() if c.explore_count > 0]
    return log, kids, scores, root.explore_count, float(root.total_reward).hex(), str(root)


@pytest.mark.reference
def test_tree_node_methods_match_the_reference_class():
    """rlzero_b200.mcts.node.TreeNode carries the reference's method set with the reference's results: select /
    expand (+ seeded noise) / update / update_recursive / uct_value / ucb_value / puct_value / is_leaf / is_root."""
    from oracle import ref_loader
    from rlzero_b200.mcts.node import TreeNode
    ref = ref_loader.load()
    for seed in (0, 1, 2):
        assert _random_tree_ops(TreeNode, seed) == _random_tree_ops(ref.TreeNode, seed)


def test_tree_node_errors_and_snapshot_methods():
    from rlzero_b200.mcts.node import TreeNode
    with pytest.raises(ValueError, match='Node has no children.'):        # node.py:38-39
        TreeNode(None, 1.0).select(5)
    root = TreeNode(None, 1.0)
    root.expand([(3, 0.5), (7, 0.25)])
    root.expand([(3, 0.9), (9, 0.125)])                                     # an existing child is kept (node.py:71-73)
    assert list(root.children) == [3, 7, 9] and root.children[3].prior == 0.5
    assert root.select(5)[0] == 3                                           # every score +inf: the first child wins
    with pytest.raises(ZeroDivisionError):
        root.children[3].puct_value(5)                                      # node.py:113 divides by explore_count
    # a device-tree snapshot answers select() like the live tree would
    snap = dict(n_nodes=1, root_N=6, root_W=-1.0,
                N=np.array([[3, -1, 2, 1]]), W=np.array([[1.0, 0.0, -1.0, 1.0]]),
                child=np.array([[-1, -1, -1, -1]]), P=np.ones((1, 4), dtype=np.float32))
    view = TreeNode.from_snapshot(snap)
    assert sorted(view._children) == [0, 2, 3] and view.explore_count == 6
    a, node = view.select(5.0)
    best = max(view._children.items(), key=lambda kv: kv[1].total_reward / kv[1].explore_count
               + 5.0 * np.sqrt(np.log(6) / kv[1].explore_count))
    assert a == best[0] and node is view._children[a]
    node.update_recursive(0.5)                                              # host copy only
    assert view.explore_count == 7 and view.total_reward == -1.5


@pytest.mark.reference
def test_reference_training_script_binds_to_this_package():
    """tools/train_alphazero.py:11-15 imports GameControl / GomokuEnv / AlphaZeroAgent / AlphaZeroPlayer /
    RolloutPlayer from rlzero.*; every call it makes into them (constructor and method, positional and keyword
    arguments, parsed from the UNMODIFIED script) must bind to the signatures of the rlzero_b200 namesakes.
    (The script itself runs for two iterations on a GPU box: tests/test_gpu_round2.py.)"""
    import ast
    import inspect
    from oracle import ref_loader
    import rlzero_b200.games.gomoku as g
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    from rlzero_b200.mcts import AlphaZeroPlayer, RolloutPlayer
    classes = dict(GameControl=g.GameControl, GomokuEnv=g.GomokuEnv, AlphaZeroAgent=AlphaZeroAgent,
                   AlphaZeroPlayer=AlphaZeroPlayer, RolloutPlayer=RolloutPlayer)
    # attribute of TrainPipeline -> class it holds, from the script's own assignments
    holder = {'board': 'GomokuEnv', 'game': 'GameControl', 'alphazero_agent': 'AlphaZeroAgent',
              'mcts_player': 'AlphaZeroPlayer'}
    src = open(os.path.join(ref_loader.REFERENCE_ROOT, 'tools', 'train_alphazero.py')).read()
    tree = ast.parse(src)
    imported = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith('rlzero'):
            imported.update(a.name for a in node.names)
    assert imported == set(classes)
    checked = 0
    for node in ast.walk(tree):
        if not isinstance(node, ast.Call):
            continue
        f = node.func
        target = None
        if isinstance(f, ast.Name) and f.id in classes:
            target = classes[f.id].__init__
            skip_self = True
        elif (isinstance(f, ast.Attribute) and isinstance(f.value, ast.Attribute)
              and isinstance(f.value.value, ast.Name) and f.value.value.id == 'self' and f.value.attr in holder):
            cls = classes[holder[f.value.attr]]
            assert hasattr(cls, f.attr) or f.attr == 'policy_value_fn', (holder[f.value.attr], f.attr)
            if f.attr == 'policy_value_fn':
                continue
            target = getattr(cls, f.attr)
            skip_self = True
        if target is None:
            continue
        sig = inspect.signature(target)
        args = [None] * (len(node.args) + (1 if skip_self else 0))
        kwargs = {kw.arg: None for kw in node.keywords if kw.arg}
        sig.bind(*args, **kwargs)            # raises TypeError if the reference's call would not bind here
        checked += 1
    assert checked >= 8
