"""GPU tests of the opt-in leaf-parallel search mode (virtual loss; rz_tree_desc.leaves_per_tree).

The reference is strictly sequential, so this mode is NOT part of reference parity (parity unpinned: the reference
has no virtual loss).  What is checked: the kernels implement exactly the wave defined in include/rlzero_b200.h --
bit for bit against its restatement ``oracle.pyoracle.Search.wave`` (visit counts, fp64 value sums, tree size, tree
reuse) -- and K = 1 stays the reference's sequential search."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_positions(size, k, G, seed):
    from oracle import pyoracle
    rs = np.random.RandomState(seed)
    move_lists, boards = [], []
    while len(move_lists) < G:
        b = pyoracle.Board(size, k)
        b.reset()
        moves = []
        for _ in range(rs.randint(0, size * size - 1)):
            a = int(b.legal[rs.randint(len(b.legal))])
            b.step(a)
            moves.append(a)
            if b.game_end_winner()[0]:
                break
        if b.game_end_winner()[0]:
            continue
        move_lists.append(moves)
        boards.append(b)
    return move_lists, boards


@pytest.mark.parametrize('size,k,n_playout,rule,K,vl', [(3, 3, 60, 'uct', 4, 1.0), (6, 4, 121, 'uct', 8, 1.0),
                                                        (9, 5, 150, 'puct', 16, 1.0), (15, 5, 200, 'uct', 32, 1.0),
                                                        (6, 4, 97, 'puct', 3, 0.5), (15, 5, 300, 'uct', 64, 2.0),
                                                        (5, 4, 400, 'uct', 8, 1.0)])
def test_leaf_parallel_waves_match_the_oracle_wave(size, k, n_playout, rule, K, vl):
    from oracle import pyoracle
    from oracle.evaluators import EVAL_HASH, make_policy_value_fn
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    G = 12
    move_lists, boards = _random_positions(size, k, G, size * 1000 + n_playout + K)
    r = L.RULE_PUCT if rule == 'puct' else L.RULE_UCT
    f = SearchForest(G, size, k, n_playout=n_playout, c_puct=2.5, rule=r, max_carry=n_playout, leaves_per_tree=K,
                     virtual_loss=vl)
    assert f.n_leaves == G * K and f.prior.shape[0] == G * K
    f.set_positions(move_lists)
    f.search(ClosedFormEvaluator(EVAL_HASH))
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    fn = make_policy_value_fn(EVAL_HASH)
    orule = pyoracle.RULE_PUCT if rule == 'puct' else pyoracle.RULE_UCT
    A = size * size
    searches = []
    for g in range(G):
        s = pyoracle.Search(fn, n_playout, 2.5, rule=orule, leaves_per_wave=K, virtual_loss=vl)
        s.simulate(boards[g], 1.0)
        searches.append(s)
        assert int(root_n[g]) == s.root.n == n_playout, g
        assert float(root_w[g]).hex() == float(s.root.w).hex(), g
        assert visits[g].tolist() == s.root_visits(A).tolist(), g
        assert [float(x).hex() for x in w[g]] == [float(x).hex() for x in s.root_values(A)], g
    # every virtual statistic was taken off again: visit counts of the root's children sum to root_N - 1 (+ the
    # visits of a re-used subtree: none here)
    assert (visits.sum(1) == n_playout - 1).all()
    # tree reuse: play the most visited move, search again (the root is expanded: full waves from the start)
    moves = [int(np.argmax(visits[g])) for g in range(G)]
    f.advance(moves, keep_subtree=True)
    f.raise_faults()
    f.search(ClosedFormEvaluator(EVAL_HASH))
    f.raise_faults()
    visits2, w2, _, root_n2, root_w2 = f.root_stats()
    for g in range(G):
        s = searches[g]
        b = boards[g]
        b.step(moves[g])
        if b.game_end_winner()[0]:
            continue
        s.update_with_move(moves[g])
        s.simulate(b, 1.0)
        assert int(root_n2[g]) == s.root.n, g
        assert visits2[g].tolist() == s.root_visits(A).tolist(), g
        assert [float(x).hex() for x in w2[g]] == [float(x).hex() for x in s.root_values(A)], g


def test_leaves_per_tree_1_is_the_sequential_reference_search():
    """K = 1 goes through the same kernels' parity branch: identical to a forest built without the option."""
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    move_lists, _ = _random_positions(9, 5, 6, 7)
    out = []
    for kw in (dict(), dict(leaves_per_tree=1)):
        f = SearchForest(6, 9, 5, n_playout=150, c_puct=5.0, **kw)
        f.set_positions(move_lists)
        f.search(ClosedFormEvaluator(EVAL_HASH))
        out.append(f.root_stats())
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_single_game_api_with_leaves_per_wave():
    """AlphaZeroPlayer(leaves_per_wave=K) behind the reference API with the tensor-core network: legal moves, pi
    sums to 1 over n_playout - 1 visits, equal to the oracle wave fed by the same network, and the self-play
    subtree is kept."""
    import torch
    from oracle import pyoracle
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    from rlzero_b200.mcts import AlphaZeroMCTS, AlphaZeroPlayer
    torch.manual_seed(3)
    agent = AlphaZeroAgent(6, net=ResNetPolicyValueNet(6, n_blocks=2))
    agent.policy_value_net.eval()
    env = GomokuEnv(6, 4)
    env.reset()
    env.step(14)
    env.step(15)
    mcts = AlphaZeroMCTS(agent.policy_value_fn, n_playout=130, c_puct=5, leaves_per_wave=16)
    acts, probs = mcts.simulate(env, 1.0)
    b = pyoracle.Board(6, 4)
    b.reset()
    b.step(14)
    b.step(15)
    s = pyoracle.Search(agent.policy_value_fn, 130, 5, leaves_per_wave=16)
    acts2, probs2 = s.simulate(b, 1.0)
    assert tuple(acts) == tuple(acts2) and np.array_equal(probs, probs2)
    assert mcts._forest.root_stats()[0][0].sum() == 129
    np.random.seed(0)
    player = AlphaZeroPlayer(agent.policy_value_fn, n_playout=64, c_puct=5, is_selfplay=True, leaves_per_wave=8)
    for _ in range(4):
        move, pi = player.get_action(env, temperature=1.0, return_prob=True)
        assert move in env.leagel_actions() and abs(pi.sum() - 1.0) < 1e-9
        env.step(move)
        if env.game_end_winner()[0]:
            break


@pytest.mark.parametrize('n,n_playout,plies,K', [(5, 150, 12, 8), (9, 200, 40, 16), (3, 90, 4, 4)])
def test_leaf_parallel_on_go_matches_the_oracle_wave(n, n_playout, plies, K):
    """The same wave over the Go rules (captures, ko, passes and two-pass terminal leaves inside the tree)."""
    from oracle import pyoracle
    from oracle.evaluators import EVAL_HASH, make_policy_value_fn
    from oracle.go_oracle import GoSearchBoard
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    G, komi = 6, 2.5
    positions = []
    for g in range(G):
        rs = np.random.RandomState(17 * g + n)
        b = GoSearchBoard(n, komi, 0)
        moves = []
        for _ in range((plies * (g + 1)) // G):
            legal = b.leagel_actions()
            a = legal[rs.randint(len(legal) - 1)] if len(legal) > 1 and rs.rand() > 0.03 else n * n
            b.step(a)
            moves.append(a)
            if b.game_end_winner()[0]:
                b.reset()
                moves = []
        positions.append((b, moves))
    f = SearchForest(G, n, 1, n_playout=n_playout, game_type=L.GAME_GO, komi=komi, leaves_per_tree=K)
    f.set_positions([m for _, m in positions])
    f.search(ClosedFormEvaluator(EVAL_HASH))
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    A = n * n + 1
    for g, (board, _) in enumerate(positions):
        s = pyoracle.Search(make_policy_value_fn(EVAL_HASH), n_playout, 5, leaves_per_wave=K)
        s.simulate(board, 1.0)
        assert np.array_equal(visits[g], s.root_visits(A)), g
        assert np.array_equal(w[g], s.root_values(A)), g
        assert root_n[g] == s.root.n and root_w[g] == s.root.w


def test_leaf_parallel_on_connect_four_matches_the_oracle_wave():
    from oracle import pyoracle
    from oracle.evaluators import EVAL_HASH, make_policy_value_fn
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    G, n_playout, K = 8, 300, 8
    boards, move_lists = [], []
    for g in range(G):
        rs = np.random.RandomState(1000 + g)
        seq = [int(m) for m in rs.permutation(np.repeat(np.arange(7), 6))[:g * 2]]
        b = pyoracle.ConnectFourBoard()
        b.reset()
        ok = []
        for m in seq:
            b.step(m)
            ok.append(m)
            if b.game_end_winner()[0]:
                b.reset()
                ok = []
        boards.append(b)
        move_lists.append(ok)
    f = SearchForest(G, 6, 4, n_playout=n_playout, board_width=7, game_type=L.GAME_CONNECT4, leaves_per_tree=K)
    f.set_positions(move_lists)
    f.search(ClosedFormEvaluator(EVAL_HASH))
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    for g in range(G):
        s = pyoracle.Search(make_policy_value_fn(EVAL_HASH), n_playout, 5, leaves_per_wave=K)
        s.simulate(boards[g], 1.0)
        assert visits[g].tolist() == s.root_visits(7).tolist(), g
        assert [float(x).hex() for x in w[g]] == [float(x).hex() for x in s.root_values(7)], g


def test_leaf_parallel_rejects_the_deepmind_flavour():
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import SearchForest
    with pytest.raises(ValueError):
        SearchForest(1, 6, 4, n_playout=10, flavour=L.FLAVOUR_DEEPMIND, leaves_per_tree=4)


def test_batched_selfplay_with_leaf_parallel_waves():
    """BatchedSelfPlay(leaves_per_tree=K): a move is 1 + ceil((n-1)/K) waves, every game gets exactly n_playout
    playouts per move (the root's child visits sum to n_playout - 1 on a fresh tree), games finish, drain works."""
    import torch
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    from rlzero_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    net = ResNetPolicyValueNet(6, n_blocks=1).cuda().eval()
    G, n, K = 24, 50, 8
    sp = BatchedSelfPlay(G, 6, 4, net=net, n_playout=n, add_noise=True, seed=3, leaves_per_tree=K)
    assert sp.waves_per_move == 1 + (n - 1 + K - 1) // K and sp.evaluator.max_batch >= G * K
    sp.warm_up()
    for _ in range(sp.waves_per_move - 2):
        sp.step_wave()
    # one wave before the commit: n_playout - (last wave's share) visits so far; after it exactly n_playout
    if sp._graph is not None:
        sp._graph.replay()
    else:
        sp._wave()
    torch.cuda.synchronize()
    sp.forest.raise_faults()
    visits, _, _, root_n, _ = sp.forest.root_stats()
    assert (root_n == n).all() and (visits.sum(1) == n - 1).all()
    sp.waves_in_move += 1
    sp.commit_move()
    sp.play(40)
    sp.forest.raise_faults()
    st = sp.stats()
    assert st['games_done'] >= G
    states, pis, zs, info = sp.drain()
    assert len(zs) > 0 and set(np.unique(zs)).issubset({-1.0, 0.0, 1.0})
    assert np.allclose(pis.sum(1), 1.0, atol=1e-5)
    # the host-buffer API in the same mode
    rows, meta = sp.forest.boards()
    meta = meta.copy()
    from rlzero_b200 import _lib as L
    meta[:, L.META_STATUS] = L.ACTIVE
    moves, pi, v = sp.get_actions(rows, meta)
    assert (v.sum(1) == n - 1).all() and np.allclose(pi.sum(1), 1.0, atol=1e-5)
