"""Parity at BASELINE.json's full size (config 3: Gomoku 15x15, 8192 games per GPU, ResNet-10 bf16)
through size-independent properties, bit-exact spot checks of individual trees against the Python
oracle fed by the same network (batch invariance at batch 8192), and -- with a closed-form evaluator
on both sides -- bit-exact comparison of ALL 8192 trees after 800 playouts with the plain-C oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G, H, K, N_PLAYOUT = 8192, 15, 5, 40


@pytest.fixture(scope='module')
def net():
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    torch.manual_seed(0)
    return ResNetPolicyValueNet(H, n_blocks=10).cuda().eval()


def _run(net, n_games, offset, ids, noise):
    from rlzero_b200.selfplay import BatchedSelfPlay
    sp = BatchedSelfPlay(n_games, H, K, net=net, n_playout=N_PLAYOUT, c_puct=5.0, temperature=1.0,
                         add_noise=noise, global_offset=offset, seed=4321)
    sp.set_random_start_positions(global_ids=ids)
    rows0, meta0 = sp.forest.boards()
    sp.warm_up()
    for _ in range(N_PLAYOUT - 1):
        sp.step_wave()          # the last wave commits the move
    torch.cuda.synchronize()
    sp.forest.raise_faults()
    return sp, rows0, meta0


def test_full_size_properties_and_determinism(net):
    from rlzero_b200 import _lib as L
    sp, rows0, meta0 = _run(net, G, 0, None, noise=True)
    f = sp.forest
    visits = f.visits.cpu().numpy()[:, :H * H]       # root statistics the move was sampled from
    pi = f.pi.cpu().numpy()[:, :H * H]
    move = f.move.cpu().numpy()
    occ = (((rows0[:, 0] | rows0[:, 1])[:, :, None] >> np.arange(H)[None, None, :]) & 1).reshape(G, H * H)
    # fresh root: the first playout expands it, the other n-1 visit its children
    assert (visits.sum(1) == N_PLAYOUT - 1).all()
    assert (visits[occ == 1] == 0).all()
    np.testing.assert_allclose(pi.sum(1), 1.0, atol=1e-5)
    assert ((move >= 0) & (move < H * H)).all() and (occ[np.arange(G), move] == 0).all()
    assert (visits[np.arange(G), move] > 0).all()    # T = 1: only visited children can be sampled
    # the move was played: one more stone, player flipped, ply advanced
    rows1, meta1 = f.boards()
    stones0 = meta0[:, L.META_STONES]
    cont = meta1[:, L.META_EPISODE] == 0             # games that did not end (ended slots restart)
    assert (meta1[cont, L.META_STONES] == stones0[cont] + 1).all()
    assert (meta1[cont, L.META_PLAYER] == 1 - meta0[cont, L.META_PLAYER]).all()
    assert (meta1[cont, L.META_LAST_MOVE] == move[cont]).all()
    # kept subtree: root count of the re-rooted tree == visits of the chosen child
    root_n = f.root_N.cpu().numpy()
    assert (root_n[cont] == visits[np.arange(G), move][cont]).all()
    assert (f.n_nodes.cpu().numpy() <= N_PLAYOUT).all()
    # determinism: an identical second run reproduces every count and every move
    sp2, _, _ = _run(net, G, 0, None, noise=True)
    assert torch.equal(sp2.forest.visits, f.visits) and torch.equal(sp2.forest.move, f.move)
    assert torch.equal(sp2.forest.root_rows, f.root_rows)
    # shard invariance at full size: global ids 4000..4007 searched alone give the same result
    ids = np.arange(4000, 4008)
    sp3, _, _ = _run(net, 8, 4000, ids, noise=True)
    assert np.array_equal(sp3.forest.visits.cpu().numpy(), f.visits.cpu().numpy()[4000:4008])
    assert np.array_equal(sp3.forest.move.cpu().numpy(), move[4000:4008])


def test_full_batch_trees_equal_the_oracle(net):
    """Noise off: trees of the 8192-game batch are bit-identical to the oracle's sequential search
    of the same position with the same network (evaluated board by board: batch invariance)."""
    from oracle import pyoracle
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    from rlzero_b200.selfplay import BatchedSelfPlay
    sp = BatchedSelfPlay(G, H, K, net=net, n_playout=N_PLAYOUT, add_noise=False, seed=1)
    sp.set_random_start_positions()
    rows0, meta0 = sp.forest.boards()
    sp.warm_up()
    for _ in range(N_PLAYOUT - 2):
        sp.step_wave()          # stop one wave before the commit: N_PLAYOUT - 1 playouts done
    torch.cuda.synchronize()
    visits, w, has, root_n, root_w = sp.forest.root_stats()
    agent = AlphaZeroAgent(H, net=net)
    for g in (0, 1, 4095, 8191):
        b = pyoracle.Board(H, K)
        b.reset()
        rs = np.random.RandomState(1000 + g)
        for m in rs.permutation(H * H)[:(1000 + g) % 31]:
            b.step(int(m))
        assert len(b.states) == int(meta0[g, L.META_STONES])
        s = pyoracle.Search(agent.policy_value_fn, N_PLAYOUT - 1, 5)
        s.simulate(b, 1.0)
        assert visits[g].tolist() == s.root_visits(H * H).tolist(), g
        assert [float(x).hex() for x in w[g]] == [float(x).hex() for x in s.root_values(H * H)], g
        assert int(root_n[g]) == s.root.n


@pytest.mark.parametrize('rule', [0, 1])
def test_full_batch_trees_at_800_playouts_with_the_real_network_equal_the_oracle(net, rule):
    """The headline configuration itself -- 8192 games, 800 playouts, ResNet-10 on the tensor cores, noise off -- and 12
    of its trees (spread over the batch) against the oracle's sequential 799-playout search of the same position with
    the same network evaluated board by board: visits, fp64 value sums and root statistics bit for bit (UCB1 and PUCT;
    VERDICT r1 found the real-network check at this size thin: 4 trees x 39 playouts)."""
    from oracle import pyoracle
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    from rlzero_b200.selfplay import BatchedSelfPlay
    n_playout = 800
    sp = BatchedSelfPlay(G, H, K, net=net, n_playout=n_playout, add_noise=False, seed=1, rule=rule)
    sp.set_random_start_positions()
    sp.warm_up()
    for _ in range(n_playout - 2):
        sp.step_wave()          # stop one wave before the commit: n_playout - 1 playouts done
    torch.cuda.synchronize()
    sp.forest.raise_faults()
    visits, w, has, root_n, root_w = sp.forest.root_stats()
    agent = AlphaZeroAgent(H, net=net)
    for g in (0, 7, 700, 1023, 2048, 3333, 4095, 4096, 5000, 6543, 8000, 8191):
        b = pyoracle.Board(H, K)
        b.reset()
        rs = np.random.RandomState(1000 + g)
        for m in rs.permutation(H * H)[:(1000 + g) % 31]:
            b.step(int(m))
        s = pyoracle.Search(agent.policy_value_fn, n_playout - 1, 5, rule=rule)
        s.simulate(b, 1.0)
        assert visits[g].tolist() == s.root_visits(H * H).tolist(), g
        assert [float(x).hex() for x in w[g]] == [float(x).hex() for x in s.root_values(H * H)], g
        assert int(root_n[g]) == s.root.n and float(root_w[g]).hex() == float(s.root.w).hex()


@pytest.mark.parametrize('rule', [0, 1])
def test_all_8192_trees_at_800_playouts_equal_the_c_oracle(rule):
    """BASELINE config-3 size, every tree checked: 8192 games from the bench start positions, 800
    playouts each, closed-form HASH evaluator on both sides -- visit counts, fp64 value sums and root
    statistics of ALL trees are bit-identical to the plain-C restatement of the reference search
    (oracle/c/rz_oracle.c, itself pinned to the live-reference fixtures in tests/test_oracle_c.py)."""
    from oracle import build_oracle
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    n_playout = 800
    lists = []
    for g in range(G):
        rs = np.random.RandomState(1000 + g)
        lists.append([int(m) for m in rs.permutation(H * H)[:(1000 + g) % 31]])
    f = SearchForest(G, H, K, n_playout=n_playout, c_puct=5.0, rule=rule, max_carry=0)
    f.set_positions(lists)
    meta = f.boards()[1]
    from rlzero_b200 import _lib as L
    live = meta[:, L.META_STATUS] == L.ACTIVE          # a random start may already be decided: skipped by both
    f.search(ClosedFormEvaluator(EVAL_HASH))
    torch.cuda.synchronize()
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    cv, cw, crn, crw = build_oracle.search_batch(H, K, lists, n_playout, 5.0, rule, EVAL_HASH)
    assert live.sum() > G - 64
    assert np.array_equal(visits[live], cv[live])
    assert np.array_equal(w[live].view(np.int64), cw[live].view(np.int64))        # bit patterns of the fp64 sums
    assert np.array_equal(root_n[live], crn[live])
    # the root's own sum: the engine keeps the reference's sign convention (root gets -v of the leaf chain)
    assert np.array_equal(root_w[live].view(np.int64), crw[live].view(np.int64))
    assert (visits[live].sum(1) == n_playout - 1).all()


@pytest.mark.parametrize('rule,leaves', [(0, 8), (1, 16)])
def test_all_8192_trees_leaf_parallel_equal_the_c_oracle_wave(rule, leaves):
    """The opt-in leaf-parallel mode at BASELINE config-3 size, every tree checked: 8192 games, 800 playouts in waves
    of up to `leaves` playouts per tree with virtual loss -- bit-identical to the plain-C restatement of that wave
    (oracle/c/rz_oracle.c rzo_search_batch_vl; parity unpinned against the reference, which has no such mode)."""
    from oracle import build_oracle
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    n_playout = 800
    lists = []
    for g in range(G):
        rs = np.random.RandomState(1000 + g)
        lists.append([int(m) for m in rs.permutation(H * H)[:(1000 + g) % 31]])
    f = SearchForest(G, H, K, n_playout=n_playout, c_puct=5.0, rule=rule, max_carry=0, leaves_per_tree=leaves)
    f.set_positions(lists)
    meta = f.boards()[1]
    live = meta[:, L.META_STATUS] == L.ACTIVE
    f.search(ClosedFormEvaluator(EVAL_HASH))
    torch.cuda.synchronize()
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    cv, cw, crn, crw = build_oracle.search_batch_vl(H, K, lists, n_playout, 5.0, rule, EVAL_HASH, leaves, 1.0)
    assert live.sum() > G - 64
    assert np.array_equal(visits[live], cv[live])
    assert np.array_equal(w[live].view(np.int64), cw[live].view(np.int64))
    assert np.array_equal(root_n[live], crn[live]) and (root_n[live] == n_playout).all()
    assert np.array_equal(root_w[live].view(np.int64), crw[live].view(np.int64))
    assert (visits[live].sum(1) == n_playout - 1).all()


@pytest.mark.parametrize('rule,leaves', [(0, 1), (1, 1), (0, 8)])
def test_config2_all_4096_connect_four_trees_equal_the_c_oracle(rule, leaves):
    """BASELINE config-2 size, every tree checked: 4096 Connect Four games (6x7, gravity, 7 actions), 200 playouts
    each, closed-form HASH evaluator on both sides: visit counts, fp64 value sums and root statistics of ALL trees are
    bit-identical to the C oracle's search over that game (sequential = the reference's MCTS algorithm, pinned through
    tests/golden/connect4.json and tests/test_oracle_c.py; leaves = 8: the leaf-parallel wave)."""
    from oracle import build_oracle
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    G2, n_playout = 4096, 200
    lists = []
    for g in range(G2):
        rs = np.random.RandomState(1000 + g)
        lists.append([int(m) for m in rs.permutation(np.repeat(np.arange(7), 6))[:(1000 + g) % 13]])
    f = SearchForest(G2, 6, 4, n_playout=n_playout, c_puct=5.0, rule=rule, max_carry=0, board_width=7,
                     game_type=L.GAME_CONNECT4, leaves_per_tree=leaves)
    f.set_positions(lists)
    live = f.boards()[1][:, L.META_STATUS] == L.ACTIVE
    f.search(ClosedFormEvaluator(EVAL_HASH))
    torch.cuda.synchronize()
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    # the C oracle refuses a move list that runs into a finished game; search only the live ones
    idx = np.nonzero(live)[0]
    cv, cw, crn, crw = build_oracle.search_batch_c4([lists[i] for i in idx], n_playout, 5.0, rule, EVAL_HASH,
                                                    leaves_per_wave=leaves)
    assert len(idx) > G2 - 256
    assert np.array_equal(visits[idx], cv)
    assert np.array_equal(w[idx].view(np.int64), cw.view(np.int64))
    assert np.array_equal(root_n[idx], crn) and np.array_equal(root_w[idx].view(np.int64), crw.view(np.int64))


@pytest.mark.parametrize('method,solve,returns_mode,sims', [('puct', True, 1, 300), ('uct', True, 0, 200)])
def test_all_8192_deepmind_mcts_trees_equal_the_c_oracle(method, solve, returns_mode, sims):
    """The DeepMindMCTS flavour (rlzero/mcts/deepmind_mcts.py) at full width: 8192 Gomoku 15x15 positions, every tree
    checked against the C restatement of the reference class (oracle/c/rz_oracle.c dm_search_game, pinned on the CPU to
    the live-reference fixtures and to oracle.dm_oracle): root children's visit counts, value-sum bits and outcomes,
    the root's N / W / outcome, best_child."""
    from oracle import build_oracle
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    lists = []
    for g in range(G):
        rs = np.random.RandomState(1000 + g)
        lists.append([int(m) for m in rs.permutation(H * H)[:(1000 + g) % 31]])
    f = SearchForest(G, H, K, n_playout=sims, c_puct=2.0, rule=L.RULE_PUCT if method == 'puct' else L.RULE_UCT,
                     flavour=L.FLAVOUR_DEEPMIND, solve=solve, returns_mode=returns_mode, max_carry=0)
    f.set_positions(lists)
    live = f.boards()[1][:, L.META_STATUS] == L.ACTIVE
    f.search(ClosedFormEvaluator(EVAL_HASH))
    torch.cuda.synchronize()
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    A = H * H
    edge_o = f.edge_O.view(G, f.max_nodes, f.AS)[:, 0, :A].cpu().numpy()
    root_o = f.root_O.cpu().numpy()
    best, _ = f.best_child()
    idx = np.nonzero(live)[0]
    r = build_oracle.dm_search_batch(H, K, [lists[i] for i in idx], sims, 2.0, method, solve, returns_mode, EVAL_HASH)
    assert len(idx) > G - 64
    child = r['visits'] >= 0
    assert np.array_equal(has[idx], child)
    assert np.array_equal(visits[idx], np.where(child, r['visits'], 0))
    wc = np.where(child & (r['visits'] > 0), r['w'], 0.0)
    assert np.array_equal(w[idx].view(np.int64), wc.view(np.int64))
    assert np.array_equal(np.where(child, edge_o[idx], 0), r['outcome'])
    assert np.array_equal(root_n[idx], r['root_n']) and np.array_equal(root_w[idx].view(np.int64), r['root_w'].view(np.int64))
    assert np.array_equal(root_o[idx], r['root_outcome']) and np.array_equal(best[idx], r['best'])


@pytest.mark.parametrize('rule', [0, 1])
def test_all_8192_trees_after_a_committed_move_equal_the_c_oracle(rule):
    """Tree reuse at full width (update_with_move, alphazero_mcts.py:96-103 = rz_tree_advance with subtree compaction):
    search 400 playouts, play every game's most visited move keeping its subtree, search 400 more -- every re-rooted
    tree equals the C oracle's (which the Python restatement pins on the CPU, tests/test_oracle_c.py)."""
    from oracle import build_oracle
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    n_playout = 400
    lists = []
    for g in range(G):
        rs = np.random.RandomState(1000 + g)
        lists.append([int(m) for m in rs.permutation(H * H)[:(1000 + g) % 31]])
    f = SearchForest(G, H, K, n_playout=n_playout, c_puct=5.0, rule=rule, max_carry=n_playout)
    f.set_positions(lists)
    live0 = f.boards()[1][:, L.META_STATUS] == L.ACTIVE
    ev = ClosedFormEvaluator(EVAL_HASH)
    f.search(ev)
    f.raise_faults()
    visits1 = f.root_stats()[0]
    moves = np.argmax(visits1, axis=1).astype(np.int32)
    moves[~live0] = -1
    f.advance(moves, keep_subtree=True)
    f.raise_faults()
    live = live0 & (f.boards()[1][:, L.META_STATUS] == L.ACTIVE)
    f.search(ev)
    torch.cuda.synchronize()
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    idx = np.nonzero(live0)[0]
    cm, cv, cw, crn, crw = build_oracle.search_batch_reuse(H, K, [lists[i] for i in idx], n_playout, 5.0, rule, EVAL_HASH)
    assert np.array_equal(moves[idx], cm)
    sel = live[idx]
    assert sel.sum() > G - 128
    assert np.array_equal(visits[idx][sel], cv[sel])
    assert np.array_equal(w[idx][sel].view(np.int64), cw[sel].view(np.int64))
    assert np.array_equal(root_n[idx][sel], crn[sel])
    assert np.array_equal(root_w[idx][sel].view(np.int64), crw[sel].view(np.int64))
    assert (crn[~sel] == 0).all()            # the games the move ended: nothing searched on either side


def test_root_policy_and_move_sampling_at_full_width():
    """AlphaZeroPlayer.get_action's pi = softmax(log(N + 1e-10) / T) and np.random.choice(acts, p=pi)
    (alphazero_mcts.py:86-94,144-148) for all 8192 trees after an 800-playout search: pi against numpy in fp64, the
    sampled move against numpy's rule (first index whose normalised cumulative probability exceeds u) for the same
    uniforms, T -> 0 picks a most-visited child."""
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    lists = []
    for g in range(G):
        rs = np.random.RandomState(1000 + g)
        lists.append([int(m) for m in rs.permutation(H * H)[:(1000 + g) % 31]])
    f = SearchForest(G, H, K, n_playout=800)
    f.set_positions(lists)
    f.search(ClosedFormEvaluator(EVAL_HASH))
    f.raise_faults()
    A = H * H
    u = np.random.RandomState(5).random_sample(G)
    f.root_policy(temperature=1.0, u01=u)
    pi = f.pi.cpu().numpy()[:, :A].astype(np.float64)
    mv = f.move.cpu().numpy()
    visits = f.visits.cpu().numpy()[:, :A].astype(np.float64)
    has = f.root_stats()[2]
    live = has.any(1)
    assert live.sum() > G - 64
    logits = np.where(has, np.log(visits + 1e-10), -np.inf)
    with np.errstate(invalid='ignore'):                    # rows of finished games are all -inf
        p = np.exp(logits - logits.max(1, keepdims=True))
    p = np.where(live[:, None], p / np.maximum(np.nan_to_num(p).sum(1, keepdims=True), 1e-300), 0.0)
    assert np.abs(pi[live] - p[live]).max() < 1e-6
    assert np.abs(pi[live].sum(1) - 1.0).max() < 1e-5
    cdf = np.cumsum(p, axis=1)
    cdf /= np.maximum(cdf[:, -1:], 1e-300)
    want = (cdf <= u[:, None]).sum(1)                       # searchsorted(u, side='right') per row
    near = np.abs(cdf[np.arange(G), np.minimum(want, mv).clip(0, A - 1)] - u) < 1e-6   # fp32 pi on the device: a boundary case
    assert ((mv == want) | near)[live].all()
    assert (mv[live] >= 0).all() and has[np.arange(G), mv.clip(0)][live].all()
    f.root_policy(temperature=1e-3, u01=u)
    mv0 = f.move.cpu().numpy()
    assert (visits[np.arange(G), mv0.clip(0)] == visits.max(1))[live].all()


@pytest.mark.parametrize('size,k', [(15, 5), (9, 4)])
def test_8192_random_games_to_the_end_equal_the_c_oracle(size, k):
    """The game kernels at full width (GomokuEnv.step / has_a_winner / game_end_winner, gomoku_env.py:49-70,116-170,
    196-203): 8192 random games played to the end, every game ends at the same ply with the same winner as in the C
    restatement (itself pinned to pyoracle.Board and through it to the live reference's env games); the legal mask of
    every final position equals the empty squares."""
    from oracle import build_oracle
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import SearchForest
    Gn, A = 8192, size * size
    rs = np.random.RandomState(size * 7)
    moves = np.stack([rs.permutation(A) for _ in range(Gn)]).astype(np.int32)
    f = SearchForest(Gn, size, k, n_playout=2)
    end_ply = np.zeros(Gn, dtype=np.int32)
    winner = np.full(Gn, -1, dtype=np.int32)
    active = np.ones(Gn, dtype=bool)
    for t in range(A):
        f.play_moves(np.where(active, moves[:, t], -1))
        meta = f.root_meta.cpu().numpy()
        over = active & (meta[:, L.META_STATUS] != L.ACTIVE)
        end_ply[over] = t + 1
        winner[over] = meta[over, L.META_WINNER]
        active &= ~over
        if not active.any():
            break
    f.raise_faults()
    assert not active.any()
    c_end, c_win, c_ended = build_oracle.replay_games(size, k, moves)
    assert c_ended.all() and np.array_equal(end_ply, c_end) and np.array_equal(winner, c_win)
    assert (winner == -1).sum() >= 0 and set(np.unique(winner)).issubset({-1, 0, 1})
    # final positions: stones on the board = plies played; legal mask = empty squares
    import ctypes as C
    mask = torch.zeros(Gn, A, dtype=torch.uint8, device='cuda')
    L.check(f.lib.rz_gomoku_legal_mask(C.byref(f.gdesc), L.ptr(f.root_rows), L.ptr(mask), Gn, L.stream_ptr()), 'legal')
    legal = mask.cpu().numpy().astype(bool)
    played = np.zeros((Gn, A), dtype=bool)
    for g in range(0, Gn, 257):
        played[g, moves[g, :end_ply[g]]] = True
        assert np.array_equal(legal[g], ~played[g]), g
    assert (legal.sum(1) == A - end_ply).all()


# ------------------------------------------------------------------ config 4: Go 19x19 at full width
@pytest.mark.parametrize('rule', [0, 1])
def test_config4_all_8192_go_trees_at_800_playouts_equal_the_c_oracle(rule):
    """BASELINE config-4 size, every tree checked: 8192 Go positions on 19x19 (komi 7.5, move cap 722) reached by up
    to 30 random legal moves, 800 playouts each with the closed-form HASH evaluator on both sides: visit counts, fp64
    value-sum bit patterns and root statistics of ALL trees equal the plain-C Go oracle (oracle/c/rz_go_oracle.c,
    pinned on the CPU to oracle/go_oracle.py by tests/test_go_oracle_c.py; parity with the reference's own engine
    remains unpinned -- pettingzoo's go_base is absent).  The positions are replayed on the device through
    rz_go_step, so a rules disagreement in the prefix would already fault."""
    from oracle import build_oracle
    from oracle.evaluators import EVAL_HASH
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator, SearchForest
    n, komi, cap, n_playout = 19, 7.5, 722, 800
    games = build_oracle.go_random_games(G, n, [(1000 + g) % 31 for g in range(G)], komi, cap, seed=5)
    lists = [games['moves'][g][:games['played'][g]].tolist() for g in range(G)]
    f = SearchForest(G, n, 1, n_playout=n_playout, c_puct=5.0, rule=rule, max_carry=0, game_type=L.GAME_GO, komi=komi,
                     max_moves=cap)
    f.set_positions(lists)
    f.search(ClosedFormEvaluator(EVAL_HASH))
    torch.cuda.synchronize()
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    cv, cw, crn, crw = build_oracle.go_search_batch(n, lists, n_playout, komi, cap, 5.0, rule, EVAL_HASH)
    live = games['over'] == 0          # a random prefix may end with two passes: the engine skips finished games
    assert live.sum() > G - 256
    assert np.array_equal(visits[live], cv[live])
    assert np.array_equal(w[live].view(np.int64), cw[live].view(np.int64))
    assert np.array_equal(root_n[live], crn[live]) and np.array_equal(root_w[live].view(np.int64), crw[live].view(np.int64))
    assert (visits[live].sum(1) == n_playout - 1).all()
    if rule == 0:
        assert (visits[live][:, n * n] > 0).all()       # UCB1 visits every child once: the pass is one of them


def test_config4_8192_random_go_games_to_the_end_equal_the_c_oracle():
    """8192 random Go games of up to 420 plies (captures, ko, suicide bans, passes, two-pass ends, the move cap) played by
    the C oracle and replayed on the device (rz_go_step): final boards, ko squares, side to move, game-over flags,
    Tromp-Taylor scores and the legal-move masks of every game."""
    from oracle import build_oracle
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.go import GoBoards
    n, komi, cap = 19, 7.5, 400
    plies = [20 + (37 * g) % 401 for g in range(G)]
    games = build_oracle.go_random_games(G, n, plies, komi, cap, seed=9)
    gb = GoBoards(G, n, komi, max_moves=cap)
    mx = int(games['played'].max())
    for t in range(mx):
        gb.step(np.where(t < games['played'], games['moves'][:, min(t, games['moves'].shape[1] - 1)], -1))
    assert not gb.faults().any()
    assert np.array_equal(gb.boards().reshape(G, -1), games['cell'])
    meta = gb.meta.cpu().numpy()
    assert np.array_equal(meta[:, L.META_KO], games['ko'])
    assert np.array_equal(meta[:, L.META_PLAYER], games['to_play'])
    assert np.array_equal((meta[:, L.META_STATUS] != L.ACTIVE).astype(np.int32), games['over'])
    score, result = (x.cpu().numpy() for x in gb.score())
    assert np.array_equal(score, games['score'])
    assert np.array_equal(gb.legal_mask().cpu().numpy().astype(np.int8), games['legal'])
    assert games['over'].sum() > 100 and (games['played'] < np.asarray(plies)).sum() > 100    # many games really ended
