"""Search parity on Go: the batched CUDA search (SearchForest, RZ_GAME_GO) against the reference's
AlphaZeroMCTS algorithm (oracle.pyoracle.Search) run over the Go oracle.  With the same closed-form
evaluator on both sides, visit counts, value sums and tree reuse must be bit-identical -- including
passes, captures inside the tree, ko and two-pass terminal leaves."""
import numpy as np
import pytest

from oracle import pyoracle
from oracle.evaluators import EVAL_HASH, EVAL_ZERO, make_policy_value_fn
from oracle.go_oracle import GoSearchBoard

pytestmark = pytest.mark.gpu


def _forest(G, n, n_playout, komi=2.5, rule=0, max_moves=0, **kw):
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import SearchForest
    return SearchForest(G, n, 1, n_playout=n_playout, rule=rule, game_type=L.GAME_GO, komi=komi,
                        max_moves=max_moves, **kw)


def _random_position(n, plies, seed, komi, max_moves=0):
    rs = np.random.RandomState(seed)
    b = GoSearchBoard(n, komi, max_moves)
    moves = []
    for _ in range(plies):
        legal = b.leagel_actions()
        a = legal[rs.randint(len(legal) - 1)] if len(legal) > 1 and rs.rand() > 0.03 else n * n
        b.step(a)
        moves.append(a)
        if b.game_end_winner()[0]:
            b.reset()
            moves = []
    return b, moves


@pytest.mark.parametrize('n,n_playout,plies,eval_id,rule', [
    (3, 80, 4, EVAL_HASH, 0), (5, 120, 12, EVAL_HASH, 0), (5, 150, 30, EVAL_HASH, 1), (5, 60, 20, EVAL_ZERO, 0),
    (9, 200, 40, EVAL_HASH, 0), (9, 120, 90, EVAL_HASH, 1)])
def test_visits_match_the_oracle(n, n_playout, plies, eval_id, rule):
    from rlzero_b200.engine import ClosedFormEvaluator
    G, komi = 6, 2.5
    positions = [_random_position(n, (plies * (g + 1)) // G, 17 * g + n, komi) for g in range(G)]
    f = _forest(G, n, n_playout, komi, rule=rule)
    f.set_positions([m for _, m in positions])
    f.search(ClosedFormEvaluator(eval_id))
    f.raise_faults()
    visits, w, has, root_n, root_w = f.root_stats()
    A = n * n + 1
    for g, (board, _) in enumerate(positions):
        s = pyoracle.Search(make_policy_value_fn(eval_id), n_playout, 5, rule=rule)
        s.simulate(board, 1.0)
        assert np.array_equal(visits[g], s.root_visits(A)), g
        assert np.array_equal(w[g], s.root_values(A)), g
        assert root_n[g] == s.root.n and root_w[g] == s.root.w
        assert sorted(np.nonzero(has[g])[0]) == sorted(s.root.children.keys())


def test_tree_reuse_and_passes_to_the_end():
    """Greedy self-play of one 3x3 game with tree reuse: every move's visit vector matches, through
    captures, passes and the two-pass end; the winner matches Tromp-Taylor scoring of the oracle."""
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator
    n, n_playout, komi = 3, 60, 0.5
    f = _forest(1, n, n_playout, komi, max_carry=n_playout)
    board = GoSearchBoard(n, komi)
    s = pyoracle.Search(make_policy_value_fn(EVAL_HASH), n_playout, 5)
    ev = ClosedFormEvaluator(EVAL_HASH)
    for ply in range(60):
        f.search(ev)
        f.raise_faults()
        visits = f.root_stats()[0][0]
        s.simulate(board, 1.0)
        want = s.root_visits(n * n + 1)
        assert np.array_equal(visits, want), ply
        move = int(np.argmax(want))
        f.advance([move], keep_subtree=True)
        board.step(move)
        s.update_with_move(move)
        end, winner = board.game_end_winner()
        meta = f.boards()[1][0]
        assert (meta[L.META_STATUS] != L.ACTIVE) == end
        if end:
            assert meta[L.META_WINNER] == winner
            break
    assert end


def test_move_cap_scores_the_game():
    from rlzero_b200 import _lib as L
    from rlzero_b200.engine import ClosedFormEvaluator
    n, cap = 5, 6
    f = _forest(2, n, 40, 0.5, max_moves=cap)
    boards = [GoSearchBoard(n, 0.5, cap), GoSearchBoard(n, 0.5, cap)]
    for g, mv in enumerate([[0, 1, 2], [12, 13, 7, 8]]):
        for a in mv:
            boards[g].step(a)
    f.set_positions([[0, 1, 2], [12, 13, 7, 8]])
    f.search(ClosedFormEvaluator(EVAL_HASH))
    f.raise_faults()
    visits = f.root_stats()[0]
    for g in range(2):
        s = pyoracle.Search(make_policy_value_fn(EVAL_HASH), 40, 5)
        s.simulate(boards[g], 1.0)
        assert np.array_equal(visits[g], s.root_visits(n * n + 1)), g


# ------------------------------------------------------------------ network + self-play on Go
def _go_net(n, blocks, seed=0):
    import torch
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    torch.manual_seed(seed)
    net = ResNetPolicyValueNet(n, n_blocks=blocks, n_actions=n * n + 1, in_planes=17).cuda().eval()
    # random-init BatchNorm statistics that are not the identity
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    return net


@pytest.mark.parametrize('n,blocks,games', [(9, 2, 12), (19, 3, 5)])
def test_go_net_forward_vs_torch(n, blocks, games):
    """Fused GoEnv.observe + stem + trunk + heads on Go positions against the fp32 PyTorch forward of
    the same module on the oracle's 17-plane observation: probabilities and value within 1e-3 (bf16)."""
    import torch
    from oracle.go_oracle import GoEnvOracle
    from rlzero_b200.games.go import GoBoards
    from rlzero_b200.games.gomoku.policy_value_net import NativeForward
    net = _go_net(n, blocks)
    rs = np.random.RandomState(3)
    gb = GoBoards(games, n, 7.5)
    envs = [GoEnvOracle(n, 7.5) for _ in range(games)]
    for e in envs:
        e.reset()
    for t in range(40):
        acts = []
        for g, e in enumerate(envs):
            if t >= 5 * g + 3 or e.is_terminal():
                acts.append(-1)
                continue
            legal = list(e.legal_actions())
            a = int(legal[rs.randint(len(legal) - 1)]) if (len(legal) > 1 and rs.rand() > 0.05) else n * n
            e.step(a)
            acts.append(a)
        gb.step(acts)
    obs = np.stack([e.observe(e.agent_selection)['observation'].transpose(2, 0, 1) for e in envs]).astype(np.float32)
    nf = NativeForward(net, max_batch=games)
    logp, v = nf.forward_boards(gb.rows, gb.meta, games, hist=gb.hist)
    torch.cuda.synchronize()
    logp, v = logp[:, :n * n + 1].cpu(), v.cpu()
    with torch.no_grad():
        lt, vt = net(torch.from_numpy(obs).cuda())
    lt, vt = lt.cpu(), vt.cpu().reshape(-1)
    assert (logp.exp() - lt.exp()).abs().max().item() < 1e-3
    assert (v - vt).abs().max().item() < 1e-3
    assert torch.allclose(logp.exp().sum(1), torch.ones(games), atol=1e-4)
    # the plane-fed entry (AlphaZeroAgent.policy_value path) runs the same kernels on the same values
    l2, v2 = nf.forward_planes(torch.from_numpy(obs))
    assert torch.equal(l2.cpu(), logp) and torch.equal(v2.cpu(), v)


def test_go_selfplay_runs_whole_games_and_is_shard_invariant():
    """Batched Go self-play with the tensor-core net: games finish (two passes or the move cap), the
    winners are what Tromp-Taylor scoring says, and sampled moves do not depend on the sharding."""
    import torch
    from rlzero_b200 import _lib as L
    from rlzero_b200.selfplay import BatchedSelfPlay
    n = 5
    net = _go_net(n, 1, seed=2)

    def run(G, offset, ids, n_moves):
        sp = BatchedSelfPlay(G, n, 1, net=net, n_playout=16, temperature=1.0, add_noise=True, global_offset=offset,
                             seed=31, game_type=L.GAME_GO, komi=0.5, max_moves=40, ring_capacity=4096)
        sp.set_random_start_positions(global_ids=ids, max_random_moves=5)
        moves = []
        for _ in range(n_moves):
            sp.play(1)
            torch.cuda.synchronize()
            moves.append(sp.forest.move.cpu().numpy().copy())
        sp.forest.raise_faults()
        return np.stack(moves), sp

    ids = np.arange(6)
    m_all, sp = run(6, 0, ids, 50)
    m_lo, _ = run(3, 0, ids[:3], 50)
    m_hi, _ = run(3, 3, ids[3:], 50)
    assert np.array_equal(m_all[:, :3], m_lo) and np.array_equal(m_all[:, 3:], m_hi)
    st = sp.stats()
    assert st['games_done'] >= 6                       # the 40-move cap guarantees finished episodes
    out = sp.forest.drain_trajectories()
    info = out['info']
    assert len(info) == st['plies_done'] and set(np.unique(info[:, 2])).issubset({-1, 1})
    assert np.allclose(out['pi'].sum(1), 1.0, atol=1e-5) and out['pi'].shape[1] == n * n + 1


def test_go_drain_rebuilds_the_17_plane_observation():
    """Drained Go training tuples: the 17 planes rebuilt from the per-ply boards equal GoEnv.observe
    of the oracle replaying the same episode (moves are the next ply's last_move)."""
    import torch
    from rlzero_b200 import _lib as L
    from rlzero_b200.selfplay import BatchedSelfPlay
    from oracle.go_oracle import GoEnvOracle
    n = 5
    net = _go_net(n, 1, seed=4)
    sp = BatchedSelfPlay(4, n, 1, net=net, n_playout=12, temperature=1.0, add_noise=True, seed=5,
                         game_type=L.GAME_GO, komi=0.5, max_moves=30, ring_capacity=4096)
    sp.play(64)
    torch.cuda.synchronize()
    sp.forest.raise_faults()
    states, pis, zs, info = sp.drain()
    assert states.shape[1:] == (17, n, n) and len(states) == len(pis) == len(zs) > 0
    checked = 0
    keys = sorted(set((int(a), int(b)) for a, b in info[:, 3:5]))
    for slot, ep in keys:
        sel = np.nonzero((info[:, 3] == slot) & (info[:, 4] == ep))[0]
        assert list(info[sel, 5]) == list(range(len(sel)))          # complete episode, plies in order
        env = GoEnvOracle(n, 0.5)
        env.reset()
        for j, i in enumerate(sel):
            want = env.observe(env.agent_selection)['observation'].transpose(2, 0, 1).astype(np.float32)
            assert np.array_equal(states[i], want), (slot, ep, j)
            assert info[i, 0] == env.current_player()
            if j + 1 < len(sel):
                env.step(int(info[sel[j + 1], 1]))
            checked += 1
    assert checked == len(states)
