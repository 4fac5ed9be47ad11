"""MuZero (BASELINE config 5) on the GPU.  The reference has no MuZero code, so PARITY IS UNPINNED:
the tree kernels are checked bit-exactly against oracle/muzero_oracle.py (the paper's pseudocode, two-
player convention) on replayed network outputs, and the three networks against their fp32 PyTorch forward."""
import numpy as np
import pytest
import torch

from oracle import muzero_oracle, pyoracle

pytestmark = pytest.mark.gpu


def _net(size, rb, db, seed=0):
    from rlzero_b200.muzero import MuZeroNet
    torch.manual_seed(seed)
    net = MuZeroNet(size, repr_blocks=rb, dyn_blocks=db).cuda().eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    return net


def _positions(G, size, k, seed):
    """G random non-terminal Gomoku positions as device records + the oracle boards."""
    from rlzero_b200.engine import SearchForest
    rs = np.random.RandomState(seed)
    lists, boards = [], []
    for g in range(G):
        while True:
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
                break
        lists.append(mv)
        boards.append(b)
    f = SearchForest(G, size, k, n_playout=4)
    f.set_positions(lists)
    return f, boards


@pytest.mark.parametrize('size,rb,db,n', [(6, 2, 1, 9), (9, 1, 2, 5), (15, 2, 2, 4), (6, 0, 0, 7)])
def test_networks_match_torch_fp32(size, rb, db, n):
    """h, g (with the action plane in the last channel) and f on the tensor cores vs fp32 PyTorch:
    hidden states within bf16 rounding, probabilities and values within 1e-3."""
    from rlzero_b200.muzero import MuZeroNative
    net = _net(size, rb, db)
    f, boards = _positions(n, size, min(5, size), 3)
    nat = MuZeroNative(net, n, 3, n_in_row=min(5, size))
    nat.representation(f.root_rows, f.root_meta, 0)
    logp0, v0 = (x.clone() for x in nat.prediction(0))
    obs = torch.from_numpy(np.stack([b.current_state() for b in boards])).float().cuda()
    with torch.no_grad():
        s_t, lt, vt = net.initial_inference(obs)
    s0 = nat.hidden_state(0)
    scale = s_t.abs().max().item()
    assert (s0 - s_t).abs().max().item() <= 3e-2 * max(scale, 1.0)
    A = size * size
    assert (logp0[:, :A].exp() - lt.exp()).abs().max().item() < 1e-3
    assert (v0 - vt.reshape(-1)).abs().max().item() < 1e-3
    # one recurrent step from the DEVICE hidden state (so only this step's error is measured)
    rs = np.random.RandomState(1)
    action = torch.from_numpy(rs.randint(0, A, size=n).astype(np.int32)).cuda()
    parent = torch.zeros(n, dtype=torch.int32, device='cuda')
    nat.dynamics(parent, action, 1)
    logp1, v1 = nat.prediction(1)
    with torch.no_grad():
        s1_t, l1t, v1t = net.recurrent_inference(s0, action)
    s1 = nat.hidden_state(1)
    assert (s1 - s1_t).abs().max().item() <= 3e-2 * max(s1_t.abs().max().item(), 1.0)
    assert (logp1[:, :A].exp() - l1t.exp()).abs().max().item() < 1e-3
    assert (v1 - v1t.reshape(-1)).abs().max().item() < 1e-3
    # the gather picks the right parent slot per tree: mixed parents 0 / 1
    parent2 = torch.from_numpy((np.arange(n) % 2).astype(np.int32)).cuda()
    nat.dynamics(parent2, action, 2)
    s2 = nat.hidden_state(2)
    with torch.no_grad():
        want = net.dynamics(torch.where((parent2 == 0)[:, None, None, None], s0, s1), action)
    assert (s2 - want).abs().max().item() <= 3e-2 * max(want.abs().max().item(), 1.0)


@pytest.mark.parametrize('size,sims,noise,bounds', [(6, 50, False, None), (6, 80, True, None), (9, 50, True, (-1, 1)),
                                                    (3, 30, False, None)])
def test_search_matches_the_pseudocode_oracle(size, sims, noise, bounds):
    """Tree statistics after run_mcts are bit-identical to the pseudocode restatement fed with the
    device's own network outputs (priors as stored, values as produced), tree by tree."""
    from rlzero_b200 import _lib as L
    from rlzero_b200.muzero import MuZeroConfig, MuZeroSearch
    G, k = 6, min(5, size)
    net = _net(size, 1, 1, seed=4)
    f, boards = _positions(G, size, k, 11)
    cfg = MuZeroConfig(num_simulations=sims, known_bounds=bounds)
    ms = MuZeroSearch(G, net, cfg, n_in_row=k, seed=9)
    legal = torch.zeros(G, size * size, dtype=torch.uint8, device='cuda')
    for g, b in enumerate(boards):
        legal[g, torch.tensor(b.leagel_actions())] = 1
    # eager run, logging what every simulation fed to expand/backup
    ms.root(f.root_rows, f.root_meta, legal, add_noise=noise, move_id=3)
    A, AS = ms.A, ms.AS
    root_P = ms.edge_P.view(G, ms.max_nodes, AS)[:, 0, :A].cpu().numpy()
    root_N = ms.edge_N.view(G, ms.max_nodes, AS)[:, 0, :A].cpu().numpy()
    log_P, log_v, log_parent, log_action = [], [], [], []
    for i in range(sims):
        ms.simulate(i)
        log_P.append(ms.edge_P.view(G, ms.max_nodes, AS)[:, i + 1, :A].cpu().numpy())
        log_v.append(ms.native.value.cpu().numpy().copy())
        log_parent.append(ms.leaf_parent.cpu().numpy().copy())
        log_action.append(ms.leaf_action.cpu().numpy().copy())
    ms.raise_faults()
    for g, b in enumerate(boards):
        pri = {int(a): float(root_P[g, a]) for a in range(A) if root_N[g, a] >= 0}
        assert sorted(pri) == sorted(b.leagel_actions())
        assert abs(sum(pri.values()) - 1.0) < 1e-5

        def recurrent(sim, parent_id, action, g=g):
            # the oracle must ask for exactly the node the device expanded
            assert parent_id == log_parent[sim][g] and action == log_action[sim][g], (g, sim)
            return [float(x) for x in log_P[sim][g]], float(log_v[sim][g])

        root, stats, nodes, trace = muzero_oracle.run_mcts(cfg, b.current_player(), pri, recurrent, A)
        d = ms.dump_tree(g)
        assert d['n_nodes'] == sims + 1 and d['root_N'] == root.visit_count == sims
        assert d['root_W'] == root.value_sum
        assert d['mm_min'] == stats.minimum and d['mm_max'] == stats.maximum
        for node in nodes:
            for a, ch in node.children.items():
                assert d['N'][node.node_id][a] == ch.visit_count
                if ch.visit_count:
                    assert d['W'][node.node_id][a] == ch.value_sum
                assert d['child'][node.node_id][a] == ch.node_id
    # the captured-graph path gives the same trees (noise keyed by the same counters)
    before = ms.edge_N.clone()
    ids = torch.full((G,), 3, dtype=torch.int32, device='cuda')
    ms.run(f.root_rows, f.root_meta, legal, add_noise=noise, move_ids=ids)
    ms.run(f.root_rows, f.root_meta, legal, add_noise=noise, move_ids=ids)
    torch.cuda.synchronize()
    assert torch.equal(before.view(G, ms.max_nodes, AS)[:, :, :A], ms.edge_N.view(G, ms.max_nodes, AS)[:, :, :A])
    assert (ms.root_visits().sum(1) == sims).all()


def test_selfplay_loop_plays_legal_moves_and_finishes_games():
    from rlzero_b200.muzero import BatchedMuZeroSelfPlay, MuZeroConfig
    net = _net(6, 1, 1, seed=2)
    sp = BatchedMuZeroSelfPlay(16, 6, 4, net=net, config=MuZeroConfig(num_simulations=20), seed=5)
    for _ in range(40):
        sp.play_move()
    torch.cuda.synchronize()
    sp.search.raise_faults()
    from rlzero_b200 import _lib as L
    assert not (sp.meta[:, L.META_FAULT].cpu().numpy() & L.FAULT_ILLEGAL_MOVE).any()
    assert sp.games_done >= 16 and sp.moves_played == 40


@pytest.mark.parametrize('size,K,n', [(6, 5, 9), (9, 3, 130)])
def test_batched_unroll_matches_torch(size, K, n):
    """MuZeroNative.unroll: h, then K times g + f along given action sequences for a whole batch, against the fp32
    PyTorch module unrolled the same way (errors accumulate over the steps: 1e-3 per step on probabilities / values), and against the step-by-step device calls (bit-identical)."""
    from rlzero_b200.muzero import MuZeroNative
    net = _net(size, 1, 1)
    f, boards = _positions(n, size, min(5, size), 5)
    nat = MuZeroNative(net, n, K + 1, n_in_row=min(5, size))
    A = size * size
    rs = np.random.RandomState(K)
    actions = rs.randint(0, A, size=(n, K)).astype(np.int32)
    logp, value = nat.unroll(f.root_rows, f.root_meta, actions)
    assert logp.shape == (K + 1, n, nat.AS) and value.shape == (K + 1, n)
    obs = torch.from_numpy(np.stack([b.current_state() for b in boards])).float().cuda()
    with torch.no_grad():
        s, lt, vt = net.initial_inference(obs)
        outs = [(lt, vt)]
        for k in range(K):
            s, lt, vt = net.recurrent_inference(s, torch.from_numpy(actions[:, k]).cuda())
            outs.append((lt, vt))
    for k, (lt, vt) in enumerate(outs):
        tol = 1e-3 * (k + 1)          # the bf16 error of every further dynamics step adds up
        assert (logp[k][:, :A].exp() - lt.exp()).abs().max().item() < tol, k
        assert (value[k] - vt.reshape(-1)).abs().max().item() < tol, k
    # the same through the single-step calls
    nat2 = MuZeroNative(net, n, K + 1, n_in_row=min(5, size))
    nat2.representation(f.root_rows, f.root_meta, 0)
    l0, v0 = (x.clone() for x in nat2.prediction(0))
    assert torch.equal(l0, logp[0]) and torch.equal(v0, value[0])
    for k in range(K):
        nat2.dynamics(torch.full((n,), k, dtype=torch.int32, device='cuda'), torch.from_numpy(actions[:, k]).cuda(), k + 1)
        lk, vk = nat2.prediction(k + 1)
        assert torch.equal(lk, logp[k + 1]) and torch.equal(vk, value[k + 1])
    with pytest.raises(ValueError):
        nat.unroll(f.root_rows, f.root_meta, np.zeros((n, K + 3), dtype=np.int32))
