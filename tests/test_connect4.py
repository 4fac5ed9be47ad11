"""Connect Four (BASELINE.json config 2; the reference has no such env): the oracle's definition,
the reference's search running on it (fixtures from the LIVE reference MCTS, which is
game-agnostic), and the device kernels (RZ_GAME_CONNECT4) against both."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = json.load(open(os.path.join(HERE, 'golden', 'connect4.json')))


def _board(pre):
    from oracle import pyoracle
    b = pyoracle.ConnectFourBoard()
    b.reset()
    for m in pre:
        b.step(m)
    return b


@pytest.mark.parametrize('case', FIX['cases'], ids=lambda c: 'c4_%d_%d' % (c['n_playout'], len(c['pre'])))
def test_oracle_search_on_connect4_matches_reference(case):
    from oracle import pyoracle
    from oracle.evaluators import make_policy_value_fn
    b = _board(case['pre'])
    s = pyoracle.Search(make_policy_value_fn(case['eval_id']), case['n_playout'], case['c_puct'])
    stages = iter(case['stages'])
    acts, _ = s.simulate(b, 1.0)
    st = next(stages)
    assert list(acts) == st['acts']
    assert s.root_visits(7).tolist() == st['visits'] and s.root_values(7).tolist() == st['W']
    assert (s.root.n, s.root.w) == (st['root_N'], st['root_W'])
    for move, n in case['chain']:
        b.step(move)
        s.update_with_move(move)
        s.n_playout = n
        if b.game_end_winner()[0]:
            break
        s.simulate(b, 1.0)
        st = next(stages)
        assert s.root_visits(7).tolist() == st['visits'] and s.root_values(7).tolist() == st['W']


def test_oracle_connect4_rules_edge_cases():
    from oracle import pyoracle
    b = pyoracle.ConnectFourBoard()
    b.reset()
    for _ in range(3):          # vertical four for player 0 in column 0
        b.step(0)
        b.step(1)
    _, reward, win, _ = b.step(0)
    assert win and reward == 1 and b.game_end_winner() == (True, 0)
    b.reset()
    for a in (0, 1, 1, 2, 3, 2, 2, 3, 4, 3):   # rising diagonal (0,0) (1,1) (2,2) (3,3) for player 0
        b.step(a)
    assert b.game_end_winner() == (False, -1)
    _, reward, win, _ = b.step(3)
    assert win and b.game_end_winner() == (True, 0)
    b.reset()
    for _ in range(6):
        b.step(6)
    assert 6 not in b.leagel_actions() and len(b.leagel_actions()) == 6
    with pytest.raises(AssertionError):
        b.step(6)


@pytest.mark.gpu
def test_device_connect4_rules_match_fixture_games():
    """ConnectFourEnv (step / legal / winner / planes computed by the CUDA kernels) replays the
    fixture games: rewards, terminal flags, legal lists and observation planes all equal."""
    from rlzero_b200.games.connect4 import ConnectFourEnv
    for plies in FIX['games']:
        env = ConnectFourEnv()
        env.reset()
        for ply in plies:
            obs, reward, win, _ = env.step(ply['a'])
            assert (reward, win) == (ply['reward'], ply['win'])
            assert env.game_end_winner() == (ply['end'], ply['winner'])
            assert list(env.leagel_actions()) == ply['legal'] and env.last_move == ply['last']
            assert obs.astype(np.int8).reshape(-1).tolist() == ply['planes']
    env = ConnectFourEnv()
    env.reset()
    for _ in range(6):
        env.step(2)
    with pytest.raises(AssertionError):
        env.step(2)


@pytest.mark.gpu
@pytest.mark.parametrize('case', FIX['cases'], ids=lambda c: 'c4_%d_%d' % (c['n_playout'], len(c['pre'])))
def test_device_search_on_connect4_matches_reference(case):
    """AlphaZeroMCTS on the device over ConnectFourEnv == the live reference's AlphaZeroMCTS over the
    oracle board: visits and fp64 value sums bit for bit, including tree reuse."""
    from oracle.evaluators import make_policy_value_fn
    from rlzero_b200.games.connect4 import ConnectFourEnv
    from rlzero_b200.mcts import AlphaZeroMCTS
    env = ConnectFourEnv()
    env.reset()
    for m in case['pre']:
        env.step(m)
    s = AlphaZeroMCTS(make_policy_value_fn(case['eval_id']), n_playout=case['n_playout'], c_puct=case['c_puct'])
    stages = iter(case['stages'])

    def check(st):
        visits, w, has, root_n, root_w = s._forest.root_stats()
        assert visits[0].tolist() == st['visits'] and w[0].tolist() == st['W']
        assert (int(root_n[0]), float(root_w[0])) == (st['root_N'], st['root_W'])
    acts, probs = s.simulate(env, 1.0)
    st = next(stages)
    assert list(acts) == st['acts']
    check(st)
    for move, n in case['chain']:
        env.step(move)
        s.update_with_move(move)
        s.n_playout = n
        if env.game_end_winner()[0]:
            break
        s.simulate(env, 1.0)
        check(next(stages))


@pytest.mark.gpu
def test_connect4_batched_selfplay_with_resnet():
    """Config 2 in miniature: batched self-play on 6x7 with a ResNet on the tensor-core path; the
    network matches PyTorch, trees match the oracle fed by the same network, games finish."""
    import torch
    from oracle import pyoracle
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    from rlzero_b200.games.gomoku.policy_value_net import NativeForward, ResNetPolicyValueNet
    from rlzero_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(0)
    net = ResNetPolicyValueNet(6, n_blocks=2, board_width=7, n_actions=7).cuda().eval()
    nf = NativeForward(net, max_batch=8)
    assert nf.mode == 'tc' and (nf.H, nf.W, nf.A) == (6, 7, 7)
    x = np.random.RandomState(0).randint(0, 2, size=(8, 4, 6, 7)).astype(np.float32)
    logp, v = (t.cpu() for t in nf.forward_planes(x))
    ref = ResNetPolicyValueNet(6, n_blocks=2, board_width=7, n_actions=7).eval()
    ref.load_state_dict({k: t.cpu() for k, t in net.state_dict().items()})
    with torch.no_grad():
        lt, vt = ref(torch.from_numpy(x))
    assert (logp[:, :7].exp() - lt.exp()).abs().max().item() < 1e-3
    assert (v - vt.reshape(-1)).abs().max().item() < 1e-3
    # search parity with the real net, noise off
    G, P = 64, 50
    sp = BatchedSelfPlay(G, 6, 4, net=net, n_playout=P, add_noise=False, seed=3, board_width=7,
                         game_type=L.GAME_CONNECT4)
    sp.set_random_start_positions(max_random_moves=9)
    rows0, meta0 = sp.forest.boards()
    sp.warm_up()
    for _ in range(P - 2):
        sp.step_wave()
    torch.cuda.synchronize()
    sp.forest.raise_faults()
    visits, w, has, root_n, root_w = sp.forest.root_stats()
    agent = AlphaZeroAgent(6, net=net)
    for g in (0, 5, 63):
        b = pyoracle.ConnectFourBoard()
        b.reset()
        rs = np.random.RandomState(1000 + g)
        seq = rs.permutation(np.repeat(np.arange(7), 6))[:(1000 + g) % 9]
        for m in seq:
            b.step(int(m))
        assert len(b.states) == int(meta0[g, L.META_STONES])
        s = pyoracle.Search(agent.policy_value_fn, P - 1, 5)
        s.simulate(b, 1.0)
        assert visits[g].tolist() == s.root_visits(7).tolist(), g
    # whole games: finish, drain, z in {-1, 0, 1}
    sp2 = BatchedSelfPlay(16, 6, 4, net=net, n_playout=12, add_noise=True, seed=5, board_width=7,
                          game_type=L.GAME_CONNECT4)
    sp2.play(42)
    assert sp2.stats()['games_done'] >= 16
    sp2.forest.raise_faults()
