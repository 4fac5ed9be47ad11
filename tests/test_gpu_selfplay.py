"""GPU tests of the batched self-play driver: shard invariance (results do not depend on how
games are split over ranks), the host-buffer API, and the drain format."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _play(n_games, offset, ids, n_moves, net):
    from rlzero_b200.selfplay import BatchedSelfPlay
    sp = BatchedSelfPlay(n_games, 6, 4, net=net, n_playout=24, c_puct=5.0, temperature=1.0, add_noise=True,
                         global_offset=offset, seed=77)
    sp.set_random_start_positions(global_ids=ids, max_random_moves=7)
    moves = []
    for _ in range(n_moves):
        sp.play(1)
        torch.cuda.synchronize()
        moves.append(sp.forest.move.cpu().numpy().copy())
    sp.forest.raise_faults()
    rows, meta = sp.forest.boards()
    return np.stack(moves), rows, meta, sp


def test_results_do_not_depend_on_sharding():
    """8 games on one 'rank' == the same 8 global ids split 4 + 4 over two 'ranks': same sampled
    moves (Dirichlet noise and move sampling are keyed by the global game id), same positions."""
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    torch.manual_seed(3)
    net = ResNetPolicyValueNet(6, n_blocks=1).cuda().eval()
    ids = np.arange(8)
    m_all, rows_all, meta_all, _ = _play(8, 0, ids, 3, net)
    m_lo, rows_lo, meta_lo, _ = _play(4, 0, ids[:4], 3, net)
    m_hi, rows_hi, meta_hi, _ = _play(4, 4, ids[4:], 3, net)
    assert np.array_equal(m_all[:, :4], m_lo) and np.array_equal(m_all[:, 4:], m_hi)
    assert np.array_equal(rows_all[:4], rows_lo) and np.array_equal(rows_all[4:], rows_hi)
    assert (m_all >= 0).all()


def test_get_actions_host_api_and_drain():
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    from rlzero_b200.selfplay import BatchedSelfPlay
    torch.manual_seed(5)
    net = ResNetPolicyValueNet(6, n_blocks=1).cuda().eval()
    G = 16
    sp = BatchedSelfPlay(G, 6, 4, net=net, n_playout=30, temperature=1.0, add_noise=False, seed=9)
    sp.set_random_start_positions(max_random_moves=5)
    rows, meta = sp.forest.boards()
    moves, pi, visits = sp.get_actions(rows, meta)
    assert moves.shape == (G,) and pi.shape == (G, 36) and visits.shape == (G, 36)
    assert (visits.sum(1) == 29).all()                 # the root expansion consumes playout #1
    np.testing.assert_allclose(pi.sum(1), 1.0, atol=1e-5)
    occ = ((rows[:, 0] | rows[:, 1])[:, :, None] >> np.arange(6)[None, None, :]) & 1
    assert (visits[occ.reshape(G, 36) == 1] == 0).all()  # no visits on occupied squares
    assert all(occ.reshape(G, 36)[g, moves[g]] == 0 for g in range(G))
    # a second call from the same positions repeats exactly (fresh trees, same seeds)
    moves2, pi2, visits2 = sp.get_actions(rows, meta)
    assert np.array_equal(visits, visits2) and np.array_equal(moves, moves2)
    # play whole games and drain them in the reference's tuple format (game.py:113-134)
    sp2 = BatchedSelfPlay(8, 6, 4, net=net, n_playout=12, temperature=1.0, add_noise=True, seed=11)
    sp2.play(36)
    states, pis, zs, info = sp2.drain()
    st = sp2.stats()
    assert st['games_done'] >= 8 and len(zs) == len(states) == len(pis) > 0
    assert states.shape[1:] == (4, 6, 6) and set(np.unique(zs)).issubset({-1.0, 0.0, 1.0})
    # plane 3 is all ones iff an even number of stones is on the board (gomoku_env.py:110-111)
    stones = states[:, 0].sum((1, 2)) + states[:, 1].sum((1, 2))
    assert np.array_equal(states[:, 3, 0, 0] == 1.0, stones % 2 == 0)
