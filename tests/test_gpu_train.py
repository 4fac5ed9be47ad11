"""GPU tests of the training-side data path: device get_equi_data against the live
reference's fixture, the device replay buffer against deque + random.sample, and a short run of
the whole TrainPipeline in both collection modes."""
import ctypes as C
import json
import os
import random
from collections import deque

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('size,k', [(6, 4), (3, 3), (15, 5)])
def test_device_augmentation_matches_host_get_equi_data(size, k):
    """rz_augment_equi on trajectory records == get_equi_data on (current_state, pi, z) of the same
    plies, sample by sample, in the reference's order."""
    from rlzero_b200 import _lib as L
    from rlzero_b200.games.gomoku import GomokuEnv
    from rlzero_b200.train_pipeline import TrainPipeline, augment_equi_device
    rs = np.random.RandomState(size)
    A = size * size
    AS = (A + 31) // 32 * 32
    recs, rows, infos, pis = [], [], [], []
    for g in range(12):
        env = GomokuEnv(size, k)
        env.reset()
        for m in rs.permutation(A)[:rs.randint(0, A - 1)]:
            env.step(int(m))
            if env.game_end_winner()[0]:
                break
        if env.game_end_winner()[0]:
            continue
        pi = rs.rand(A).astype(np.float32)
        z = float(rs.choice([-1, 0, 1]))
        recs.append((env.current_state(), pi.astype(np.float64), z))
        r, m = env.device_state()
        rows.append(r)
        infos.append([env.current_player(), env.last_move, int(z), 0, 0, 0])
        pis.append(np.concatenate([pi, np.zeros(AS - A, dtype=np.float32)]))
    tp = TrainPipeline.__new__(TrainPipeline)
    tp.board_size = size
    want = tp.get_equi_data(recs)
    gd = L.GameDesc(size, k, A, AS)
    s, p, z = augment_equi_device(gd, torch.cat(rows).cuda().contiguous(),
                                  torch.tensor(infos, dtype=torch.int32, device='cuda'),
                                  torch.tensor(np.stack(pis), device='cuda'))
    s, p, z = s.cpu().numpy(), p.cpu().numpy(), z.cpu().numpy()
    assert len(want) == len(z)
    for j, (ws, wp, wz) in enumerate(want):
        assert np.array_equal(s[j], ws.astype(np.float32)), j
        assert np.array_equal(p[j], wp.astype(np.float32)), j
        assert z[j] == wz


def test_device_replay_buffer_is_a_deque_with_random_sample():
    from rlzero_b200.train_pipeline import DeviceReplayBuffer
    H, cap = 3, 50
    buf = DeviceReplayBuffer(cap, H)
    dq = deque(maxlen=cap)
    rs = np.random.RandomState(0)
    serial = 0
    for n in (7, 20, 30, 64, 5, 120, 3):
        st = torch.zeros(n, 4, H, H)
        st[:, 0, 0, 0] = torch.arange(serial, serial + n, dtype=torch.float32)
        pi = torch.tensor(rs.rand(n, 9), dtype=torch.float32)
        z = torch.arange(serial, serial + n, dtype=torch.float32)
        buf.extend(st.cuda(), pi.cuda(), z.cuda())
        for i in range(n):
            dq.append((st[i].numpy(), pi[i].numpy(), float(z[i])))
        serial += n
        assert len(buf) == len(dq)
        random.seed(serial)
        want = random.sample(dq, min(8, len(dq)))
        random.seed(serial)
        s, p, zz = buf.sample(min(8, len(dq)))
        for j, (ws, wp, wz) in enumerate(want):
            assert np.array_equal(s[j].cpu().numpy(), ws) and np.array_equal(p[j].cpu().numpy(), wp)
            assert float(zz[j]) == wz


def test_train_pipeline_runs_in_both_modes(tmp_path, monkeypatch):
    """A few iterations of the whole loop: reference-style single game collection, then batched
    collection with device augmentation + device replay buffer; losses finite, buffers fill, weights
    change, evaluation against RolloutPlayer returns a ratio."""
    from rlzero_b200.train_pipeline import TrainPipeline
    monkeypatch.chdir(tmp_path)
    np.random.seed(0)
    random.seed(0)
    torch.manual_seed(0)
    tp = TrainPipeline(board_size=5, n_in_row=4, n_playout=20, game_batch_num=3, check_freq=3,
                       pure_mcts_playout_num=20)
    tp.batch_size = 16
    before = [p.detach().clone() for p in tp.alphazero_agent.policy_value_net.parameters()]
    tp.run()
    assert len(tp.data_buffer) > tp.batch_size and tp.episode_len > 0
    after = list(tp.alphazero_agent.policy_value_net.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(after, before))
    assert os.path.exists(os.path.join('current_policy.model', 'model.th'))
    # batched mode
    tp2 = TrainPipeline(board_size=5, n_in_row=4, n_playout=16, n_parallel_games=32, game_batch_num=2,
                        check_freq=100)
    tp2.batch_size = 16
    tp2.play_batch_size = 8
    tp2.collect_selfplay_data(8)
    assert len(tp2.device_buffer) > tp2.batch_size
    loss, entropy = tp2.policy_update()
    assert np.isfinite(loss) and np.isfinite(entropy)
    # the search must now run with the updated weights (graph re-captured)
    tp2.collect_selfplay_data(4)
    ratio = tp2.policy_evaluate(n_games=2)
    assert 0.0 <= ratio <= 1.0


def test_learn_and_policy_value_accept_device_tensors():
    """AlphaZeroAgent.learn / policy_value_device on CUDA tensors (the batched pipeline's mini-batches, which never
    visit the host) give exactly what the reference-style list / numpy inputs give."""
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    rs = np.random.RandomState(0)
    states = rs.randint(0, 2, size=(24, 4, 6, 6)).astype(np.float32)
    pis = rs.dirichlet(np.ones(36), size=24).astype(np.float32)
    zs = rs.choice([-1.0, 0.0, 1.0], size=24).astype(np.float32)
    torch.manual_seed(1)
    a = AlphaZeroAgent(6)
    torch.manual_seed(1)
    b = AlphaZeroAgent(6)
    la = a.learn([s for s in states], [p for p in pis], [z for z in zs])
    lb = b.learn(torch.from_numpy(states).cuda(), torch.from_numpy(pis).cuda(), torch.from_numpy(zs).cuda())
    assert la == lb                  # same forward pass, same loss and entropy
    for pa, pb in zip(a.policy_value_net.parameters(), b.policy_value_net.parameters()):
        assert torch.allclose(pa, pb, atol=1e-6)      # cuDNN's weight gradients are not run-to-run deterministic
    pa, va = a.policy_value(states)
    pb, vb = a.policy_value_device(torch.from_numpy(states).cuda())
    # same log-probabilities; exp is numpy's on one side and CUDA's on the other (last-ulp differences)
    assert np.allclose(pa, pb.cpu().numpy(), rtol=1e-6, atol=0) and np.array_equal(va, vb.cpu().numpy())


def test_train_pipeline_single_game_loop_with_leaf_parallel_search(tmp_path, monkeypatch):
    """TrainPipeline(leaves_per_wave=K): the reference-style one-game-at-a-time loop with the leaf-parallel search
    (GameControl.start_self_play -> AlphaZeroPlayer.get_action): episodes finish, 8-fold augmented records land in
    the buffer with pi vectors that sum to 1, and an update step runs."""
    from rlzero_b200.train_pipeline import TrainPipeline
    monkeypatch.chdir(tmp_path)
    np.random.seed(1)
    random.seed(1)
    torch.manual_seed(1)
    tp = TrainPipeline(board_size=5, n_in_row=4, n_playout=24, game_batch_num=1, leaves_per_wave=4)
    assert tp.mcts_player.mcts.leaves_per_wave == 4
    tp.batch_size = 16
    tp.collect_selfplay_data(2)
    assert tp.episode_len > 0 and len(tp.data_buffer) >= 8 * tp.episode_len
    for state, pi, z in list(tp.data_buffer)[:16]:
        assert state.shape == (4, 5, 5) and abs(float(np.sum(pi)) - 1.0) < 1e-6 and z in (-1.0, 0.0, 1.0)
    loss, entropy = tp.policy_update()
    assert np.isfinite(loss) and np.isfinite(entropy)
