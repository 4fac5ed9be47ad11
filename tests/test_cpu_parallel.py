"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: sharding arithmetic, ragged
trajectory all-gather, weight broadcast.  The search path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from rlzero_b200 import parallel
    from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet
    try:
        # shard arithmetic: contiguous, disjoint, covering
        lo, hi = parallel.shard_range(11)
        spans = [parallel.shard_range(11, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 11
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert (lo, hi) == spans[rank]
        # ragged gather: rank r contributes 3 + 2 r plies (rank 1 more than rank 0), one rank may be empty
        for counts in ([3, 5], [0, 4]):
            n = counts[rank]
            rs = np.random.RandomState(100 + rank)
            states = rs.rand(n, 4, 6, 6).astype(np.float32)
            pis = rs.rand(n, 36).astype(np.float32)
            zs = rs.choice([-1.0, 0.0, 1.0], size=n).astype(np.float32)
            info = rs.randint(0, 50, size=(n, 6)).astype(np.int32)
            g_states, g_pis, g_zs, g_info, got_counts = parallel.gather_trajectories(states, pis, zs, info)
            assert got_counts == counts
            exp = []
            for r in range(world):
                rr = np.random.RandomState(100 + r)
                m = counts[r]
                exp.append((rr.rand(m, 4, 6, 6).astype(np.float32), rr.rand(m, 36).astype(np.float32),
                            rr.choice([-1.0, 0.0, 1.0], size=m).astype(np.float32),
                            rr.randint(0, 50, size=(m, 6)).astype(np.int32)))
            assert np.array_equal(g_states, np.concatenate([e[0] for e in exp]))
            assert np.array_equal(g_pis, np.concatenate([e[1] for e in exp]))
            assert np.array_equal(g_zs, np.concatenate([e[2] for e in exp]))
            assert np.array_equal(g_info, np.concatenate([e[3] for e in exp]))
        # compact device-record gather (the NCCL path's function on gloo/CPU tensors): ragged, one rank empty,
        # local slot -> global game id, pi travels bit-cast and comes back bit for bit
        for counts in ([3, 5], [0, 4], [0, 0]):
            n = counts[rank]
            rs = np.random.RandomState(200 + rank)
            rows = torch.from_numpy(rs.randint(0, 1 << 15, size=(n, 2, 15)).astype(np.int32))
            info = torch.from_numpy(rs.randint(0, 50, size=(n, 6)).astype(np.int32))
            pi = torch.from_numpy(rs.rand(n, 256).astype(np.float32))
            g_rows, g_info, g_pi, got = parallel.gather_records_device(rows, info, pi, global_offset=1000 * rank)
            assert got == counts
            if sum(counts) == 0:
                continue
            exp = []
            for r in range(world):
                rr = np.random.RandomState(200 + r)
                m = counts[r]
                e_rows = rr.randint(0, 1 << 15, size=(m, 2, 15)).astype(np.int32)
                e_info = rr.randint(0, 50, size=(m, 6)).astype(np.int32)
                e_info[:, 3] += 1000 * r
                exp.append((e_rows, e_info, rr.rand(m, 256).astype(np.float32)))
            assert np.array_equal(g_rows.numpy(), np.concatenate([e[0] for e in exp]))
            assert np.array_equal(g_info.numpy(), np.concatenate([e[1] for e in exp]))
            assert np.array_equal(g_pi.numpy(), np.concatenate([e[2] for e in exp]))
        p_rows, p_info, p_pi = parallel.unpack_records(parallel.pack_records(rows, info, pi), 15)
        assert torch.equal(p_rows, rows) and torch.equal(p_info, info) and torch.equal(p_pi, pi)
        # weight broadcast: every rank ends with rank 0's parameters, bit for bit
        torch.manual_seed(1234 + rank)
        net = PolicyValueNet(6)
        torch.manual_seed(1234)
        ref = PolicyValueNet(6)
        moved = parallel.broadcast_weights(net, src=0)
        assert moved == sum(p.numel() * p.element_size() for p in ref.parameters())
        for (k, a), (_, b) in zip(net.state_dict().items(), ref.state_dict().items()):
            assert torch.equal(a, b), k
        open(os.path.join(out_dir, 'ok%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), 'ok%d' % r)) for r in range(world))
