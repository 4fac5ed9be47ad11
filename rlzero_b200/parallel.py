"""Multi-GPU plumbing around the self-play path (one process per GPU, ``torch.distributed``).

Self-play games are independent (the reference plays them one after another,
tools/train_alphazero.py:81-90), so they shard over ranks with NO collective on the search path:
rank r owns global game ids ``[r*G, (r+1)*G)`` and every per-game random stream is keyed by the
global id, which makes results independent of the shard count.

The only exchanges are off the hot path, once per generation / training step:

* ``gather_trajectories``  -- all ranks contribute the finished plies they drained
  (``BatchedSelfPlay.drain``) and every rank (or only the learner) receives the union, in
  global-game-id order, ready for ``TrainPipeline.get_equi_data`` (tools/train_alphazero.py:59-79);
* ``broadcast_weights``    -- after ``AlphaZeroAgent.learn`` on the learner rank the parameters and
  buffers of the policy-value module go to every rank, which then re-packs them for the kernels.

Backend: NCCL over NVLink on the GPUs; the same code runs on ``gloo`` (CPU tensors) in the tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_global_games, rank=None, world=None):
    """[lo, hi) global game ids of this rank: contiguous blocks, remainder to the low ranks."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    base, rem = divmod(int(n_global_games), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_for_backend():
    if dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def gather_trajectories(states, pis, zs, info=None, group=None):
    """All-gather ragged per-rank trajectory records.

    states float32 [n,4,H,W], pis float32 [n,A], zs float32 [n], info int32 [n,6] (optional; column
    3 is the LOCAL slot, column 4 the episode, column 5 the ply -- see include/rlzero_b200.h
    ``ring_info``).  Every rank returns the concatenation over ranks in rank order.
    """
    world = dist.get_world_size(group)
    dev = _device_for_backend()
    n = int(len(zs))
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[dist.get_rank(group)] = n
    dist.all_reduce(counts, group=group)
    counts = counts.cpu().tolist()
    n_max = max(counts) if counts else 0
    out = []
    for arr in (states, pis, zs, info):
        if arr is None:
            out.append(None)
            continue
        t = torch.as_tensor(np.ascontiguousarray(arr))
        pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:n] = t.to(dev)
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out.append(torch.cat([b[:c] for b, c in zip(bufs, counts)]).cpu().numpy())
    return tuple(out) + (counts,)


def broadcast_weights(module, src=0, group=None):
    """Broadcast parameters and buffers of ``module`` from rank ``src`` (flattened: one collective
    per dtype, sized for launch latency, not link count)."""
    dev = _device_for_backend()
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.reshape(-1).to(dev) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            k = t.numel()
            t.copy_(flat[off:off + k].reshape(t.shape).to(t.device))
            off += k
    return module
