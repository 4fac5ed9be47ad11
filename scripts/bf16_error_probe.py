#!/usr/bin/env python
"""How far is the bf16 tensor-core forward from fp32 once the network is TRAINED?  (VERDICT r1, weak point 1:
the 1e-3 bound was only shown at random init.)

Trains ResNet-10/128 on 15x15 with the native tensor-core step on self-play records that BatchedSelfPlay produces
with the network itself (search targets pi, outcomes z, 8-fold augmented on the device), and after every few
generations compares the inference path (NativeForward, bf16 activations between the 21 layers) with the PyTorch
fp32 forward of the same weights (CPU, no TF32) on held-out positions: max |dp|, max |dv|, logit range.

    python scripts/bf16_error_probe.py [generations] [steps_per_generation]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent  # noqa: E402
from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet  # noqa: E402
from rlzero_b200.selfplay import BatchedSelfPlay  # noqa: E402
from rlzero_b200.train_pipeline import augment_equi_device  # noqa: E402

H, BLOCKS = int(os.environ.get('RZ_PROBE_BOARD', 15)), int(os.environ.get('RZ_PROBE_BLOCKS', 10))
G, PLAYOUTS, BATCH = 2048, 48, 2048


def compare(agent, x):
    ref = ResNetPolicyValueNet(H, n_blocks=BLOCKS).eval()
    ref.load_state_dict({k: v.detach().cpu() for k, v in agent.policy_value_net.state_dict().items()})
    with torch.no_grad():
        lt, vt = ref(x.cpu())
    probs, vals = agent.policy_value(x.cpu().numpy())
    dp = float(np.abs(probs - lt.exp().numpy()).max())
    dv = float(np.abs(vals.reshape(-1) - vt.reshape(-1).numpy()).max())
    # the same network in PyTorch fp32 arithmetic with bf16 STORAGE where the kernels have it (folded weights and the
    # activation between layers): what any bf16-activation implementation of this net computes
    import torch.nn.functional as F
    from rlzero_b200.games.gomoku.policy_value_net import _fold_bn
    r = lambda t: t.to(torch.bfloat16).float()
    with torch.no_grad():
        layers = ref.trunk_layers()
        outs = []
        a = x.cpu()
        for conv, bn, skip, relu in layers:
            w, b = _fold_bn(conv, bn)
            y = F.conv2d(a, r(w.float()), b.float(), padding=1)
            if skip is not None:
                y = y + outs[skip]
            a = r(F.relu(y))
            outs.append(a)
        le, ve = ref.heads(a)
        # A/B by emulation: the residual stream kept in fp32 (only the convolutions' inputs are rounded to bf16) -- what
        # an fp32-skip variant of the kernels would compute
        outs, a32 = [], x.cpu()
        for conv, bn, skip, relu in layers:
            w, b = _fold_bn(conv, bn)
            y = F.conv2d(r(a32), r(w.float()), b.float(), padding=1)
            if skip is not None:
                y = y + outs[skip]
            a32 = F.relu(y)
            outs.append(a32)
        ls, vs = ref.heads(a32)
    emu_dp = float((le.exp() - lt.exp()).abs().max())
    emu_dv = float((ve.reshape(-1) - vt.reshape(-1)).abs().max())
    return (dp, dv, float((lt.max(dim=1).values - lt.min(dim=1).values).max()), float(lt.exp().max()), float(vt.abs().max()),
            emu_dp, emu_dv, float((ls.exp() - lt.exp()).abs().max()), float((vs.reshape(-1) - vt.reshape(-1)).abs().max()))


def run(board=15, blocks=10, gens=8, steps=40, games=2048, playouts=48, batch=2048, log=print):
    """Generator-free driver: returns the list of per-generation records (also passed to ``log`` as JSON lines)."""
    global H, BLOCKS, G, PLAYOUTS, BATCH
    H, BLOCKS, G, PLAYOUTS, BATCH = board, blocks, games, playouts, batch
    GENS, STEPS = gens, steps
    records = []
    torch.manual_seed(0)
    agent = AlphaZeroAgent(H, net=ResNetPolicyValueNet(H, n_blocks=BLOCKS), learning_rate=2e-3)
    sp = BatchedSelfPlay(G, H, 5, evaluator=agent.native, n_playout=PLAYOUTS, add_noise=True, seed=1, ring_capacity=1 << 18)
    sp.set_random_start_positions(max_random_moves=60)
    held = None
    step = 0
    for gen in range(GENS + 1):
        if gen > 0:
            # play until enough finished plies, then train on them
            states = None
            for _ in range(40):
                sp.play(2)
                rec = sp.forest.drain_trajectories_device()
                if rec['info'].shape[0]:
                    s_, p_, z_ = augment_equi_device(sp.forest.gdesc, rec['rows'], rec['info'], rec['pi'])
                    states = s_ if states is None else torch.cat([states, s_])
                    pis = p_ if states is s_ else torch.cat([pis, p_])
                    zs = z_ if states is s_ else torch.cat([zs, z_])
                if states is not None and states.shape[0] >= 4 * BATCH:
                    break
            if states is None:
                log(json.dumps({'generation': gen, 'note': 'no finished games yet'}))
                continue
            if held is None:
                held = states[:256].clone()
            gpu = torch.Generator(device='cpu').manual_seed(gen)
            losses = []
            for _ in range(STEPS):
                idx = torch.randperm(states.shape[0], generator=gpu)[:BATCH].to(states.device)
                losses.append(agent.learn(states[idx], pis[idx], zs[idx])[0])
                step += 1
        x = held if held is not None else torch.from_numpy((np.random.RandomState(0).rand(256, 4, H, H) < 0.15).astype(np.float32))
        dp, dv, rng, pmax, vmax, emu_dp, emu_dv, skip_dp, skip_dv = compare(agent, x)
        out = {'generation': gen, 'steps': step, 'max_dp': dp, 'max_dv': dv, 'max_logit_range': rng, 'max_p': pmax,
               'max_abs_v': vmax, 'torch_bf16_storage_emulation_max_dp': emu_dp, 'torch_bf16_storage_emulation_max_dv': emu_dv,
               'emulated_fp32_skip_stream_max_dp': skip_dp, 'emulated_fp32_skip_stream_max_dv': skip_dv}
        if gen > 0:
            out['loss_first'], out['loss_last'] = losses[0], losses[-1]
            out['records'] = int(states.shape[0])
        records.append(out)
        log(json.dumps(out))
    return records


if __name__ == '__main__':
    run(H, BLOCKS, int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 40,
        log=lambda s: print(s, flush=True))
