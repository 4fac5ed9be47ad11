"""GPU parity tests of the hand-written training step (rlzero_b200/learn.py, csrc/rz_learn.cu) against
oracle/train_oracle.py, the numpy float64 restatement of AlphaZeroAgent.learn
(rlzero/games/gomoku/alphazero_agent.py:59-86) that tests/test_train_oracle.py pins to autograd and to the live
reference agent.  Tolerance: the north star's float32 figure, 1e-5, written out below."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(size, B, seed):
    rs = np.random.RandomState(seed)
    x = np.zeros((B, 4, size, size), dtype=np.float32)
    for i in range(B):
        k = rs.randint(0, size * size // 2)
        sq = rs.permutation(size * size)[:k]
        for j, s in enumerate(sq):
            x[i, j % 2, s // size, s % size] = 1.0
        if k:
            x[i, 2, sq[-1] // size, sq[-1] % size] = 1.0
        if k % 2 == 0:
            x[i, 3] = 1.0
    pi = rs.dirichlet(0.3 * np.ones(size * size), size=B).astype(np.float32)
    z = rs.choice([-1.0, 0.0, 1.0], size=B).astype(np.float32)
    return x, pi, z


def _oracle_params(net):
    return {k: v.detach().cpu().double().numpy().copy() for k, v in net.state_dict().items()}


@pytest.mark.parametrize('path', ['tc', 'tc_wgrad', 'f32'])
@pytest.mark.parametrize('size,B', [(6, 32), (3, 8), (15, 48), (9, 130)])
def test_forward_loss_and_every_gradient_match_the_oracle(size, B, path):
    """path 'tc_wgrad' (the default up to 15x15): forward and data-gradient convolutions on the float32 CUDA-core
    kernel, the weight gradients on the tensor cores (bf16 high/low pairs, four partial products in fp32 accumulators);
    'f32': everything on CUDA cores.  Both within 1e-5 of the oracle (relative to each tensor's largest gradient).
    'tc' (AlphaZeroAgent(trainer='native_tc')): the whole trunk on the tensor cores (16-mantissa-bit pairs).  Its forward
    outputs are still within 1e-5, but a forward pass that differs by 2e-6 flips the ReLU mask of the pre-activations
    that close to zero, and a flipped mask is a 100 % error on that gradient element: measured 1e-3 ... 7e-3 of each
    trunk tensor's largest gradient (heads: 4e-6).  That is the class of ANY lower-precision forward -- PyTorch's default
    TF32 convolutions, three decimal digits, sit well above it -- so the test holds 'tc' to 2e-2."""
    from oracle import train_oracle
    from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet
    from rlzero_b200.learn import NativeTrainer
    torch.manual_seed(size)
    net = PolicyValueNet(size).cuda()
    tr = NativeTrainer(net)
    tr.use_tc_trunk, tr.use_tc_wgrad = path == 'tc', path != 'f32'
    tr._repack()
    x, pi, z = _batch(size, B, 1)
    p = _oracle_params(net)
    logp_o, v_o, _ = train_oracle.forward(p, x.astype(np.float64))
    loss_o, ent_o, g_o = train_oracle.loss_and_grads(p, x.astype(np.float64), pi.astype(np.float64), z.astype(np.float64))
    logp, v = tr.forward(torch.from_numpy(x).cuda())
    A = size * size
    np.testing.assert_allclose(logp[:, :A].cpu().numpy(), logp_o, atol=1e-5, rtol=0)
    np.testing.assert_allclose(v.cpu().numpy(), v_o.reshape(-1), atol=1e-5, rtol=0)
    vl, pl, ent = tr.backward(torch.from_numpy(pi).cuda(), torch.from_numpy(z).cuda()).tolist()
    assert abs((vl + pl) - loss_o) < 1e-5 and abs(ent - ent_o) < 1e-5
    worst = 0.0
    for name, g in tr.grads().items():
        ref = g_o[name]
        err = np.abs(g.cpu().numpy().astype(np.float64) - ref).max()
        scale = max(np.abs(ref).max(), 1e-3)
        worst = max(worst, err / scale)
        assert err <= (2e-2 if path == 'tc' else 1e-5) * scale + 1e-7, (name, err, scale)
    print('%s %dx%d B=%d: worst gradient error / largest gradient = %.1e' % (path, size, size, B, worst))


@pytest.mark.parametrize('size,B,steps', [(6, 32, 3), (15, 64, 2)])
def test_adam_steps_match_the_oracle(size, B, steps):
    """Whole learn() steps (forward, backward, Adam with weight decay): every parameter within 1e-5 of the float64
    oracle after each step, loss and entropy within 1e-5.  (Adam normalises by |g|: the update of a component is
    lr * g / (|g| + 1e-8), so the float32 rounding of a gradient of size 1e-5 -- about 1e-7 absolute here -- moves
    that parameter by 1e-2 * lr.  Components whose gradient was below 2e-5 in any step so far are therefore held to
    two learning rates instead of 1e-5; everything else to 1e-5.)"""
    from oracle import train_oracle
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    torch.manual_seed(7)
    agent = AlphaZeroAgent(size, learning_rate=1e-3)
    assert agent.trainer is not None
    p = _oracle_params(agent.policy_value_net)
    state = {}
    tiny = {}
    for step in range(steps):
        x, pi, z = _batch(size, B, 10 + step)
        loss_o, ent_o, g_o = train_oracle.loss_and_grads(p, x.astype(np.float64), pi.astype(np.float64), z.astype(np.float64))
        p = train_oracle.adam_step(p, g_o, state, lr=1e-3, weight_decay=1e-4)
        loss, ent = agent.learn(x, pi, z)
        assert abs(loss - loss_o) < 1e-5 and abs(ent - ent_o) < 1e-5
        for name, t in agent.policy_value_net.state_dict().items():
            err = np.abs(t.cpu().numpy().astype(np.float64) - p[name])
            tiny[name] = tiny.get(name, False) | (np.abs(g_o[name]) < 2e-5)
            assert err[~tiny[name]].max(initial=0.0) <= 1e-5, (step, name, err[~tiny[name]].max())
            assert err[tiny[name]].max(initial=0.0) <= 2.1e-3, (step, name)
    # the inference path saw every update (weights re-packed for the kernels)
    probs, vals = agent.policy_value(x)
    logp_o, v_o, _ = train_oracle.forward(p, x.astype(np.float64))
    np.testing.assert_allclose(probs, np.exp(logp_o), atol=1e-5, rtol=0)
    np.testing.assert_allclose(vals.reshape(-1), v_o.reshape(-1), atol=1e-5, rtol=0)


def test_native_step_equals_the_autograd_step():
    """The same batch through learn() (hand-written kernels) and learn_autograd() (PyTorch): loss, entropy, the
    updated parameters and the KL / explained-variance monitors of policy_update (tools/train_alphazero.py:99-137)."""
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x, pi, z = _batch(6, 32, 3)
    torch.manual_seed(5)
    a = AlphaZeroAgent(6, trainer='native')
    torch.manual_seed(5)
    b = AlphaZeroAgent(6, trainer='autograd')
    old_a, oldv_a = a.policy_value(x)
    old_b, oldv_b = b.policy_value(x)
    for _ in range(3):
        la, ea = a.learn(x, pi, z)
        lb, eb = b.learn(x, pi, z)
        assert abs(la - lb) < 1e-5 and abs(ea - eb) < 1e-5
    for (k, ta), (_, tb) in zip(a.policy_value_net.state_dict().items(), b.policy_value_net.state_dict().items()):
        d = (ta - tb).abs().flatten().float()
        # two float32 implementations: Adam turns rounding noise on near-zero gradient components into steps of up to
        # a learning rate (see test_adam_steps_match_the_oracle), everything else agrees to 1e-5
        assert torch.quantile(d, 0.99).item() < 1e-5 and d.max().item() < 3.1e-3, k
    new_a, newv_a = a.policy_value(x)
    new_b, newv_b = b.policy_value(x)
    kl = lambda o, n: np.mean(np.sum(o * (np.log(o + 1e-10) - np.log(n + 1e-10)), axis=1))
    ev = lambda v: 1 - np.var(z - v.flatten()) / np.var(z)
    assert abs(kl(old_a, new_a) - kl(old_b, new_b)) < 1e-5
    assert abs(ev(newv_a) - ev(newv_b)) < 1e-4


def test_save_restore_round_trip_with_the_native_trainer(tmp_path):
    """save_model / restore (alphazero_agent.py:99-125): same files and state_dict keys; the optimiser file is in
    torch.optim.Adam's format and restores the moments, so training continues identically."""
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    x, pi, z = _batch(6, 16, 4)
    torch.manual_seed(2)
    a = AlphaZeroAgent(6)
    a.learn(x, pi, z)
    a.save_model(str(tmp_path / 'ckpt'))
    sd = torch.load(str(tmp_path / 'ckpt' / 'optimizer.th'))
    assert set(sd) == {'state', 'param_groups'} and len(sd['state']) == 16
    assert sd['param_groups'][0]['weight_decay'] == 1e-4 and float(sd['state'][0]['step']) == 1.0
    ref = torch.optim.Adam(AlphaZeroAgent(6, trainer='autograd').policy_value_net.parameters())
    ref.load_state_dict(sd)                                    # a stock torch optimiser accepts the file
    torch.manual_seed(99)
    b = AlphaZeroAgent(6)
    b.restore(str(tmp_path / 'ckpt'))
    la, lb = a.learn(x, pi, z), b.learn(x, pi, z)
    assert la == lb
    for ta, tb in zip(a.policy_value_net.parameters(), b.policy_value_net.parameters()):
        assert torch.equal(ta, tb)


def test_the_step_is_deterministic():
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    x, pi, z = _batch(9, 64, 6)
    outs = []
    for _ in range(2):
        torch.manual_seed(3)
        a = AlphaZeroAgent(9)
        for _ in range(2):
            res = a.learn(x, pi, z)
        outs.append((res, [t.clone() for t in a.policy_value_net.parameters()]))
    assert outs[0][0] == outs[1][0]
    assert all(torch.equal(u, w) for u, w in zip(outs[0][1], outs[1][1]))
