"""CPU tests: the numpy restatement of the reference's training step (oracle/train_oracle.py: forward, hand-derived
backward, Adam) against PyTorch autograd / torch.optim.Adam in float64 -- and against the LIVE reference's
AlphaZeroAgent.learn when /root/reference is present (authoring container)."""
import numpy as np
import pytest
import torch

from oracle import ref_loader, train_oracle


def _module(size, seed):
    from rlzero_b200.games.gomoku.policy_value_net import PolicyValueNet
    torch.manual_seed(seed)
    return PolicyValueNet(size).double()


def _batch(size, n, seed):
    rs = np.random.RandomState(seed)
    x = (rs.rand(n, 4, size, size) < 0.35).astype(np.float64)
    pi = rs.dirichlet(np.ones(size * size), size=n)
    z = rs.choice([-1.0, 0.0, 1.0], size=n)
    return x, pi, z


@pytest.mark.parametrize('size,n', [(3, 5), (6, 8), (9, 3)])
def test_loss_and_gradients_match_autograd(size, n):
    net = _module(size, size)
    x, pi, z = _batch(size, n, size + 1)
    p = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    loss, entropy, g = train_oracle.loss_and_grads(p, x, pi, z)
    log_p, v = net(torch.from_numpy(x))
    tl = torch.nn.functional.mse_loss(v.view(-1), torch.from_numpy(z)) - torch.mean(torch.sum(torch.from_numpy(pi) * log_p, dim=1))
    tl.backward()
    te = -torch.mean(torch.sum(torch.exp(log_p) * log_p, dim=1))
    assert abs(loss - tl.item()) < 1e-12 and abs(entropy - te.item()) < 1e-12
    for k, prm in net.named_parameters():
        assert g[k].shape == tuple(prm.shape), k
        assert np.abs(g[k] - prm.grad.numpy()).max() < 1e-12, k


def test_three_adam_steps_match_torch_optim():
    size, n = 6, 8
    net = _module(size, 1)
    opt = torch.optim.Adam(net.parameters(), lr=2e-3, weight_decay=1e-4)
    p = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    state = {}
    for step in range(3):
        x, pi, z = _batch(size, n, 10 + step)
        loss, _, g = train_oracle.loss_and_grads(p, x, pi, z)
        p = train_oracle.adam_step(p, g, state, lr=2e-3, weight_decay=1e-4)
        log_p, v = net(torch.from_numpy(x))
        tl = torch.nn.functional.mse_loss(v.view(-1), torch.from_numpy(z)) - torch.mean(torch.sum(torch.from_numpy(pi) * log_p, dim=1))
        opt.zero_grad()
        tl.backward()
        opt.step()
        assert abs(loss - tl.item()) < 1e-11
    for k, prm in net.named_parameters():
        assert np.abs(p[k] - prm.detach().numpy()).max() < 1e-11, k


@pytest.mark.skipif(not ref_loader.available(), reason='needs the live reference under /root/reference')
def test_one_learn_step_matches_the_live_reference_agent():
    """AlphaZeroAgent.learn of the UNMODIFIED reference (float32, its own Adam) against the oracle in float64 from the
    same initial weights: loss / entropy within float32 accuracy, updated weights within 1e-5."""
    ref = ref_loader.load()
    size, n = 6, 16
    torch.manual_seed(4)
    agent = ref.AlphaZeroAgent(size, device='cpu')
    p = {k: v.detach().numpy().astype(np.float64) for k, v in agent.policy_value_net.state_dict().items()}
    x, pi, z = _batch(size, n, 5)
    loss, entropy, g = train_oracle.loss_and_grads(p, x, pi, z)
    p = train_oracle.adam_step(p, g, {}, lr=agent.optimizer.param_groups[0]['lr'],
                               weight_decay=agent.optimizer.param_groups[0]['weight_decay'])
    rl, re_ = agent.learn(list(x.astype(np.float32)), list(pi.astype(np.float32)), list(z.astype(np.float32)))
    assert abs(rl - loss) < 1e-5 and abs(re_ - entropy) < 1e-5
    for k, v in agent.policy_value_net.state_dict().items():
        assert np.abs(v.numpy() - p[k]).max() < 1e-5, k
