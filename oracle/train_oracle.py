"""TEST INFRASTRUCTURE ONLY -- numpy (float64) restatement of the reference's training step.

Follows ``AlphaZeroAgent.learn`` (rlzero/games/gomoku/alphazero_agent.py:59-86) on the stock
``PolicyValueNet`` (rlzero/games/gomoku/policy_value_net.py:6-52):

    loss = mse(v, z) - mean_b sum_a pi[b, a] * log p[b, a]          (:70-75; the L2 term lives in Adam's weight_decay)
    entropy = -mean_b sum_a p * log p                               (:83-85, monitoring only)
    Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)       (:22-24, torch.optim.Adam semantics)

with the forward pass, the analytic backward pass and the optimiser update written out by hand -- no autograd -- so
that a hand-written backward pass on the device (a next-round item: today the product's ``learn`` is PyTorch autograd
on device tensors) has an oracle to be checked against.  Pinned by ``tests/test_train_oracle.py`` against PyTorch
autograd in float64 on the CPU.  Parameters use the module's own names and layouts (``state_dict``).
"""
import numpy as np


def _conv3x3(x, w, b):
    """x [B,Cin,H,W], w [Cout,Cin,3,3], b [Cout] -> [B,Cout,H,W], zero padding 1 (nn.Conv2d(..., 3, padding=1))."""
    B, C, H, W = x.shape
    xp = np.zeros((B, C, H + 2, W + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((B, w.shape[0], H, W), dtype=x.dtype)
    for kh in range(3):
        for kw in range(3):
            out += np.einsum('bchw,oc->bohw', xp[:, :, kh:kh + H, kw:kw + W], w[:, :, kh, kw])
    return out + b[None, :, None, None]


def _conv3x3_backward(x, w, dy):
    """Gradients of _conv3x3: (dx, dw, db)."""
    B, C, H, W = x.shape
    xp = np.zeros((B, C, H + 2, W + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    dxp = np.zeros_like(xp)
    dw = np.zeros_like(w)
    for kh in range(3):
        for kw in range(3):
            patch = xp[:, :, kh:kh + H, kw:kw + W]
            dw[:, :, kh, kw] = np.einsum('bohw,bchw->oc', dy, patch)
            dxp[:, :, kh:kh + H, kw:kw + W] += np.einsum('bohw,oc->bchw', dy, w[:, :, kh, kw])
    return dxp[:, :, 1:-1, 1:-1], dw, dy.sum(axis=(0, 2, 3))


def forward(p, x):
    """The stock network (policy_value_net.py:34-52).  Returns (log_p [B,HW], v [B,1], cache of activations)."""
    B = x.shape[0]
    a1 = np.maximum(_conv3x3(x, p['conv1.weight'], p['conv1.bias']), 0.0)
    a2 = np.maximum(_conv3x3(a1, p['conv2.weight'], p['conv2.bias']), 0.0)
    a3 = np.maximum(_conv3x3(a2, p['conv3.weight'], p['conv3.bias']), 0.0)
    pa = np.maximum(np.einsum('bchw,oc->bohw', a3, p['act_conv1.weight'][:, :, 0, 0])
                    + p['act_conv1.bias'][None, :, None, None], 0.0)
    pf = pa.reshape(B, -1)                                   # x.view(-1, 4*H*W): channel-major flatten
    logits = pf @ p['act_fc1.weight'].T + p['act_fc1.bias']
    m = logits.max(axis=1, keepdims=True)
    log_p = logits - m - np.log(np.exp(logits - m).sum(axis=1, keepdims=True))
    va = np.maximum(np.einsum('bchw,oc->bohw', a3, p['val_conv1.weight'][:, :, 0, 0])
                    + p['val_conv1.bias'][None, :, None, None], 0.0)
    vf = va.reshape(B, -1)
    h = np.maximum(vf @ p['val_fc1.weight'].T + p['val_fc1.bias'], 0.0)
    v = np.tanh(h @ p['val_fc2.weight'].T + p['val_fc2.bias'])
    return log_p, v, dict(x=x, a1=a1, a2=a2, a3=a3, pa=pa, pf=pf, va=va, vf=vf, h=h, v=v, log_p=log_p)


def loss_and_grads(p, x, pi, z):
    """(loss, entropy, grads) of alphazero_agent.py:66-85; grads keyed like the state_dict."""
    B = x.shape[0]
    log_p, v, c = forward(p, x)
    value_loss = np.mean((v.reshape(-1) - z) ** 2)
    policy_loss = -np.mean(np.sum(pi * log_p, axis=1))
    loss = value_loss + policy_loss
    prob = np.exp(log_p)
    entropy = -np.mean(np.sum(prob * log_p, axis=1))
    g = {}
    # value head
    dv = (2.0 * (v.reshape(-1) - z) / B).reshape(B, 1)
    dpre2 = dv * (1.0 - v ** 2)                              # tanh'
    g['val_fc2.weight'] = dpre2.T @ c['h']
    g['val_fc2.bias'] = dpre2.sum(axis=0)
    dh = (dpre2 @ p['val_fc2.weight']) * (c['h'] > 0)
    g['val_fc1.weight'] = dh.T @ c['vf']
    g['val_fc1.bias'] = dh.sum(axis=0)
    dva = (dh @ p['val_fc1.weight']).reshape(c['va'].shape) * (c['va'] > 0)
    g['val_conv1.weight'] = np.einsum('bohw,bchw->oc', dva, c['a3'])[:, :, None, None]
    g['val_conv1.bias'] = dva.sum(axis=(0, 2, 3))
    da3 = np.einsum('bohw,oc->bchw', dva, p['val_conv1.weight'][:, :, 0, 0])
    # policy head: d(-mean sum pi log_softmax) / dlogits = (softmax * sum_a pi - pi) / B
    dlogits = (prob * pi.sum(axis=1, keepdims=True) - pi) / B
    g['act_fc1.weight'] = dlogits.T @ c['pf']
    g['act_fc1.bias'] = dlogits.sum(axis=0)
    dpa = (dlogits @ p['act_fc1.weight']).reshape(c['pa'].shape) * (c['pa'] > 0)
    g['act_conv1.weight'] = np.einsum('bohw,bchw->oc', dpa, c['a3'])[:, :, None, None]
    g['act_conv1.bias'] = dpa.sum(axis=(0, 2, 3))
    da3 = da3 + np.einsum('bohw,oc->bchw', dpa, p['act_conv1.weight'][:, :, 0, 0])
    # trunk
    dz3 = da3 * (c['a3'] > 0)
    da2, g['conv3.weight'], g['conv3.bias'] = _conv3x3_backward(c['a2'], p['conv3.weight'], dz3)
    dz2 = da2 * (c['a2'] > 0)
    da1, g['conv2.weight'], g['conv2.bias'] = _conv3x3_backward(c['a1'], p['conv2.weight'], dz2)
    dz1 = da1 * (c['a1'] > 0)
    _, g['conv1.weight'], g['conv1.bias'] = _conv3x3_backward(c['x'], p['conv1.weight'], dz1)
    return float(loss), float(entropy), g


def adam_step(p, g, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
    """torch.optim.Adam (L2 weight decay folded into the gradient, bias-corrected moments):
    g += wd * p; m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
    ``state``: {'step': t, 'm': {...}, 'v': {...}} updated in place (created on first use)."""
    if 'step' not in state:
        state.update(step=0, m={k: np.zeros_like(v) for k, v in p.items()}, v={k: np.zeros_like(v) for k, v in p.items()})
    state['step'] += 1
    t = state['step']
    b1, b2 = betas
    for k in p:
        grad = g[k] + weight_decay * p[k]
        state['m'][k] = b1 * state['m'][k] + (1 - b1) * grad
        state['v'][k] = b2 * state['v'][k] + (1 - b2) * grad * grad
        denom = np.sqrt(state['v'][k]) / np.sqrt(1 - b2 ** t) + eps
        p[k] = p[k] - (lr / (1 - b1 ** t)) * state['m'][k] / denom
    return p
