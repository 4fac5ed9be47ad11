"""GPU tests of the tensor-core training kernels (csrc/rz_learn_tc.cu) against plain PyTorch fp32 on the same
bf16-rounded inputs: the tcgen05 weight gradient (MN-major operands), the data gradient through the forward kernel,
BatchNorm forward / backward in training mode, and the whole ResNet training step against autograd."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _tile(x_nchw):
    """float [n,C=128,H,W] -> bf16 padded tile layout [n*256,128] (zero pad squares)."""
    n, c, h, w = x_nchw.shape
    t = torch.zeros(n, 16, 16, c, device=x_nchw.device)
    t[:, :h, :w, :] = x_nchw.permute(0, 2, 3, 1)
    return t.reshape(n * 256, c).to(torch.bfloat16).contiguous()


def _untile(t, n, h, w):
    return t.float().reshape(n, 16, 16, -1)[:, :h, :w, :].permute(0, 3, 1, 2).contiguous()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


@pytest.mark.parametrize('n,h,w', [(1, 15, 15), (7, 15, 15), (64, 9, 9), (300, 6, 6), (130, 15, 15)])
def test_weight_gradient_on_the_tensor_cores(n, h, w):
    """rz_learn_conv_wgrad_tc == d/dW of F.conv2d(x, W, padding=1) contracted with dy (fp32 on the bf16 values)."""
    from rlzero_b200 import _lib as L
    lib = L.load()
    g = torch.Generator(device='cuda').manual_seed(n)
    x = torch.randn(n, 128, h, w, device='cuda', generator=g)
    dy = torch.randn(n, 128, h, w, device='cuda', generator=g)
    xt, dyt = _tile(x), _tile(dy)
    xr, dyr = _untile(xt, n, h, w), _untile(dyt, n, h, w)
    ref = torch.nn.grad.conv2d_weight(xr.double(), (128, 128, 3, 3), dyr.double(), padding=1)
    dw = torch.empty(128, 128, 3, 3, device='cuda')
    scratch = torch.empty(49 * 9 * 128 * 128, device='cuda')
    L.check(lib.rz_learn_conv_wgrad_tc(L.ptr(xt), L.ptr(dyt), L.ptr(dw), L.ptr(scratch), scratch.numel(), n, 0,
                                       L.stream_ptr()), 'rz_learn_conv_wgrad_tc')
    err = (dw.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-5 * scale, (err, scale)            # exact bf16 products, fp32 accumulation over n*225 positions
    # deterministic: a second launch gives the same bits
    dw2 = torch.empty_like(dw)
    L.check(lib.rz_learn_conv_wgrad_tc(L.ptr(xt), L.ptr(dyt), L.ptr(dw2), L.ptr(scratch), scratch.numel(), n, 0,
                                       L.stream_ptr()), 'rz_learn_conv_wgrad_tc')
    assert torch.equal(dw, dw2)


@pytest.mark.parametrize('n,h,w', [(5, 15, 15), (64, 9, 9)])
def test_data_gradient_is_the_forward_kernel_with_repacked_weights(n, h, w):
    from rlzero_b200 import _lib as L
    lib = L.load()
    g = torch.Generator(device='cuda').manual_seed(3)
    wgt = torch.randn(128, 128, 3, 3, device='cuda', generator=g) * 0.05
    dy = torch.randn(n, 128, h, w, device='cuda', generator=g)
    skip = torch.randn(n, 128, h, w, device='cuda', generator=g)
    wf = torch.empty(9 * 128 * 128, dtype=torch.bfloat16, device='cuda')
    wb = torch.empty_like(wf)
    L.check(lib.rz_learn_pack_conv_tc(L.ptr(wgt), L.ptr(wf), L.ptr(wb), L.stream_ptr()), 'rz_learn_pack_conv_tc')
    assert torch.equal(wf.view(9, 128, 128), wgt.permute(2, 3, 0, 1).reshape(9, 128, 128).to(torch.bfloat16))
    dyt, skt = _tile(dy), _tile(skip)
    out = torch.zeros_like(dyt)
    zero = torch.zeros(128, device='cuda')
    L.check(lib.rz_net_conv3x3_tc2(L.ptr(dyt), L.ptr(wb), L.ptr(zero), L.ptr(skt), L.ptr(out), n, h, w, 128, 0, 2, 2, 0,
                                   L.stream_ptr()), 'rz_net_conv3x3_tc2')
    wr = wgt.to(torch.bfloat16).double()
    ref = F.conv_transpose2d(_untile(dyt, n, h, w).double(), wr, padding=1) + _untile(skt, n, h, w).double()
    got = _untile(out, n, h, w).double()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()       # the output is rounded to bf16
    assert out.float().reshape(n, 16, 16, 128)[:, h:, :, :].abs().max().item() == 0.0


@pytest.mark.parametrize('n,h,w,with_skip', [(9, 15, 15, True), (64, 6, 6, False)])
def test_batchnorm_training_forward_and_backward(n, h, w, with_skip):
    from rlzero_b200 import _lib as L
    lib = L.load()
    g = torch.Generator(device='cuda').manual_seed(11)
    y = torch.randn(n, 128, h, w, device='cuda', generator=g) * 2 + 0.5
    skip = torch.randn(n, 128, h, w, device='cuda', generator=g)
    dout = torch.randn(n, 128, h, w, device='cuda', generator=g)
    yt, skt, dt = _tile(y), _tile(skip), _tile(dout)
    bn = torch.nn.BatchNorm2d(128).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-0.5, 0.5, generator=g)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    stats = torch.empty(4 * 128, device='cuda')
    scratch = torch.empty(148 * 4 * 256 + 256, device='cuda')      # (the kernels need 148 * 2 * 256 + 256)
    out = torch.empty_like(yt)
    L.check(lib.rz_learn_bn_forward(L.ptr(yt), L.ptr(skt) if with_skip else None, L.ptr(out), L.ptr(bn.weight.data),
                                    L.ptr(bn.bias.data), L.ptr(rm), L.ptr(rv), bn.eps, bn.momentum, L.ptr(stats),
                                    L.ptr(scratch), n, h, w, L.stream_ptr()), 'rz_learn_bn_forward')
    yr = _untile(yt, n, h, w).requires_grad_(True)
    sr = _untile(skt, n, h, w).requires_grad_(True)
    ref = bn(yr) + (sr if with_skip else 0.0)
    ref = F.relu(ref)
    got = _untile(out, n, h, w)
    assert (got - ref).abs().max().item() <= 2e-2                  # bf16 output
    assert torch.allclose(rm, bn.running_mean, atol=1e-5) and torch.allclose(rv, bn.running_var, rtol=1e-4, atol=1e-5)
    assert out.float().reshape(n, 16, 16, 128)[:, :, w:, :].abs().max().item() == 0.0
    # backward: the mask comes from the kernel's own (bf16) output, like autograd's from its own
    dgamma, dbeta = torch.empty(128, device='cuda'), torch.empty(128, device='cuda')
    dy, dz = torch.empty_like(yt), torch.empty_like(yt)
    L.check(lib.rz_learn_bn_backward(L.ptr(dt), L.ptr(out), L.ptr(yt), L.ptr(stats), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dy),
                                     L.ptr(dz), L.ptr(scratch), n, h, w, L.stream_ptr()), 'rz_learn_bn_backward')
    ref.backward(_untile(dt, n, h, w))
    assert (dgamma - bn.weight.grad).abs().max().item() <= 2e-3 * bn.weight.grad.abs().max().item() + 2e-2
    assert (dbeta - bn.bias.grad).abs().max().item() <= 2e-3 * bn.bias.grad.abs().max().item() + 2e-2
    assert (_untile(dy, n, h, w) - yr.grad).abs().max().item() <= 2e-2 * yr.grad.abs().max().item()
    if with_skip:
        assert (_untile(dz, n, h, w) - sr.grad).abs().max().item() <= 1e-2 * sr.grad.abs().max().item()


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


class _RoundBF16(torch.autograd.Function):
    """Round to bf16 in the forward pass AND round the gradient in the backward pass: the places where the kernels
    store a tensor (activations, raw convolution outputs and their gradients all live in HBM as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def _emulated_forward(net, x):
    """ResNetPolicyValueNet.forward in training mode with the kernels' storage points emulated in PyTorch fp32: conv
    weights and every stored tensor rounded to bf16, everything else (accumulation, BatchNorm statistics, heads) fp32.
    Autograd through this graph is the reference the hand-written backward pass is checked against: the ReLU masks are
    the kernels' masks, so what remains is rounding order."""
    r = _RoundBF16.apply
    wq = lambda conv: conv.weight + (conv.weight.to(torch.bfloat16).float() - conv.weight).detach()
    a = r(F.relu(F.conv2d(x, wq(net.stem), net.stem.bias, padding=1)))
    for blk in net.blocks:
        y1 = r(F.conv2d(a, wq(blk.conv1), blk.conv1.bias, padding=1))
        a1 = r(F.relu(blk.bn1(y1)))
        y2 = r(F.conv2d(a1, wq(blk.conv2), blk.conv2.bias, padding=1))
        a = r(F.relu(blk.bn2(y2) + a))
    return net.heads(a)


@pytest.mark.parametrize('size,blocks,B', [(6, 2, 32), (9, 1, 16), (6, 1, 8), (15, 3, 24), (15, 10, 64)])
def test_resnet_training_step_against_autograd(size, blocks, B):
    """ResNetTrainer (tensor-core trunk, bf16 activations, BatchNorm in training mode) against PyTorch autograd on the
    same weights and batch.

    (1) Against plain fp32 autograd: forward outputs, loss, entropy and running statistics (bf16 tolerances).  The
        gradients of the two agree only loosely, and that is the precision, not the backward pass: a pre-activation
        within the bf16 forward error of zero (~0.4 % of them) flips its ReLU mask, each flip changes that gradient
        element by 100 %, i.e. ~6 % of the tensor's norm per layer, accumulating over the depth (measured 0.06 at the
        last block to 0.30 at the stem of ResNet-10) -- PyTorch's own bf16 autocast differs from fp32 the same way.
    (2) Against autograd through a graph that rounds to bf16 exactly where the kernels store tensors.  Where that
        graph's forward pass reproduces the kernels' (same roundings, hence the same masks: the small cases) every
        parameter gradient agrees within 3 % of its norm -- the hand-written backward pass (tcgen05 weight gradient,
        data gradient, BatchNorm backward, skip, stem, heads) IS the derivative of the forward.  Where the two
        forwards differ in the last bf16 bit (fp32 summation order) some masks flip again; there the test asks for
        the same direction and size: cosine > 0.93, norms within 15 %."""
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    from rlzero_b200.learn import ResNetTrainer
    torch.manual_seed(size + blocks)
    net = ResNetPolicyValueNet(size, n_blocks=blocks).cuda()
    ref = ResNetPolicyValueNet(size, n_blocks=blocks).cuda()
    emu = ResNetPolicyValueNet(size, n_blocks=blocks).cuda()
    ref.load_state_dict(net.state_dict())
    emu.load_state_dict(net.state_dict())
    tr = ResNetTrainer(net)
    x, pi, z = _batch(size, B, 2)
    xt, pit, zt = (torch.from_numpy(t).cuda() for t in (x, pi, z))

    def loss_of(lp, v):
        return F.mse_loss(v.view(-1), zt) - torch.mean(torch.sum(pit * lp, dim=1))
    ref.train()
    emu.train()
    lp_ref, v_ref = ref(xt)
    loss_ref = loss_of(lp_ref, v_ref)
    ent_ref = -torch.mean(torch.sum(torch.exp(lp_ref) * lp_ref, dim=1))
    loss_ref.backward()
    lp_emu, v_emu = _emulated_forward(emu, xt)
    loss_of(lp_emu, v_emu).backward()
    logp, v = tr.forward(xt)
    A = size * size
    p_err = (logp[:, :A].exp() - lp_ref.exp()).abs().max().item()
    v_err = (v - v_ref.view(-1)).abs().max().item()
    vl, pl, ent = tr.backward(pit, zt).tolist()
    print('resnet-%d %dx%d B=%d: |dp| %.2e |dv| %.2e loss %.5f vs %.5f; vs the emulation |dp| %.2e |dv| %.2e' % (
        blocks, size, size, B, p_err, v_err, vl + pl, loss_ref.item(),
        (logp[:, :A].exp() - lp_emu.exp()).abs().max().item(), (v - v_emu.view(-1)).abs().max().item()))
    assert p_err < 5e-3 and v_err < 3e-2
    assert abs((vl + pl) - loss_ref.item()) < 2e-2 and abs(ent - ent_ref.item()) < 1e-2
    for blk_n, blk_r in zip(net.blocks, ref.blocks):
        for bn_n, bn_r in ((blk_n.bn1, blk_r.bn1), (blk_n.bn2, blk_r.bn2)):
            assert (bn_n.running_mean - bn_r.running_mean).abs().max().item() < 5e-3
            assert (bn_n.running_var - bn_r.running_var).abs().max().item() < 5e-3
            assert int(bn_n.num_batches_tracked) == 1
    grads = tr.grads()
    same_forward = (v - v_emu.view(-1)).abs().max().item() < 2e-4
    e_ref, e_emu, cos, ratio = {}, {}, {}, {}
    for (name, prm), (_, pe) in zip(ref.named_parameters(), emu.named_parameters()):
        g, r, q = grads[name].double(), prm.grad.double(), pe.grad.double()
        if name.startswith('blocks.') and name.endswith(('conv1.bias', 'conv2.bias')):
            # a bias in front of a BatchNorm: the true gradient is zero (autograd returns rounding noise)
            assert g.abs().max().item() == 0.0 and r.abs().max().item() < 1e-4
            continue
        e_ref[name] = ((g - r).norm() / (r.norm() + 1e-12)).item()
        e_emu[name] = ((g - q).norm() / (q.norm() + 1e-12)).item()
        cos[name] = (torch.dot(g.flatten(), q.flatten()) / (g.norm() * q.norm() + 1e-30)).item()
        ratio[name] = (g.norm() / (q.norm() + 1e-30)).item()
    short = lambda d: ' '.join('%s=%.3f' % (k.replace('blocks.', 'b'), e) for k, e in d.items())
    print('relative gradient errors vs fp32 autograd:', short(e_ref))
    print('relative gradient errors vs the bf16-storage emulation:', short(e_emu))
    if same_forward:
        assert max(e_emu.values()) < 0.03, max(e_emu.items(), key=lambda kv: kv[1])
    else:
        print('cosines:', short(cos))
        print('norm ratios:', short(ratio))
        assert min(cos.values()) > 0.93, min(cos.items(), key=lambda kv: kv[1])
        # (a two- or four-element bias gradient is a sum with heavy cancellation: direction only)
        assert all(0.85 < r < 1.18 for k, r in ratio.items() if grads[k].numel() >= 16), ratio
    assert (size, blocks, B) != (9, 1, 16) or same_forward          # at least this case takes the strict branch
    assert max(e_ref.values()) < 0.5, max(e_ref.items(), key=lambda kv: kv[1])


def test_resnet_learn_through_the_agent_tracks_autograd():
    """AlphaZeroAgent(net=ResNet).learn on the native tensor-core step for 20 steps on one batch: the loss falls, and
    its trajectory follows the fp32 autograd agent's; inference (eval-mode BatchNorm folded into the kernels) sees the
    updated weights."""
    from rlzero_b200.games.gomoku.alphazero_agent import AlphaZeroAgent
    from rlzero_b200.games.gomoku.policy_value_net import ResNetPolicyValueNet
    from rlzero_b200.learn import ResNetTrainer
    x, pi, z = _batch(9, 64, 5)
    torch.manual_seed(4)
    a = AlphaZeroAgent(9, net=ResNetPolicyValueNet(9, n_blocks=2))
    torch.manual_seed(4)
    b = AlphaZeroAgent(9, net=ResNetPolicyValueNet(9, n_blocks=2), trainer='autograd')
    assert isinstance(a.trainer, ResNetTrainer) and b.trainer is None
    p0, _ = a.policy_value(x)
    la, lb = [], []
    for _ in range(20):
        la.append(a.learn(x, pi, z)[0])
        lb.append(b.learn(x, pi, z)[0])
    assert la[-1] < la[0] - 0.3 and lb[-1] < lb[0] - 0.3
    assert max(abs(u - w) for u, w in zip(la, lb)) < 0.15
    p1, _ = a.policy_value(x)
    assert not np.allclose(p0, p1)


def test_bf16_inference_error_with_trained_weights():
    """the 1e-3 bound of the bf16 tensor-core forward was only shown at random init.  A short run of the
    whole loop -- self-play with the network, device augmentation, the native tensor-core training step, weight
    re-pack -- and then NativeForward against PyTorch fp32 on real positions: the loss falls, the policy sharpens
    (the logit range grows severalfold).  MEASURED: the action probabilities stay within 1e-3 (also over the long run,
    ResNet-10 / 15x15 / 320 steps, max |dp| 1.5e-4 -> 9.9e-4 while the logit range grows 0.5 -> 18:
    profiles/r2_run15_bf16_error_probe.log), but the VALUE output of a trained net leaves the 1e-3 band: up to 4e-3
    here after 40 steps (the value head sums 2*H*W features whose bf16 errors are correlated through the residual
    stream), 1e-2 ... 3e-2 once |v| approaches 1.  A PyTorch forward with fp32 arithmetic and bf16 STORAGE at the same
    places shows the same error, so this is the precision class of bf16 activations, not of these kernels.  The test
    asserts what holds -- |dp| < 1e-3 throughout, |dv| < 1e-3 at random init, and afterwards |dv| within a factor of
    two of that emulation's -- and DESIGN.md section 3.4 states the bound and the paths for stricter needs (mode 'f32';
    'tc32' for the reference's own network)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        'bf16_error_probe', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scripts',
                                         'bf16_error_probe.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    recs = mod.run(board=9, blocks=4, gens=3, steps=40, games=512, playouts=32, batch=1024, log=lambda s: None)
    trained = [r for r in recs if r.get('steps', 0) > 0]
    assert trained and trained[-1]['steps'] == 120
    assert trained[-1]['loss_last'] < trained[0]['loss_first'] - 0.3
    assert trained[-1]['max_logit_range'] > 2.5 * recs[0]['max_logit_range']
    print('bf16 vs fp32 after training (steps, |dp|, |dv|, emulated |dv|):',
          [(r['steps'], '%.1e' % r['max_dp'], '%.1e' % r['max_dv'], '%.1e' % r['torch_bf16_storage_emulation_max_dv'])
           for r in recs])
    assert recs[0]['max_dp'] < 1e-3 and recs[0]['max_dv'] < 1e-3          # random init: the north star's bound
    for r in recs:
        emu = r['torch_bf16_storage_emulation_max_dv']
        assert r['max_dp'] < 1e-3 and r['max_dv'] < 0.1, r
        assert r['max_dv'] <= 2.0 * emu + 1e-3 and emu <= 2.0 * r['max_dv'] + 1e-3, r
