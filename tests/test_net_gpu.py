"""The caller of MultiMaskConv in BASELINE config 2: `Net` of graph_regression/mma.py:62-127
(embeddings -> 4 x [MMAConv -> BatchNorm -> ReLU] -> global_add_pool -> MLP) around the drop-in layer, against the
oracle's restatement of the same op sequence on the CPU with the same weights and the same injected dropout masks."""
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

REL = 1e-5


def close(a, b, rel=REL, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= rel * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


def test_global_add_pool_sorted_and_unsorted():
    from mma_b200.graph_regression.net import global_add_pool
    g = torch.Generator().manual_seed(0)
    x = torch.randn(500, 75, generator=g)
    for batch in (torch.sort(torch.randint(0, 40, (500,), generator=g)).values, torch.randint(0, 40, (500,), generator=g)):
        xg = x.cuda().requires_grad_()
        out = global_add_pool(xg, batch.cuda(), 41)                         # graph 40 may be empty -> zero row
        ref = torch.zeros(41, 75, dtype=torch.float64).index_add_(0, batch, x.double())
        close(out, ref, what="global_add_pool")
        gy = torch.randn(41, 75, generator=g)
        (gx,) = torch.autograd.grad(out, [xg], gy.cuda())
        assert torch.equal(gx.cpu(), gy[batch]), "backward of the pool is a row gather"


def test_net_forward_backward_vs_oracle():
    from mma_b200.graph_regression.net import Net
    from oracle import restate
    torch.manual_seed(42)
    ei, batch = restate.zinc_like_batch(16, seed=3)
    n, E, G = batch.numel(), ei.shape[1], 16
    deg = restate.degree_histogram(ei, n)
    g = torch.Generator().manual_seed(5)
    x = torch.randint(0, 21, (n, 1), generator=g)
    ea = torch.randint(0, 4, (E,), generator=g)
    y = torch.randn(G, generator=g)
    net = Net(types.SimpleNamespace(mask=True), ["min", "max"], ["identity", "amplification", "linear"], deg).cuda()
    net.train()
    keeps = [(torch.rand(E, 5, 75, generator=g) < 0.5).float() * 2 for _ in net.convs]
    for conv, k in zip(net.convs, keeps):
        conv._inject_keep = k.cuda()
    out = net(x.cuda(), ei.cuda(), ea.cuda(), batch.cuda())
    loss = (out.squeeze() - y.cuda()).abs().mean()                          # mma.py:155
    # the oracle on leaf copies of the same weights
    ws = [restate.weights_from_module(c) for c in net.convs]
    emb = net.node_emb.weight.detach().cpu().clone().requires_grad_()
    eemb = net.edge_emb.weight.detach().cpu().clone().requires_grad_()
    for w in ws:
        for t in w.tensors():
            t.requires_grad_()
    # mma.py:116-127 with torch ops on the CPU; conv layers through oracle.restate.mmaconv_forward
    h = F.embedding(x.squeeze(), emb)
    e = F.embedding(ea, eemb)
    for w, bn, keep in zip(ws, net.batch_norms, keeps):
        h = restate.mmaconv_forward(w, h, ei, e, keep)
        h = F.relu(F.batch_norm(h, None, None, bn.weight.detach().cpu(), bn.bias.detach().cpu(), True, 0.1, bn.eps))
    h = torch.zeros(G, h.shape[1]).index_add_(0, batch, h)
    m = net.mlp
    cpu = lambda t: t.detach().cpu()
    h = F.relu(F.linear(h, cpu(m[0].weight), cpu(m[0].bias)))
    h = F.relu(F.linear(h, cpu(m[2].weight), cpu(m[2].bias)))
    ref = F.linear(h, cpu(m[4].weight), cpu(m[4].bias))
    ref_loss = (ref.squeeze() - y).abs().mean()
    close(out, ref, rel=5e-5, what="Net output (4 stacked layers)")
    close(loss, ref_loss, rel=5e-5, what="loss")
    # gradients: embeddings (through all four layers) and the first layer's live mask Linear (unregistered, Q1)
    lin0 = net.convs[0].pre_nns["max"][0][0].aggregation_layers["max"]
    got = torch.autograd.grad(loss, [net.node_emb.weight, net.edge_emb.weight, lin0.weight, net.convs[3].lin.weight])
    want = torch.autograd.grad(ref_loss, [emb, eemb, ws[0].pre[0][0][0], ws[3].lin[0]])
    for a, b, what in zip(got, want, ("d node_emb", "d edge_emb", "d mask Linear (layer 1)", "d lin (layer 4)")):
        close(a, b, rel=2e-4, what=what)


@pytest.mark.parametrize("n,Fd", [(2973, 75), (40, 128), (1, 8), (5000, 33)])
def test_fused_batch_norm_relu_vs_torch(n, Fd):
    """mma_bn_relu_fwd / _bwd (one kernel per direction) against F.relu(F.batch_norm(...)) in float64: training mode
    (batch statistics, running statistics updated with the unbiased variance) and eval mode (given statistics).
    reference: graph_regression/mma.py:120-121."""
    from mma_b200.graph_regression.net import batch_norm_relu
    g = torch.Generator().manual_seed(n + Fd)
    x = torch.randn(n, Fd, generator=g) * 2 + 0.5
    gy = torch.randn(n, Fd, generator=g)
    for training in (True, False):
        if training and n == 1:
            continue                                                    # torch refuses batch statistics of one row
        bn = torch.nn.BatchNorm1d(Fd).cuda()
        ref = torch.nn.BatchNorm1d(Fd).double()
        with torch.no_grad():
            for m in (bn, ref):
                m.weight.copy_(torch.linspace(0.5, 1.5, Fd)); m.bias.copy_(torch.linspace(-0.3, 0.3, Fd))
                m.running_mean.copy_(torch.linspace(-1, 1, Fd)); m.running_var.copy_(torch.linspace(0.5, 2, Fd))
        bn.train(training); ref.train(training)
        xg = x.cuda().requires_grad_()
        y = batch_norm_relu(xg, bn)
        xr = x.double().requires_grad_()
        yr = F.relu(ref(xr))
        close(y, yr, what=f"bn+relu y (training={training})")
        got = torch.autograd.grad(y, [xg, bn.weight, bn.bias], gy.cuda())
        want = torch.autograd.grad(yr, [xr, ref.weight, ref.bias], gy.double())
        for a, b, what in zip(got, want, ("dx", "dgamma", "dbeta")):
            close(a, b, rel=2e-5, what=f"bn+relu {what} (training={training})")
        close(bn.running_mean, ref.running_mean, what="running_mean")
        close(bn.running_var, ref.running_var, what="running_var")
        assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked)
        assert torch.equal(y, batch_norm_relu(xg, bn)) or training         # eval: deterministic, no state change


def test_net_eval_folds_batch_norm_and_relu_into_lin():
    """Net.eval(): BatchNorm is a per-channel affine map and is folded, with the ReLU, into the layer's `lin` GEMM
    (MMA_GEMM_RELU epilogue) -- same result and same gradients as the unfused sequence conv -> bn -> relu."""
    from mma_b200.graph_regression.net import Net
    from mma_b200.synthetic import zinc_like_batch, degree_histogram
    torch.manual_seed(1)
    ei, batch = zinc_like_batch(8, seed=4)
    n, E = batch.numel(), ei.shape[1]
    net = Net(types.SimpleNamespace(mask=True), ["min", "max"], ["identity", "amplification", "linear"],
              degree_histogram(ei, n)).cuda()
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for bn in net.batch_norms:
            bn.running_mean.copy_(torch.randn(75, generator=g) * 0.1); bn.running_var.copy_(torch.rand(75, generator=g) + 0.5)
            bn.weight.copy_(torch.rand(75, generator=g) + 0.5); bn.bias.copy_(torch.randn(75, generator=g) * 0.1)
    net.eval()
    conv, bn = net.convs[0], net.batch_norms[0]
    conv._inject_keep = ((torch.rand(E, 5, 75, generator=g) < 0.5).float() * 2).cuda()
    x = torch.randn(n, 75, generator=g).cuda().requires_grad_()
    ea = torch.randn(E, 50, generator=g).cuda()
    scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    y = conv.forward_affine_relu(x, ei.cuda(), ea, scale, shift)
    yr = F.relu(bn(conv(x, ei.cuda(), ea)))
    close(y, yr, what="folded eval BatchNorm + ReLU")
    gy = torch.randn(n, 75, generator=g).cuda()
    ga = torch.autograd.grad(y, [x, conv.lin.weight, bn.weight, bn.bias], gy)
    gb = torch.autograd.grad(yr, [x, conv.lin.weight, bn.weight, bn.bias], gy)
    for a, b, what in zip(ga, gb, ("dx", "d lin.weight", "d bn.weight", "d bn.bias")):
        close(a, b, rel=2e-5, what=what)
    # the whole net in eval mode runs (always-on dropout makes its output stochastic, Q3: shapes and finiteness only)
    xi = torch.randint(0, 21, (n, 1), generator=g).cuda()
    eai = torch.randint(0, 4, (E,), generator=g).cuda()
    out = net(xi, ei.cuda(), eai, batch.cuda())
    assert out.shape == (8, 1) and bool(torch.isfinite(out).all())
