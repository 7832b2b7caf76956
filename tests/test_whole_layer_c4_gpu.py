"""Whole-layer parity AT THE BENCHMARK CONFIGURATION (BASELINE config 4's layer: hidden 128, towers 1,
aggregators mean,sum,min,max,std x scalers identity,amplification,attenuation,linear) on a graph small enough for
the oracle: forward output and EVERY gradient (x, the unregistered mask Linear, post Linear, lin) of the drop-in
`MMAConv` -- the fused autograd node of fused_layer.py: tensor-memory-resident and streaming tcgen05 3xTF32 GEMMs,
grouped post transform with folded scalers, K1, the transpose pass -- against the op-for-op restatement of
mma_conv.py:121-196 (oracle/restate.py, test infrastructure), tolerance 1e-5 of each tensor's largest magnitude
(BASELINE.json north_star).

Min/max selections of the reference are discontinuous in the messages (tests/near_ties.py: a near-tie flips the arg
edge under ANY change of fp32 summation order and moves the gradients by ~1e-3), and at the layer level the CUDA
path's messages come from another evaluation order (P[dst] + Q[src] on 3xTF32 tensor cores instead of one sgemm
over cat([x_i, x_j]), mma_conv.py:146-152).  The case therefore makes the MESSAGES exact: x holds small integers
and the mask Linear multiples of 1/16, so every product and partial sum of the mask projection is exactly
representable in tf32 / fp32 on both sides -- the messages are bit-identical, exact ties (plentiful on such a grid)
are resolved by the same first-occurrence rule, and near-ties cannot occur (asserted on the oracle's side).  All other
weights are the layer's own random initialisers, so everything after the messages is ordinary fp32.

Two dropout modes:
* "inject": an explicit keep mask on both sides (K1's keep-mask kernels).
* "philox": the in-kernel Philox stream of the persistent stream kernels -- the path bench.py times -- replayed
  into the oracle through `dropout_keep_scale` with the seed the layer used.

Measured on a B200 (tests/tools/whole_layer_errors.py): output 1.5e-6 / 1.8e-6 (inject / philox), dx 1.5e-6 / 1.6e-6,
mask weight 9e-7 / 1.5e-6, mask bias 1.6e-6 / 1e-6, post weight 6e-7 / 9e-7, lin weight 7e-7 / 1e-6, biases 2-3e-7.
"""
import pytest
import torch

from near_ties import ambiguous_entries

AGGR = ["mean", "sum", "min", "max", "std"]
SCAL = ["identity", "amplification", "attenuation", "linear"]
REL = 1e-5
INJECT_CASE = dict(seed=7, n=500, E=8000, F=128)
PHILOX_CASE = dict(seed=11, n=1500, E=24000, F=128)


def build_case(seed, n, E, F):
    """CPU: graph with 3 empty rows, a few duplicate edges and self loops (uniform ids), the layer with its own
    initialisers under the seed, input, upstream gradient, keep mask."""
    from mma_b200 import MMAConv
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (E,), generator=g)
    dst = torch.randint(0, n - 3, (E,), generator=g)
    hist = torch.bincount(torch.bincount(dst, minlength=n))
    torch.manual_seed(seed)
    conv = MMAConv(F, F, AGGR, SCAL, hist, towers=1, strict_reference=False)
    with torch.no_grad():                        # exact messages: integers x multiples of 1/16 (see the module docstring)
        for p in conv.mask_parameters():
            p.copy_(torch.randint(-4, 5, p.shape, generator=g).float() / 16)
    x = torch.randint(-3, 4, (n, F), generator=g).float()
    gy = torch.randn(n, F, generator=g)
    keep = (torch.rand(E, 1, F, generator=g) < 0.5).float() * 2
    return conv, src, dst, x, gy, keep


def oracle_run(conv, src, dst, x, gy, keep):
    """Reference op sequence on CPU: output, gradients (x first, then MMAConvWeights.tensors() order) and the
    numerically ambiguous min/max selections."""
    from oracle import restate
    n, E, F = x.shape[0], src.numel(), x.shape[1]
    ei = torch.stack([src, dst])
    w = restate.weights_from_module(conv)
    for t in w.tensors():
        t.requires_grad_()
    xr = x.clone().requires_grad_()
    yr = restate.mmaconv_forward(w, xr, ei, None, keep, strict=False)
    grads = torch.autograd.grad(yr, [xr] + w.tensors(), gy)
    with torch.no_grad():
        xt = x.view(n, 1, F)
        msg = restate.mmaconv_message(w, xt.index_select(0, dst), xt.index_select(0, src), None, keep, False).view(E, F)
        _, amin = restate.scatter_with_arg(msg, dst, n, "min")
        _, amax = restate.scatter_with_arg(msg, dst, n, "max")
        rows, edges = ambiguous_entries(msg, dst, n, amin, amax, tol=REL * msg.abs().max().item())
    return yr.detach(), [t.detach() for t in grads], rows, edges


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["inject", "philox"])
def test_config4_layer_forward_and_all_gradients_vs_oracle(mode):
    import mma_b200
    from mma_b200 import Graph
    from oracle import restate
    case = INJECT_CASE if mode == "inject" else PHILOX_CASE
    conv, src, dst, x, gy, keep = build_case(**case)
    n, E, F = case["n"], case["E"], case["F"]
    conv = conv.cuda()
    conv.fold_min_rows = 16                      # both the grouped tcgen05 GEMM (big degree ranges) and the tail path
    graph = Graph(src.cuda(), dst.cuda(), n, sort_rows=True)
    xg = x.cuda().requires_grad_()
    if mode == "inject":
        conv._inject_keep = keep.cuda()
    y = conv(xg, graph)
    if mode == "philox":
        assert conv.dropout == 0.5 and conv.last_seed is not None
        keep = mma_b200.dropout_keep_scale(0.5, conv.last_seed, E, F, "cuda", graph=graph).cpu().view(E, 1, F)
        assert abs((keep > 0).float().mean().item() - 0.5) < 0.01
    params = restate.weights_from_module(conv, clone=False).tensors()
    got = torch.autograd.grad(y, [xg] + params, gy.cuda())
    yr, ref, rows, _ = oracle_run(conv, src, dst, x, gy, keep)

    assert rel_err(y, yr) <= REL, f"{mode}: layer output rel err {rel_err(y, yr):.3e}"
    assert not rows, ("near-ties among exact messages", rows[:5])
    names = ["dx", "d mask W", "d mask b", "d post W", "d post b", "d lin W", "d lin b"]
    assert len(got) == len(ref) == len(names)
    for name, a, b in zip(names, got, ref):
        assert rel_err(a, b) <= REL, f"{mode}: {name} rel err {rel_err(a, b):.3e}"


def test_near_tie_finder_cpu():
    """CPU check of the helper on a hand-made case: one genuine near-tie, one exact tie (not ambiguous), one
    dropped-zero tie (not ambiguous), one clear winner."""
    from oracle import restate
    index = torch.tensor([0, 0, 1, 1, 2, 2, 3, 3])
    m = torch.tensor([[1.0], [1.0 + 2e-7], [0.5], [0.5], [-0.0], [0.0], [3.0], [1.0]])
    _, amin = restate.scatter_with_arg(m, index, 5, "min")
    _, amax = restate.scatter_with_arg(m, index, 5, "max")
    rows, edges = ambiguous_entries(m, index, 5, amin, amax, tol=1e-5)
    assert rows == [0] and edges == [0, 1]
    assert amax[4, 0] == 8 and amin[1, 0] == 2 and amax[2, 0] == 4     # empty row; first occurrence wins exact ties
