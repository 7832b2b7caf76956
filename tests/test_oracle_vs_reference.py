"""CPU, build container only: re-runs the VERBATIM reference (from /root/reference, under
oracle/ref_shims.py) side by side with the restatement on fresh random inputs.  Skipped on
the GPU box, where /root/reference does not exist (the committed golden vectors stand in)."""
import numpy as np
import pytest
import torch

from oracle import ref_shims, restate

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present")


def test_mmaconv_verbatim_vs_restatement():
    mma_conv, _ = ref_shims.load_graph_regression()
    ei, _ = restate.zinc_like_batch(6, seed=3)
    n, E = int(ei.max()) + 1, ei.shape[1]
    deg = restate.degree_histogram(ei, n)
    torch.manual_seed(0)
    conv = mma_conv.MMAConv(20, 20, ["mean", "min", "max"], ["identity", "attenuation"], deg, edge_dim=4, towers=2)
    x, ea = torch.randn(n, 20, requires_grad=True), torch.randn(E, 4)
    keep = (torch.rand(E, 2, 20) < 0.5).float() * 2
    ref_shims.set_dropout(mma_conv, ref_shims.KeepMaskFeeder([keep]))
    y = conv(x, ei, ea)
    ref_shims.set_dropout(mma_conv, None)
    y2 = restate.mmaconv_forward(restate.weights_from_module(conv), x, ei, ea, keep)
    assert torch.equal(y, y2)
    with pytest.raises(ValueError):
        bad = mma_conv.MMAConv(20, 20, ["std"], ["identity"], deg)
        bad(x, ei)


def test_nc_verbatim_vs_restatement():
    layers, _ = ref_shims.load_node_classification("cpu")
    from oracle.make_golden import NC_PARAM_ORDER
    rs = np.random.RandomState(5)
    rows = [np.unique(rs.randint(0, 40, size=rs.randint(1, 6))) for _ in range(40)]
    rowptr, col = restate.add_all_to_csr(rows)
    adj = restate.csr_to_sparse_adj(rowptr, col, 40)
    torch.manual_seed(1)
    Fd, C, names = 6, 3, ["sum2", "mean", "max", "min3"]
    ps = {nm: torch.nn.Parameter(torch.empty(2 * Fd, Fd)) for nm in NC_PARAM_ORDER}
    W, b = torch.nn.Parameter(torch.empty(Fd, C)), torch.nn.Parameter(torch.empty(C))
    L = layers.MMA(rows, "new_sigmoid", 2, Fd, C, W, b, *[ps[nm] for nm in NC_PARAM_ORDER], 0.5, names, "cpu")
    x = torch.relu(torch.randn(40, Fd))
    keeps = {nm: (torch.rand(col.numel(), Fd) < 0.5).float() * 2 for nm in names}
    chunks = [keeps[nm][rowptr[i]:rowptr[i + 1]] for nm in names for i in range(40)]
    ref_shims.set_dropout(layers, ref_shims.KeepMaskFeeder(chunks))
    y = L(x, adj)
    ref_shims.set_dropout(layers, None)
    y2 = restate.nc_forward(x, adj, rowptr, col, {k: v.detach() for k, v in ps.items()}, W.detach(), b.detach(),
                            names, "new_sigmoid", 0.5, keeps)
    assert (y - y2).abs().max().item() <= 2e-6 * y.abs().max().item()


def _live_tensors(w):
    ts = [t for layers in w.pre for pair in layers for t in pair if t is not None]
    ts += [t for layers in w.post for pair in layers for t in pair if t is not None]
    ts += [t for t in (w.enc or ()) if t is not None]
    ts += [t for t in w.lin if t is not None]
    return ts


MMACONV_CASES = [
    # in, out, aggregators, scalers, edge_dim, towers, pre_layers, post_layers, divide_input
    (12, 12, ["sum", "max"], ["identity", "amplification", "attenuation", "linear", "inverse_linear"], None, 1, 1, 1, False),
    (12, 9, ["mean", "min", "max", "sum"], ["linear"], 3, 3, 1, 1, False),
    (12, 12, ["min"], ["identity", "amplification"], 5, 4, 2, 2, True),
    (10, 10, ["max", "mean"], ["attenuation", "identity"], None, 2, 2, 1, False),
    (8, 8, ["mean", "sum", "min", "max"], ["identity", "amplification", "attenuation", "linear"], None, 1, 1, 1, False),
]


@pytest.mark.parametrize("case", MMACONV_CASES, ids=lambda c: f"in{c[0]}_out{c[1]}_T{c[5]}_pre{c[6]}_post{c[7]}_{'div' if c[8] else 'rep'}")
def test_mmaconv_verbatim_vs_restatement_forward_and_gradients(case):
    """Constructor space of mma_conv.py:47-107 (towers, divide_input, pre / post depth, edge encoder, all five
    scalers): forward identical, and the gradients of x, edge_attr and every weight the reference trains -- including
    the unregistered mask Linears of the pre_nns dict (Q1) -- equal to autograd through the restatement."""
    cin, cout, aggr, scal, edge_dim, towers, pre, post, divide = case
    mma_conv, _ = ref_shims.load_graph_regression()
    ei, _ = restate.zinc_like_batch(5, seed=11)
    n, E = int(ei.max()) + 1, ei.shape[1]
    deg = restate.degree_histogram(ei, n)
    torch.manual_seed(2)
    conv = mma_conv.MMAConv(cin, cout, aggr, scal, deg, edge_dim=edge_dim, towers=towers, pre_layers=pre,
                            post_layers=post, divide_input=divide)
    F_in = cin // towers if divide else cin
    x = torch.randn(n, cin, requires_grad=True)
    ea = torch.randn(E, edge_dim, requires_grad=True) if edge_dim else None
    keep = (torch.rand(E, towers, F_in) < 0.5).float() * 2
    cot = torch.randn(n, cout)
    w = restate.weights_from_module(conv, clone=False)          # the live parameters: both passes differentiate them
    live = _live_tensors(w)
    assert all(t.requires_grad for t in live)

    ref_shims.set_dropout(mma_conv, ref_shims.KeepMaskFeeder([keep]))
    y = conv(x, ei, ea) if edge_dim else conv(x, ei)
    ref_shims.set_dropout(mma_conv, None)
    g_ref = torch.autograd.grad((y * cot).sum(), [x] + ([ea] if edge_dim else []) + live, allow_unused=True)

    y2 = restate.mmaconv_forward(w, x, ei, ea, keep)
    g_res = torch.autograd.grad((y2 * cot).sum(), [x] + ([ea] if edge_dim else []) + live, allow_unused=True)

    assert torch.equal(y, y2)
    for a, b in zip(g_ref, g_res):
        assert (a is None) == (b is None)
        if a is not None:
            assert (a - b).abs().max().item() <= 1e-6 * max(1.0, a.abs().max().item())


NC_CASES = [(grp, act) for act in ("new_sigmoid", "sigmoid")
            for grp in (["sum", "sum2", "sum3", "sum4"], ["mean", "mean2", "mean3", "mean4"], ["max", "max2", "max3", "max4"],
                        ["min", "min2", "min3", "min4"], ["softmax"], ["softmin"], ["mean3", "max", "min", "sum"])]


@pytest.mark.parametrize("names,activation", NC_CASES, ids=lambda v: "-".join(v) if isinstance(v, list) else v)
def test_nc_verbatim_vs_restatement_every_aggregator_with_gradients(names, activation):
    """All 18 working aggregators of layers.py:201-728 under both activations (`new_sigmoid` makes mean3 / max / min use
    the RAW mask, Q8), injected dropout, ties in x: output and the gradients of x, every mask in use, the class weight
    and the bias against autograd through the restatement."""
    layers, _ = ref_shims.load_node_classification("cpu")
    from oracle.make_golden import NC_PARAM_ORDER
    n, Fd, C = 30, 5, 4
    rs = np.random.RandomState(7)
    rows = [np.unique(rs.randint(0, n, size=rs.randint(1, 7))) for _ in range(n)]
    rowptr, col = restate.add_all_to_csr(rows)
    adj = restate.csr_to_sparse_adj(rowptr, col, n)
    torch.manual_seed(3)
    ps = {nm: torch.nn.Parameter(torch.empty(2 * Fd, Fd)) for nm in NC_PARAM_ORDER}
    W, b = torch.nn.Parameter(torch.empty(Fd, C)), torch.nn.Parameter(torch.empty(C))
    L = layers.MMA(rows, activation, 2, Fd, C, W, b, *[ps[nm] for nm in NC_PARAM_ORDER], 0.5, names, "cpu")
    x = torch.relu(torch.randn(n, Fd))
    x[torch.rand(n, Fd) < 0.3] = 0.0                              # ties for the elementwise max / min with x_i
    x.requires_grad_()
    keeps = {nm: (torch.rand(col.numel(), Fd) < 0.5).float() * 2 for nm in names}
    chunks = [keeps[nm][rowptr[i]:rowptr[i + 1]] for nm in names for i in range(n)]
    cot = torch.randn(n, C)
    wrt = [x, W, b] + [ps[nm] for nm in names]

    ref_shims.set_dropout(layers, ref_shims.KeepMaskFeeder(chunks))
    y = L(x, adj)
    ref_shims.set_dropout(layers, None)
    g_ref = torch.autograd.grad((y * cot).sum(), wrt)

    y2 = restate.nc_forward(x, adj, rowptr, col, ps, W, b, names, activation, 0.5, keeps)
    g_res = torch.autograd.grad((y2 * cot).sum(), wrt)

    assert (y - y2).abs().max().item() <= 2e-6 * max(1.0, y.abs().max().item())
    for a, g in zip(g_ref, g_res):
        assert (a - g).abs().max().item() <= 5e-6 * max(1.0, a.abs().max().item())
