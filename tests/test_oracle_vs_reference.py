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
