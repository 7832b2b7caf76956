"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the
reference's hot-path files VERBATIM (from /root/reference, under
oracle/ref_shims.py) on seeded inputs with an injected dropout keep-mask.

Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_golden

Each fixture stores inputs, live weights, the reference's outputs and its
autograd gradients.  The reference has no golden vectors of its own
(SURVEY.md 8(c)); these are "outputs of the reference itself run here".
Also exports the Planetoid topologies used by configs 1 and 3 (Cora, Pubmed;
data fixtures, not source) as compact CSR arrays.
"""
from __future__ import annotations

import os
import pickle
import sys
import warnings

import numpy as np
import torch

from . import ref_shims, restate

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NC_PARAM_ORDER = ["moment_3", "sum", "sum2", "sum3", "sum4", "mean", "mean2", "mean3", "mean4",
                  "max", "max2", "max3", "max4", "min", "min2", "min3", "min4", "softmax",
                  "softmin", "std", "normalized_mean"]


def _save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def _weights_dict(w: restate.MMAConvWeights):
    d = dict(w.__dict__)
    return d


def golden_mmaconv(mma_conv, name, *, n_graphs, in_ch, out_ch, aggregators, scalers, edge_dim,
                   towers, divide_input, seed, p_keep=0.5, pre_layers=1, post_layers=1,
                   random_graph=None):
    torch.manual_seed(seed)
    if random_graph is None:
        ei, _ = restate.zinc_like_batch(n_graphs, seed=seed)
        n = int(ei.max()) + 1
    else:
        n, E = random_graph
        ei = torch.randint(0, n, (2, E))
        ei[1, ei[1] >= n - 3] = 0          # leave a few empty rows at the end
    E = ei.shape[1]
    deg = restate.degree_histogram(ei, n)
    conv = mma_conv.MMAConv(in_ch, out_ch, aggregators, scalers, deg, edge_dim=edge_dim,
                            towers=towers, pre_layers=pre_layers, post_layers=post_layers,
                            divide_input=divide_input)
    x = torch.randn(n, in_ch, requires_grad=True)
    ea = torch.randn(E, edge_dim, requires_grad=True) if edge_dim else None
    F_in = conv.F_in
    keep = (torch.rand(E, towers, F_in) < p_keep).float() / p_keep
    ref_shims.set_dropout(mma_conv, ref_shims.KeepMaskFeeder([keep]))
    y = conv(x, ei, ea)
    gy = torch.randn_like(y)
    w = restate.weights_from_module(conv, clone=False)
    params = w.tensors()
    grads = torch.autograd.grad(y, [x] + ([ea] if ea is not None else []) + params, gy)
    gi = 1 + (1 if ea is not None else 0)
    _save(name, {
        "x": x.detach(), "edge_index": ei, "edge_attr": None if ea is None else ea.detach(),
        "deg_hist": deg, "keep_bits": (keep > 0), "p_keep": p_keep, "gy": gy,
        "y": y.detach(), "gx": grads[0], "gea": grads[1] if ea is not None else None,
        "gparams": [g for g in grads[gi:]],
        "weights": _weights_dict(restate.weights_from_module(conv)),
        "ctor": dict(in_channels=in_ch, out_channels=out_ch, aggregators=aggregators,
                     scalers=scalers, edge_dim=edge_dim, towers=towers, pre_layers=pre_layers,
                     post_layers=post_layers, divide_input=divide_input),
    })
    ref_shims.set_dropout(mma_conv, None)


def synthetic_grad(out):
    """Deterministic upstream gradient (not stored, to keep fixtures small)."""
    k = torch.arange(out.numel(), dtype=torch.float64)
    return (((k * 0.6180339887) % 1.0) * 2.0 - 1.0).to(torch.float32).view_as(out)


def golden_aggregate(mma_conv, name, *, n, E, T, F_in, aggregators, scalers, seed):
    """MMAConv.aggregate called on its own (mma_conv.py:159): reaches var/std (Q6)."""
    torch.manual_seed(seed)
    index = torch.randint(0, n - 2, (E,))
    deg = torch.bincount(torch.bincount(index, minlength=n))
    conv = mma_conv.MMAConv(F_in * T, F_in * T, ["sum"], ["identity"], deg, towers=T, divide_input=True)
    conv.aggregators, conv.scalers = aggregators, scalers
    inputs = torch.randn(E, T, F_in)
    inputs[torch.rand(E, T, F_in) < 0.4] = 0.0          # exact-zero ties like always-on dropout
    inputs[torch.rand(E, T, F_in) < 0.05] = -0.0
    inputs.requires_grad_()
    out = conv.aggregate(inputs, index, dim_size=n)
    g = synthetic_grad(out)
    (gin,) = torch.autograd.grad(out, [inputs], g)
    _save(name, {"inputs": inputs.detach(), "index": index, "n": n, "aggregators": aggregators,
                 "scalers": scalers, "avg_deg": dict(conv.avg_deg), "deg_hist": deg,
                 "out": out.detach(), "ginputs": gin})   # gout = synthetic_grad(out)


def golden_nc(layers, name, *, rowptr, col, Fd, C, names, activation, p, seed, k=2):
    torch.manual_seed(seed)
    n = rowptr.numel() - 1
    add_all = [col[rowptr[i]:rowptr[i + 1]].numpy() for i in range(n)]
    adj = restate.csr_to_sparse_adj(rowptr, col, n)
    ps = {nm: torch.nn.Parameter(torch.empty(2 * Fd, Fd)) for nm in NC_PARAM_ORDER}
    W = torch.nn.Parameter(torch.empty(Fd, C))
    b = torch.nn.Parameter(torch.empty(C))
    L = layers.MMA(add_all, activation, k, Fd, C, W, b, *[ps[nm] for nm in NC_PARAM_ORDER],
                   p, names, "cpu")
    x = torch.relu(torch.randn(n, Fd))
    x[torch.rand(n, Fd) < 0.3] = 0.0                     # post-ReLU+dropout style zeros (ties in max/min)
    x.requires_grad_()
    E = col.numel()
    keeps = {nm: ((torch.rand(E, Fd) >= p).float() / (1.0 - p)) for nm in names}
    chunks = [keeps[nm][rowptr[i]:rowptr[i + 1]] for nm in names for i in range(n)]
    ref_shims.set_dropout(layers, ref_shims.KeepMaskFeeder(chunks))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = L(x, adj)
    gy = torch.randn_like(y)
    plist = [W, b] + [ps[nm] for nm in names]
    grads = torch.autograd.grad(y, [x] + plist, gy)
    _save(name, {"rowptr": rowptr, "col": col, "x": x.detach(), "names": names,
                 "activation": activation, "k": k, "p": p, "weight": W.detach(), "bias": b.detach(),
                 "masks": {nm: ps[nm].detach() for nm in names},
                 "keep_bits": {nm: keeps[nm] > 0 for nm in names}, "gy": gy, "y": y.detach(),
                 "gx": grads[0], "gweight": grads[1], "gbias": grads[2],
                 "gmasks": {nm: grads[3 + i] for i, nm in enumerate(names)}})
    ref_shims.set_dropout(layers, None)


def planetoid_csr(dataset):
    """utils.py:71 + 98-100 restated for networkx>=3: adjacency of
    nx.from_dict_of_lists(graph), rows in node order, neighbours ascending."""
    import networkx as nx
    path = os.path.join(ref_shims.REF_ROOT, "node_classification", "data", f"ind.{dataset}.graph")
    with open(path, "rb") as fh:
        graph = pickle.load(fh, encoding="latin1")
    adj = nx.adjacency_matrix(nx.from_dict_of_lists(graph)).tocsr()
    adj.sort_indices()
    return torch.from_numpy(adj.indptr.astype(np.int64)), torch.from_numpy(adj.indices.astype(np.int64))


def main():
    assert ref_shims.reference_available(), "needs /root/reference"
    mma_conv, _ = ref_shims.load_graph_regression()
    layers, _ = ref_shims.load_node_classification("cpu")

    # --- MultiMaskConv (graph_regression) -------------------------------------------------
    golden_mmaconv(mma_conv, "mmaconv_zinc.pt", n_graphs=12, in_ch=75, out_ch=75,
                   aggregators=["min", "max"], scalers=["identity", "amplification", "linear"],
                   edge_dim=50, towers=5, divide_input=False, seed=42)
    golden_mmaconv(mma_conv, "mmaconv_t1_noedge.pt", n_graphs=0, in_ch=16, out_ch=16,
                   aggregators=["mean", "sum", "min", "max"],
                   scalers=["identity", "amplification", "attenuation", "linear", "inverse_linear"],
                   edge_dim=None, towers=1, divide_input=False, seed=7, random_graph=(200, 1500))
    golden_mmaconv(mma_conv, "mmaconv_divide_prepost2.pt", n_graphs=0, in_ch=32, out_ch=24,
                   aggregators=["sum", "max"], scalers=["attenuation", "identity"],
                   edge_dim=6, towers=4, divide_input=True, seed=11, random_graph=(90, 700),
                   pre_layers=2, post_layers=2)
    golden_aggregate(mma_conv, "aggregate_all.pt", n=150, E=2000, T=2, F_in=12,
                     aggregators=["mean", "sum", "min", "max", "std", "var"],
                     scalers=["identity", "amplification", "attenuation", "linear", "inverse_linear"],
                     seed=3)
    golden_aggregate(mma_conv, "aggregate_c4_small.pt", n=100, E=1000, T=1, F_in=128,
                     aggregators=["mean", "sum", "min", "max", "std"],
                     scalers=["identity", "amplification", "attenuation", "linear"], seed=5)

    # --- masked multi-aggregator layer (node_classification) ---------------------------------
    rs = np.random.RandomState(0)
    rows = [np.unique(rs.randint(0, 70, size=rs.randint(1, 9))) for _ in range(70)]
    rowptr, col = restate.add_all_to_csr(rows)
    golden_nc(layers, "nc_small_mean.pt", rowptr=rowptr, col=col, Fd=8, C=3,
              names=["mean", "mean2"], activation="new_sigmoid", p=0.75, seed=1)
    golden_nc(layers, "nc_small_min4.pt", rowptr=rowptr, col=col, Fd=16, C=3,
              names=["min", "min2", "min3", "min4"], activation="new_sigmoid", p=0.5, seed=2)
    golden_nc(layers, "nc_small_mixed.pt", rowptr=rowptr, col=col, Fd=12, C=5,
              names=["sum", "max", "mean3", "softmax"], activation="new_sigmoid", p=0.5, seed=3)
    golden_nc(layers, "nc_small_sigmoid.pt", rowptr=rowptr, col=col, Fd=8, C=4,
              names=["max2", "min", "sum3", "mean3"], activation="sigmoid", p=0.5, seed=4)

    # --- Planetoid topologies (configs 1 and 3) + one real-Cora layer fixture --------------
    topo = {}
    for ds in ("cora", "pubmed"):
        rp, cl = planetoid_csr(ds)
        topo[ds] = {"rowptr": rp.to(torch.int32), "col": cl.to(torch.int32)}
        print(ds, "N", rp.numel() - 1, "nnz", cl.numel())
    _save("planetoid_topology.pt", topo)
    rp, cl = planetoid_csr("cora")
    golden_nc(layers, "nc_cora_mean_f8.pt", rowptr=rp, col=cl, Fd=8, C=7,
              names=["mean", "mean2"], activation="new_sigmoid", p=0.75, seed=42)


if __name__ == "__main__":
    sys.exit(main())
