"""CPU: the oracle restatement (oracle/restate.py, oracle/scatter_seq.c) against the golden
vectors produced by the verbatim reference (oracle/make_golden.py) -- this is what pins the
oracle ("outputs of the reference itself run here"; the reference ships no golden vectors)."""
import pytest
import torch

from conftest import load_golden
from oracle import restate, seq
from oracle.make_golden import synthetic_grad


def close(a, b, rel=1e-6, what=""):
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= rel * scale, f"{what}: {err:.3e} / {scale:.3e}"


@pytest.mark.parametrize("name", ["mmaconv_zinc.pt", "mmaconv_t1_noedge.pt", "mmaconv_divide_prepost2.pt"])
def test_mmaconv_restatement_matches_reference(name):
    gd = load_golden(name)
    w = restate.MMAConvWeights(**gd["weights"])
    for t in w.tensors():
        t.requires_grad_()
    x = gd["x"].clone().requires_grad_()
    ea = None if gd["edge_attr"] is None else gd["edge_attr"].clone().requires_grad_()
    keep = gd["keep_bits"].float() / gd["p_keep"]
    y = restate.mmaconv_forward(w, x, gd["edge_index"], ea, keep)
    assert torch.equal(y, gd["y"]), "restatement must reproduce the reference bit-for-bit on CPU"
    grads = torch.autograd.grad(y, [x] + ([ea] if ea is not None else []) + w.tensors(), gd["gy"])
    close(grads[0], gd["gx"], what="gx")
    k = 1
    if ea is not None:
        close(grads[1], gd["gea"], what="gea"); k = 2
    for a, b in zip(grads[k:], gd["gparams"]):
        close(a, b, what="gparam")


@pytest.mark.parametrize("name", ["aggregate_all.pt", "aggregate_c4_small.pt"])
def test_aggregate_restatement_and_sequential_c(name):
    gd = load_golden(name)
    inputs = gd["inputs"].clone().requires_grad_()
    out = restate.mmaconv_aggregate(inputs, gd["index"], gd["n"], gd["aggregators"], gd["scalers"], gd["avg_deg"])
    assert torch.equal(out, gd["out"])
    (g,) = torch.autograd.grad(out, [inputs], synthetic_grad(out))
    close(g, gd["ginputs"], what="ginputs")
    # sequential C ground truth for the raw reductions (identity block)
    E, T, F = gd["inputs"].shape
    flat = gd["inputs"].reshape(E, T * F)
    so = seq.mmconv_aggregate(None, None, flat, None, gd["index"], gd["index"], gd["n"], T * F)
    A = len(gd["aggregators"])
    blk = gd["out"].view(gd["n"], T, len(gd["scalers"]), A, F)[:, :, 0]
    for ai, a in enumerate(gd["aggregators"]):
        ref = blk[:, :, ai].reshape(gd["n"], T * F)
        if a in ("min", "max"):
            assert torch.equal(so[a].view(torch.int32), ref.contiguous().view(torch.int32)), a
        else:
            close(so[a], ref, what=f"seq {a}")


def test_scatter_semantics_edge_cases():
    src = torch.tensor([[0.0, 5.0], [-0.0, 5.0], [1.0, -1.0], [float("nan"), 2.0]])
    index = torch.tensor([0, 0, 2, 2])
    for is_max in (False, True):
        o, a = restate._minmax_first(src, index, 4, is_max)
        o2, a2 = seq.scatter_minmax(src, index, 4, is_max)
        assert torch.equal(a, a2)
        assert torch.equal(torch.nan_to_num(o, nan=-7).view(torch.int32), torch.nan_to_num(o2, nan=-7).view(torch.int32))
    o, a = restate._minmax_first(src, index, 4, False)
    assert a[0].tolist() == [0, 0] and a[1].tolist() == [4, 4] and o[1].tolist() == [0.0, 0.0]   # first wins; empty -> 0, arg=E
    assert not torch.signbit(o[0, 0])                                                             # +0.0 came first
    with pytest.raises(ValueError):
        restate.scatter(src, index, 0, None, 4, reduce="min2")                                    # Q6
    e = restate.scatter(torch.zeros(0, 3), torch.zeros(0, dtype=torch.int64), 0, None, 2, "mean")
    assert e.shape == (2, 3) and e.abs().sum() == 0


def test_scatter_restatement_vs_sequential_c_property():
    """Property test (hypothesis): the torch restatement of torch_scatter's reductions against the sequential C ground
    truth on random inputs drawn from a small value set -- many exact ties, +-0, NaN, empty and single-edge rows."""
    from hypothesis import given, settings, strategies as st

    vals = st.sampled_from([0.0, -0.0, 1.0, -1.0, 2.5, -2.5, 1e-30, float("nan"), 3.0, 3.0])

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 6), st.integers(0, 40), st.integers(1, 5), st.data())
    def run(n, E, F, data):
        src = torch.tensor(data.draw(st.lists(vals, min_size=E * F, max_size=E * F)), dtype=torch.float32).view(E, F)
        index = torch.tensor(data.draw(st.lists(st.integers(0, n - 1), min_size=E, max_size=E)), dtype=torch.int64)
        for is_max in (False, True):
            o, a = restate._minmax_first(src, index, n, is_max)
            o2, a2 = seq.scatter_minmax(src, index, n, is_max)
            assert torch.equal(a, a2), (src, index, is_max)
            assert torch.equal(torch.nan_to_num(o, nan=-7.0).view(torch.int32),
                               torch.nan_to_num(o2, nan=-7.0).view(torch.int32)), (src, index, is_max)
        finite = torch.nan_to_num(src, nan=0.5)
        for mean in (False, True):
            a = restate.scatter(finite, index, 0, None, n, "mean" if mean else "sum")
            b = seq.scatter_sum(finite, index, n, mean=mean)
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-37), (finite, index, mean)

    run()


@pytest.mark.parametrize("name", ["nc_small_mean.pt", "nc_small_min4.pt", "nc_small_mixed.pt",
                                  "nc_small_sigmoid.pt", "nc_cora_mean_f8.pt"])
def test_node_classification_restatement(name):
    gd = load_golden(name)
    n = gd["rowptr"].numel() - 1
    adj = restate.csr_to_sparse_adj(gd["rowptr"], gd["col"], n)
    x = gd["x"].clone().requires_grad_()
    W = gd["weight"].clone().requires_grad_()
    b = gd["bias"].clone().requires_grad_()
    masks = {k: v.clone().requires_grad_() for k, v in gd["masks"].items()}
    keeps = {k: v.float() / (1.0 - gd["p"]) for k, v in gd["keep_bits"].items()}
    y = restate.nc_forward(x, adj, gd["rowptr"], gd["col"], masks, W, b, gd["names"], gd["activation"], gd["p"], keeps)
    close(y, gd["y"], rel=2e-6, what="y")
    grads = torch.autograd.grad(y, [x, W, b] + [masks[k] for k in gd["names"]], gd["gy"])
    close(grads[0], gd["gx"], rel=2e-6, what="gx")
    close(grads[1], gd["gweight"], rel=2e-6, what="gW")
    close(grads[2], gd["gbias"], rel=2e-6, what="gb")
    for g, k in zip(grads[3:], gd["names"]):
        close(g, gd["gmasks"][k], rel=2e-6, what=f"gmask {k}")


def test_nc_reference_failure_modes():
    rowptr = torch.tensor([0, 1, 2]); col = torch.tensor([1, 0])
    adj = restate.csr_to_sparse_adj(rowptr, col, 2)
    x = torch.randn(2, 4); W = torch.randn(4, 3); M = {k: torch.randn(8, 4) for k in restate.NC_WORKING + restate.NC_BROKEN}
    with pytest.raises(RuntimeError):
        restate.nc_forward(x, adj, rowptr, col, M, W, None, ["std"], "sigmoid", 0.0)
    with pytest.raises(KeyError):
        restate.nc_forward(x, adj, rowptr, col, M, W, None, ["median"], "sigmoid", 0.0)
    with pytest.raises(RuntimeError):
        restate.nc_forward(x, adj, rowptr, col, M, W, None, ["sum", "sum2", "sum3", "sum4", "mean"], "sigmoid", 0.0)
    amp, att = restate.nc_scale_factors(2708)
    assert abs(amp[0, 0].item() - 1.0) < 1e-6 and abs(att[0, 0].item() - 1.0) < 1e-6             # Q7: every "degree" is N
