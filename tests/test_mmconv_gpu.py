"""Parity of the CUDA MultiMaskConv path (K1 + drop-in MMAConv) against the oracle and the
golden vectors produced by the verbatim reference.  Bar (BASELINE.json north_star):
min/max selections and arg indices BIT-EXACT; sum/mean/std and all gradients within 1e-5
relative error (fp32), measured against the largest magnitude of the reference tensor."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

REL = 1e-5      # north_star tolerance for sum/mean/std/gradients


def close(a, b, rel=REL, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if b.numel() == 0:
        return
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= rel * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


def bitexact(a, b, what=""):
    a, b = a.detach().cpu().contiguous(), b.detach().cpu().contiguous()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)), f"{what}: not bit-exact"


def rand_graph(n, E, seed, empty_tail=3):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (E,), generator=g)
    dst = torch.randint(0, max(n - empty_tail, 1), (E,), generator=g)
    return src, dst


ALL_AGGR = ["mean", "sum", "min", "max", "std", "var"]
ALL_SCALE = ["identity", "amplification", "attenuation", "linear", "inverse_linear"]


@pytest.mark.parametrize("n,E,T,F_in,use", [
    (300, 4000, 1, 128, "PQ"),       # warp per row, float4
    (257, 3000, 1, 64, "PQR"),       # 16 lanes per row
    (100, 900, 2, 8, "PQRK"),        # 4 lanes per row, towers, explicit keep
    (64, 700, 5, 75, "PQRK"),        # c2 shape: F_in % 4 != 0 -> scalar path, 12 chunks
    (50, 600, 1, 200, "QK"),         # F > 128: two float4 chunks
    (40, 300, 1, 16, "R"),           # materialised messages only
    (30, 0, 1, 16, "PQ"),            # no edges at all
    (1, 50, 1, 4, "PQ"),             # single node, self loops / multi-edges
])
def test_fused_aggregate_vs_oracle(n, E, T, F_in, use):
    import mma_b200
    from oracle import restate, seq
    F = T * F_in
    src, dst = rand_graph(n, E, seed=n + E, empty_tail=3 if n > 4 else 0)
    g = torch.Generator().manual_seed(1)
    def mk(rows):
        t = torch.randn(rows, F, generator=g)
        t[torch.rand(rows, F, generator=g) < 0.3] = 0.0       # exact ties
        t[torch.rand(rows, F, generator=g) < 0.05] = -0.0     # signed zeros
        return t
    P = mk(n) if "P" in use else None
    Q = mk(n) if "Q" in use else None
    R = mk(E) if "R" in use else None
    keep = (torch.rand(E, F, generator=g) < 0.5).float() * 2 if "K" in use else None
    hist = torch.bincount(torch.bincount(dst, minlength=n)) if E else torch.tensor([n])
    avg = restate.avg_deg_from_hist(hist)

    dev = "cuda"
    graph = mma_b200.Graph(src.to(dev), dst.to(dev), n)
    gl = [t.clone().to(dev).requires_grad_() for t in (P, Q, R) if t is not None]
    it = iter(gl)
    Pg, Qg, Rg = (next(it) if t is not None else None for t in (P, Q, R))
    Y, amin, amax = mma_b200.mmconv_aggregate(Pg, Qg, Rg, graph, towers=T, F_in=F_in, aggregators=ALL_AGGR,
                                              scalers=ALL_SCALE, avg_deg=avg,
                                              keep=None if keep is None else keep.to(dev), return_args=True)
    assert Y.shape == (n, T, len(ALL_SCALE) * len(ALL_AGGR) * F_in)

    # the oracle's layout is [n, S*A*F] with F the flat width; re-tile to [n,T,(s,a),F_in]
    ref_flat, rargs = restate.mmconv_fused_op(P, Q, R, keep, src, dst, n, ALL_AGGR, ALL_SCALE, avg, return_args=True)
    S, A = len(ALL_SCALE), len(ALL_AGGR)
    ref_t = ref_flat.view(n, S * A, T, F_in).permute(0, 2, 1, 3).reshape(n, T, S * A * F_in)
    Yc = Y.detach().cpu()
    Yv, Rv = Yc.view(n, T, S, A, F_in), ref_t.view(n, T, S, A, F_in)
    for ai, name in enumerate(ALL_AGGR):
        if name in ("min", "max"):
            bitexact(Yv[:, :, :, ai], Rv[:, :, :, ai], f"{name} (all scaler blocks)")
        else:
            close(Yv[:, :, :, ai], Rv[:, :, :, ai], what=name)
    assert torch.equal(amin.cpu().long(), rargs["min"].view(n, F)), "argmin indices"
    assert torch.equal(amax.cpu().long(), rargs["max"].view(n, F)), "argmax indices"

    # sequential C ground truth (tie-breaking order) on the raw aggregates
    if E > 0:
        so = seq.mmconv_aggregate(P, Q, R, keep, src, dst, n, F)
        assert torch.equal(amin.cpu().long(), so["arg_min"]) and torch.equal(amax.cpu().long(), so["arg_max"])
        id_block = Yc.view(n, T, S, A, F_in)[:, :, 0]                        # identity scaler block
        for ai, name in enumerate(ALL_AGGR):
            got = id_block[:, :, ai].reshape(n, F)
            if name in ("min", "max"):
                bitexact(got, so[name], f"seq {name}")
            else:
                close(got, so[name], what=f"seq {name}")

    # gradients (deterministic upstream gradient)
    gy = torch.cos(torch.arange(Y.numel(), dtype=torch.float32) * 0.37).view_as(Y)
    gy_ref = gy.view(n, T, S * A, F_in).permute(0, 2, 1, 3).reshape(n, -1)
    leaves_ref = [t.clone().requires_grad_() for t in (P, Q, R) if t is not None]
    it = iter(leaves_ref)
    Pr, Qr, Rr = (next(it) if t is not None else None for t in (P, Q, R))
    out = restate.mmconv_fused_op(Pr, Qr, Rr, keep, src, dst, n, ALL_AGGR, ALL_SCALE, avg)
    gref = torch.autograd.grad(out, leaves_ref, gy_ref)
    ggot = torch.autograd.grad(Y, gl, gy.to(dev))
    for name, a, b in zip([k for k, t in zip("PQR", (P, Q, R)) if t is not None], ggot, gref):
        close(a, b, what=f"d{name}")


def test_determinism_and_edge_order_invariance():
    import mma_b200
    n, E, F = 500, 9000, 64
    src, dst = rand_graph(n, E, 5)
    P, Q = torch.randn(n, F).cuda(), torch.randn(n, F).cuda()
    hist = torch.bincount(torch.bincount(dst, minlength=n))
    from oracle import restate
    avg = restate.avg_deg_from_hist(hist)
    def run(s, d):
        g = mma_b200.Graph(s.cuda(), d.cuda(), n)
        Pg, Qg = P.clone().requires_grad_(), Q.clone().requires_grad_()
        Y, amin, amax = mma_b200.mmconv_aggregate(Pg, Qg, None, g, towers=1, F_in=F, aggregators=ALL_AGGR[:5],
                                                  scalers=ALL_SCALE[:4], avg_deg=avg, return_args=True)
        gP, gQ = torch.autograd.grad(Y, [Pg, Qg], torch.ones_like(Y))
        return Y, amin, amax, gP, gQ
    a, b = run(src, dst), run(src, dst)
    for x, y in zip(a, b):
        assert torch.equal(x, y), "two identical runs differ (must be bit-reproducible: no atomics)"
    # permuting the edge list: min/max VALUES stay bit-exact, sums within tolerance
    perm = torch.randperm(E, generator=torch.Generator().manual_seed(0))
    c = run(src[perm], dst[perm])
    S, A = 4, 5
    Ya, Yc = a[0].view(n, S, A, F), c[0].view(n, S, A, F)
    bitexact(Ya[:, :, 2:4], Yc[:, :, 2:4], "min/max under edge permutation")
    close(Ya, Yc, what="sum/mean/std under edge permutation")
    # arg follows the lowest ORIGINAL edge id among ties: map back through the permutation
    m = (P.cpu()[dst] + Q.cpu()[src])
    amin = a[1].cpu().long()
    rows = torch.arange(n).view(n, 1).expand(n, F)
    valid = amin < E
    picked = m[amin.clamp(max=E - 1), torch.arange(F).view(1, F).expand(n, F)]
    assert torch.equal(picked[valid], Ya[:, 0, 2].cpu()[valid]), "m[argmin] must equal the min value"
    assert torch.equal(dst[amin.clamp(max=E - 1)][valid], rows[valid]), "argmin edge must end at its row"


def test_philox_dropout_matches_injected_mask():
    import mma_b200
    from oracle import restate
    n, E, F, p = 200, 3000, 32, 0.5
    src, dst = rand_graph(n, E, 9)
    P, Q = torch.randn(n, F), torch.randn(n, F)
    g = mma_b200.Graph(src.cuda(), dst.cuda(), n)
    seed = 1234567
    keep = mma_b200.dropout_keep_scale(p, seed, E, F, "cuda", graph=g)
    frac = (keep > 0).float().mean().item()
    assert abs(frac - (1 - p)) < 0.01 and set(keep.unique().tolist()) == {0.0, 2.0}
    Pg, Qg = P.cuda().requires_grad_(), Q.cuda().requires_grad_()
    Y = mma_b200.mmconv_aggregate(Pg, Qg, None, g, towers=1, F_in=F, aggregators=["sum", "max", "std"],
                                  scalers=["identity"], p_drop=p, seed=seed)
    Pr, Qr = P.clone().requires_grad_(), Q.clone().requires_grad_()
    ref = restate.mmconv_fused_op(Pr, Qr, None, keep.cpu(), src, dst, n, ["sum", "max", "std"], ["identity"], {})
    close(Y.view(n, -1), ref, what="philox fwd")
    bitexact(Y.view(n, 3, F)[:, 1], ref.view(n, 3, F)[:, 1], "max under philox dropout")
    gy = torch.randn(n, 3 * F)
    a = torch.autograd.grad(Y, [Pg, Qg], gy.cuda().view_as(Y))
    b = torch.autograd.grad(ref, [Pr, Qr], gy)
    close(a[0], b[0], what="philox dP"); close(a[1], b[1], what="philox dQ")
    # different seeds / p
    k2 = mma_b200.dropout_keep_scale(p, seed + 1, E, F, "cuda", graph=g)
    assert not torch.equal(keep, k2)
    k3 = mma_b200.dropout_keep_scale(0.75, seed, E, F, "cuda", graph=g)
    assert abs((k3 > 0).float().mean().item() - 0.25) < 0.01 and k3.max().item() == 4.0


@pytest.mark.parametrize("name", ["aggregate_all.pt", "aggregate_c4_small.pt"])
def test_aggregate_method_vs_reference_golden(name):
    """MMAConv.aggregate on its own (mma_conv.py:159) against the verbatim reference's output."""
    from mma_b200 import MMAConv
    from oracle.make_golden import synthetic_grad
    gd = load_golden(name)
    inputs, index, n = gd["inputs"], gd["index"], gd["n"]
    E, T, F_in = inputs.shape
    conv = MMAConv(T * F_in, T * F_in, ["sum"], ["identity"], gd["deg_hist"], towers=T, divide_input=True)
    conv.aggregators, conv.scalers = gd["aggregators"], gd["scalers"]
    assert conv.avg_deg == gd["avg_deg"]
    x = inputs.cuda().requires_grad_()
    out = conv.aggregate(x, index.cuda(), dim_size=n)
    A, S = len(gd["aggregators"]), len(gd["scalers"])
    ov, rv = out.detach().cpu().view(n, T, S, A, F_in), gd["out"].view(n, T, S, A, F_in)
    for ai, a in enumerate(gd["aggregators"]):
        if a in ("min", "max"):
            bitexact(ov[:, :, :, ai], rv[:, :, :, ai], f"{name}:{a}")
        else:
            close(ov[:, :, :, ai], rv[:, :, :, ai], what=f"{name}:{a}")
    (gin,) = torch.autograd.grad(out, [x], synthetic_grad(gd["out"]).cuda())
    close(gin, gd["ginputs"], what=f"{name}: d inputs")


def _load_weights(conv, w):
    with torch.no_grad():
        if w["enc"] is not None:
            conv.edge_encoder.weight.copy_(w["enc"][0]); conv.edge_encoder.bias.copy_(w["enc"][1])
        a_star = conv.aggregators[-1]
        for t, seqm in enumerate(conv.pre_nns[a_star]):
            lins = [m for m in seqm if hasattr(m, "aggregation_layers")]
            for li, m in enumerate(lins):
                m.live().weight.copy_(w["pre"][t][li][0]); m.live().bias.copy_(w["pre"][t][li][1])
        for t, seqm in enumerate(conv.post_nns):
            lins = [m for m in seqm if hasattr(m, "weight")]
            for li, m in enumerate(lins):
                m.weight.copy_(w["post"][t][li][0]); m.bias.copy_(w["post"][t][li][1])
        conv.lin.weight.copy_(w["lin"][0]); conv.lin.bias.copy_(w["lin"][1])


@pytest.mark.parametrize("name,fold,tc", [("mmaconv_zinc.pt", None, False), ("mmaconv_t1_noedge.pt", 512, False),
                                          ("mmaconv_t1_noedge.pt", 2, False), ("mmaconv_t1_noedge.pt", 0, False),
                                          ("mmaconv_t1_noedge.pt", 512, True), ("mmaconv_t1_noedge.pt", 1, True),
                                          ("mmaconv_divide_prepost2.pt", None, False)])
def test_mmaconv_layer_vs_reference_golden(name, fold, tc):
    """Whole drop-in layer (fwd + all gradients) against the verbatim reference's outputs.
    fold: towers == 1 only -- 0 = materialised scaler blocks, k = scalers folded into the post weight
    with degree ranges of >= k rows as one GEMM each (512: everything through the literal tail path)."""
    from mma_b200 import MMAConv
    from oracle import restate
    gd = load_golden(name)
    conv = MMAConv(deg=gd["deg_hist"], **gd["ctor"]).cuda()
    if fold is not None:
        conv.fold_scalers, conv.fold_min_rows = fold > 0, max(fold, 1)
    conv.use_tensor_cores = tc      # tc: the whole layer as one autograd node over the tcgen05 3xTF32 GEMMs
    _load_weights(conv, gd["weights"])
    assert conv.avg_deg == gd["weights"]["avg_deg"]
    x = gd["x"].cuda().requires_grad_()
    ea = None if gd["edge_attr"] is None else gd["edge_attr"].cuda().requires_grad_()
    conv._inject_keep = gd["keep_bits"].float().cuda() / gd["p_keep"]
    y = conv(x, gd["edge_index"].cuda(), ea)
    close(y, gd["y"], what=f"{name}: y")
    params = restate.weights_from_module(conv, clone=False).tensors()
    grads = torch.autograd.grad(y, [x] + ([ea] if ea is not None else []) + params, gd["gy"].cuda())
    close(grads[0], gd["gx"], what="dx")
    k = 1
    if ea is not None:
        close(grads[1], gd["gea"], what="d edge_attr"); k = 2
    for i, (a, b) in enumerate(zip(grads[k:], gd["gparams"])):
        close(a, b, what=f"d param {i}")


def test_mmaconv_api_and_errors():
    from mma_b200 import MMAConv, MaskAggregateLinear
    deg = torch.tensor([0, 3, 5, 2])
    conv = MMAConv(8, 8, ["mean", "max"], ["identity", "amplification"], deg, edge_dim=3, towers=2)
    assert conv.dropout == 0.5 and conv.F_in == 8 and conv.F_out == 4
    assert set(conv.pre_nns) == {"mean", "max"} and isinstance(conv.pre_nns, dict)
    names = [n for n, _ in conv.named_parameters()]
    assert not any("pre_nns" in n or "aggregation_layers" in n for n in names)      # Q1
    assert isinstance(conv.pre_nns["max"][0][0], MaskAggregateLinear)
    assert len(conv.mask_parameters()) == 2 * 2
    x = torch.randn(6, 8).cuda(); ei = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 0]]).cuda(); ea = torch.randn(4, 3).cuda()
    conv = conv.cuda()
    assert conv(x, ei, ea).shape == (6, 8)
    with pytest.raises(RuntimeError):
        conv.cpu()(x.cpu(), ei.cpu(), ea.cpu())                                      # no CPU fallback
    conv = conv.cuda()
    bad = MMAConv(8, 8, ["mean", "std"], ["identity"], deg).cuda()
    with pytest.raises(ValueError, match='Unknown aggregator "std"'):                # Q6
        bad(x, ei)
    ok = MMAConv(8, 8, ["mean", "std"], ["identity"], deg, strict_reference=False).cuda()
    assert ok(x, ei).shape == (6, 8)
    bad2 = MMAConv(8, 8, ["min2"], ["identity"], deg).cuda()
    with pytest.raises(ValueError):
        bad2(x, ei)
    bad3 = MMAConv(8, 8, ["min"], ["squash"], deg).cuda()
    with pytest.raises(ValueError, match='Unknown scaler "squash"'):
        bad3(x, ei)
    # stochastic in eval mode too (Q3), reproducible per seed
    conv.eval()
    torch.manual_seed(3); conv._calls = 0; a = conv(x, ei, ea)
    b = conv(x, ei, ea)
    torch.manual_seed(3); conv._calls = 0; c = conv(x, ei, ea)
    assert not torch.equal(a, b) and torch.equal(a, c)


@pytest.mark.parametrize("edge_dim,tc", [(None, False), (6, False), (None, True), (6, True)])
def test_folded_post_transform_vs_oracle(edge_dim, tc):
    """towers == 1 fast path (raw aggregates in degree-sorted rows + per-degree effective post
    weight) on a graph with many distinct degrees, vs the op-for-op oracle of the reference."""
    from mma_b200 import MMAConv
    from oracle import restate
    n, E, Fd = 3000, 40000, 32
    g = torch.Generator().manual_seed(3)
    src = torch.randint(0, n, (E,), generator=g)
    dst = (torch.rand(E, generator=g) ** 2 * (n - 5)).long()          # skewed in-degrees, a few empty rows
    ei = torch.stack([src, dst])
    hist = restate.degree_histogram(ei, n)
    torch.manual_seed(0)
    aggr, scal = ["mean", "sum", "min", "max"], ["identity", "amplification", "attenuation", "linear", "inverse_linear"]
    conv = MMAConv(Fd, Fd, aggr, scal, hist, edge_dim=edge_dim, towers=1).cuda()
    conv.fold_min_rows = 16
    conv.use_tensor_cores = tc
    x = torch.randn(n, Fd, generator=g)
    ea = torch.randn(E, edge_dim, generator=g) if edge_dim else None
    keep = (torch.rand(E, 1, Fd, generator=g) < 0.5).float() * 2
    conv._inject_keep = keep.cuda()
    xg = x.cuda().requires_grad_()
    eag = None if ea is None else ea.cuda().requires_grad_()
    y = conv(xg, ei.cuda(), eag)
    w = restate.weights_from_module(conv)
    for t in w.tensors():
        t.requires_grad_()
    xr = x.clone().requires_grad_()
    ear = None if ea is None else ea.clone().requires_grad_()
    yr = restate.mmaconv_forward(w, xr, ei, ear, keep)
    close(y, yr, what="folded y")
    gy = torch.randn(n, Fd, generator=g)
    params = restate.weights_from_module(conv, clone=False).tensors()
    ins = [xg] + ([eag] if ea is not None else [])
    insr = [xr] + ([ear] if ea is not None else [])
    got = torch.autograd.grad(y, ins + params, gy.cuda())
    ref = torch.autograd.grad(yr, insr + w.tensors(), gy)
    for i, (a, b) in enumerate(zip(got, ref)):
        close(a, b, what=f"folded grad {i}")
    # unfolded path gives the same numbers to fp32 rounding
    conv.fold_scalers = False
    y2 = conv(xg, ei.cuda(), eag)
    close(y2, yr, what="unfolded y")


@pytest.mark.parametrize("F,p,seg", [(128, 0.5, 4096), (64, 0.5, 64), (128, 0.25, 96), (32, 0.0, 32), (128, 0.5, 64),
                                     (128, 0.5, 1 << 20)])
def test_stream_kernels_long_rows_and_hubs(F, p, seg, monkeypatch):
    """Skewed degrees: a few hub rows (thousands of in-edges: many index chunks, ring wrap-arounds, dropout words
    refreshed every 32 edges), runs of empty rows, rows of every small length -- forward (bit-exact min/max +
    arg indices) and backward of the persistent stream kernels against the oracle fed the same dropout mask.
    seg: rows longer than this are cut into segments walked by different warps and merged by a second kernel
    (4096 is the default: only the largest hub splits; 64 splits dozens of rows; 2^20: no row splits)."""
    import mma_b200
    from oracle import restate
    monkeypatch.setattr(mma_b200.Graph, "K1_MAX_SEG", seg)
    g = torch.Generator().manual_seed(F + int(p * 100))
    n = 700
    deg = torch.cat([torch.tensor([5000, 1537, 260, 129, 97, 65, 64, 63, 33, 32, 31]),
                     torch.randint(0, 9, (n - 11 - 40,), generator=g), torch.zeros(40, dtype=torch.long)])
    deg = deg[torch.randperm(n, generator=g)]
    dst = torch.repeat_interleave(torch.arange(n), deg)
    E = dst.numel()
    shuffle = torch.randperm(E, generator=g)
    dst = dst[shuffle]
    src = torch.randint(0, n, (E,), generator=g)
    P, Q = torch.randn(n, F, generator=g), torch.randn(n, F, generator=g)
    P[torch.rand(n, F, generator=g) < 0.2] = 0.0; Q[torch.rand(n, F, generator=g) < 0.2] = 0.0     # exact ties
    aggr = ["mean", "sum", "min", "max", "std"]
    for sort_rows in (False, True):
        graph = mma_b200.Graph(src.cuda(), dst.cuda(), n, sort_rows=sort_rows)
        seed = 99
        keep = mma_b200.dropout_keep_scale(p, seed, E, F, "cuda", graph=graph).cpu() if p > 0 else None
        if keep is not None:
            assert abs((keep > 0).float().mean().item() - (1 - p)) < 0.02
        Pg, Qg = P.cuda().requires_grad_(), Q.cuda().requires_grad_()
        Y, amin, amax = mma_b200.mmconv_aggregate(Pg, Qg, None, graph, towers=1, F_in=F, aggregators=aggr,
                                                  scalers=["identity"], p_drop=p, seed=seed, return_args=True)
        P32, Q32 = P.clone().requires_grad_(), Q.clone().requires_grad_()
        ref, rargs = restate.mmconv_fused_op(P32, Q32, None, keep, src, dst, n, aggr, ["identity"], {}, return_args=True)
        # sums over thousands of edges: fp32 accumulation alone is only good to ~1e-5 on the hub rows (the
        # reference's own fp32 result is that far from exact), so sums and gradients are judged against the same
        # op evaluated in float64: within 1e-5, or no further from it than twice the reference's fp32 error
        Pr, Qr = P.double().requires_grad_(), Q.double().requires_grad_()
        ref64 = restate.mmconv_fused_op(Pr, Qr, None, None if keep is None else keep.double(), src, dst, n, aggr,
                                        ["identity"], {})
        # with sorted rows Y is in CSR-row order: bring it back to node order
        Yn = Y.view(n, -1)
        if sort_rows:
            inv = graph.row_map.long()
            Yn = torch.empty_like(Yn).index_copy_(0, inv, Yn)
            amin = torch.empty_like(amin).index_copy_(0, inv, amin)
            amax = torch.empty_like(amax).index_copy_(0, inv, amax)
        Yv, Rv = Yn.detach().cpu().view(n, 5, F), ref.detach().view(n, 5, F)
        bitexact(Yv[:, 2:4], Rv[:, 2:4], f"min/max F={F} p={p} sorted={sort_rows}")
        close(Yv, ref64.detach().view(n, 5, F), what="sum/mean/std")
        assert torch.equal(amin.cpu().long(), rargs["min"].view(n, F)), "argmin"
        assert torch.equal(amax.cpu().long(), rargs["max"].view(n, F)), "argmax"
        gy = torch.randn(n, 5 * F, generator=g)
        gyg = gy.cuda()
        if sort_rows:
            gyg = gyg.index_select(0, graph.row_map.long())
        a = torch.autograd.grad(Y, [Pg, Qg], gyg.view_as(Y))
        b = torch.autograd.grad(ref64, [Pr, Qr], gy.double())
        b32 = torch.autograd.grad(ref, [P32, Q32], gy)
        for got, exact, r32, what in zip(a, b, b32, ("dP", "dQ")):
            scale = exact.abs().max().item()
            ref_err = (r32.double() - exact).abs().max().item()
            err = (got.cpu().double() - exact).abs().max().item()
            assert err <= max(REL * scale, 2 * ref_err), f"{what}: err {err:.3e}, reference fp32 err {ref_err:.3e}, scale {scale:.3e}"


def test_gather_rows_colsum():
    from mma_b200 import tc_gemm as tg
    g = torch.Generator().manual_seed(0)
    for n, F in ((1000, 128), (77, 4), (5000, 200)):
        src = torch.randn(n, F, generator=g).cuda()
        idx = torch.randperm(n, generator=g).to(torch.int32).cuda()
        out, cs = tg.gather_rows_colsum(src, idx)
        assert torch.equal(out, src.index_select(0, idx.long()))
        ref = src.double().sum(0)
        assert (cs.double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-4
        out2, cs2 = tg.gather_rows_colsum(src, idx)
        assert torch.equal(cs, cs2), "fixed-order reduction must be bit-reproducible"


def test_cuda_graph_replay_draws_fresh_dropout():
    """The fused layer is capture-safe; with device_seed every replay advances the seed on the device: replay k
    must equal the k-th eager call started from the same seed state (forward and gradients)."""
    from mma_b200 import MMAConv, Graph
    from oracle import restate
    torch.manual_seed(0)
    n, E, Fd = 3000, 40000, 64
    src, dst = rand_graph(n, E, 21)
    hist = torch.bincount(torch.bincount(dst, minlength=n))
    conv = MMAConv(Fd, Fd, ["mean", "max", "std"], ["identity", "amplification"], hist, towers=1,
                   strict_reference=False).cuda()
    conv.fold_min_rows = 32
    conv.device_seed = True
    graph = Graph(src.cuda(), dst.cuda(), n, sort_rows=True)
    x = torch.randn(n, Fd).cuda().requires_grad_()
    gy = torch.randn(n, Fd).cuda()
    params = list(conv.parameters()) + conv.mask_parameters()

    def step():
        y = conv(x, graph)
        return (y,) + torch.autograd.grad(y, [x] + params, gy)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    start = conv._seed_dev.clone()
    eager = [[t.detach().clone() for t in step()] for _ in range(2)]     # detached: no autograd graph of an eager
    torch.cuda.synchronize()                                              # step may outlive into the capture
    assert not torch.equal(eager[0][0], eager[1][0]), "two calls must draw different masks"
    conv._seed_dev.copy_(start)
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        outs = step()
    conv._seed_dev.copy_(start)          # the capture itself does not run the kernels
    for k in range(2):
        cg.replay()
        torch.cuda.synchronize()
        for a, b in zip(outs, eager[k]):
            assert torch.equal(a, b), f"replay {k} differs from eager call {k}"


def test_integration_md_ctypes_stub_runs():
    """The binding INTEGRATION.md section 2 shows (plain ctypes over the C ABI, pasted into the reference's
    MMAConv.aggregate) is executed as written and reproduces the verbatim reference's output."""
    import os
    import re
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(import ctypes as C, torch\n.*?)```", txt, re.S).group(1)
    ns = {}
    cwd = os.getcwd()
    os.chdir(root)                              # the stub loads "mma_b200/libmma_b200.so" relative to the repo root
    try:
        exec(compile(code, "INTEGRATION.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    gd = load_golden("aggregate_all.pt")
    inputs, index, n = gd["inputs"], gd["index"], gd["n"]
    E, T, F_in = inputs.shape
    me = types.SimpleNamespace(avg_deg=gd["avg_deg"], aggregators=gd["aggregators"], scalers=gd["scalers"])
    out = ns["aggregate"](me, inputs.cuda(), index.cuda(), n)
    torch.cuda.synchronize()
    A, S = len(gd["aggregators"]), len(gd["scalers"])
    ov, rv = out.cpu().view(n, T, S, A, F_in), gd["out"].view(n, T, S, A, F_in)
    for ai, a in enumerate(gd["aggregators"]):
        if a in ("min", "max"):
            bitexact(ov[:, :, :, ai], rv[:, :, :, ai], f"stub:{a}")
        else:
            close(ov[:, :, :, ai], rv[:, :, :, ai], what=f"stub:{a}")

