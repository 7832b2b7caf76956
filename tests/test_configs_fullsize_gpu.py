"""BASELINE configs 1-3 at FULL size against the oracle port (oracle/restate.py), which is fast enough on the host to
run them whole: the real Cora and Pubmed topologies (tests/golden/planetoid_topology.pt) and a 128-graph ZINC-shaped
batch.  Same weights, same injected dropout keep masks on both sides; output and every gradient within 1e-5 of the
oracle, relative to the tensor's largest magnitude (north_star).  Also: the column-window ABI of K1 (the sharded
pipeline's building block) bit for bit against the full-width call on one GPU, and the graph cache against address
reuse."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
REL = 1e-5

ORDER = ["moment_3", "sum", "sum2", "sum3", "sum4", "mean", "mean2", "mean3", "mean4", "max", "max2", "max3",
         "max4", "min", "min2", "min3", "min4", "softmax", "softmin", "std", "normalized_mean"]


def close(a, b, rel=REL, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= rel * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


@pytest.mark.parametrize("topology,Fd,C,names,p", [
    ("pubmed", 16, 3, ["min", "min2", "min3", "min4"], 0.5),       # config 3 (node_classification README.md:58)
    ("cora", 64, 7, ["mean", "mean2"], 0.75),                      # config 1's layer (README.md:70)
])
def test_nc_layer_full_topology_vs_oracle(topology, Fd, C, names, p):
    """reference: node_classification/layers.py:540-651 (min*), :305-364 (mean*), :853-867 (forward); `min` runs on the
    RAW mask under the default activation (Q8)."""
    from mma_b200.node_classification.layers import MMA
    from oracle import restate
    topo = load_golden("planetoid_topology.pt")[topology]
    rowptr, col = topo["rowptr"], topo["col"]
    n, E = rowptr.numel() - 1, col.numel()
    g = torch.Generator().manual_seed(42)
    x = torch.randn(n, Fd, generator=g)
    masks = {nm: torch.randn(2 * Fd, Fd, generator=g) * 0.1 for nm in ORDER}
    W, b = torch.randn(Fd, C, generator=g) * 0.1, torch.randn(C, generator=g) * 0.1
    gy = torch.randn(n, C, generator=g)
    keeps = {nm: (torch.rand(E, Fd, generator=g) >= p).float() / (1.0 - p) for nm in names}
    dev = "cuda"
    add_all = [col[rowptr[i]:rowptr[i + 1]].numpy() for i in range(n)]
    ps = {nm: torch.nn.Parameter(torch.empty(2 * Fd, Fd, device=dev)) for nm in ORDER}
    Wd, bd = torch.nn.Parameter(torch.empty(Fd, C, device=dev)), torch.nn.Parameter(torch.empty(C, device=dev))
    L = MMA(add_all, "new_sigmoid", 2, Fd, C, Wd, bd, *[ps[nm] for nm in ORDER], p, names, dev)
    with torch.no_grad():
        Wd.copy_(W); bd.copy_(b)
        for nm in ORDER:
            ps[nm].copy_(masks[nm])
    L._inject_keep = {nm: keeps[nm].to(dev) for nm in names}
    adj = restate.csr_to_sparse_adj(rowptr, col, n)
    xd = x.to(dev).requires_grad_()
    y = L(xd, adj.to(dev))
    grads = torch.autograd.grad(y, [xd, Wd, bd] + [ps[nm] for nm in names], gy.to(dev))
    xr, Wr, br = x.clone().requires_grad_(), W.clone().requires_grad_(), b.clone().requires_grad_()
    mr = {nm: masks[nm].clone().requires_grad_() for nm in names}
    yr = restate.nc_forward(xr, adj, rowptr, col, mr, Wr, br, names, "new_sigmoid", p, keeps)
    gr = torch.autograd.grad(yr, [xr, Wr, br] + [mr[nm] for nm in names], gy)
    close(y, yr, what=f"{topology}: y")
    for a, r, nm in zip(grads, gr, ["dx", "dW", "db"] + [f"dmask_{nm}" for nm in names]):
        close(a, r, what=f"{topology}: {nm}")


def test_config2_zinc_batch_128_graphs_vs_oracle():
    """reference: graph_regression/mma.py:92-95 (layer hyper-parameters), mma_conv.py:121-196; towers 5, F_in 75,
    edge features through the encoder, min,max x identity,amplification,linear."""
    import mma_b200
    from mma_b200.synthetic import zinc_like_batch, degree_histogram
    from oracle import restate
    ei, batch = zinc_like_batch(128, seed=42)
    n, E = int(batch.numel()), int(ei.shape[1])
    assert 2500 < n < 3500 and 5500 < E < 7500
    g = torch.Generator().manual_seed(42)
    torch.manual_seed(42)
    conv = mma_b200.MMAConv(75, 75, ["min", "max"], ["identity", "amplification", "linear"], degree_histogram(ei, n),
                            edge_dim=50, towers=5).cuda()
    x, ea, gy = torch.randn(n, 75, generator=g), torch.randn(E, 50, generator=g), torch.randn(n, 75, generator=g)
    keep = (torch.rand(E, 5, 75, generator=g) < 0.5).float() * 2
    conv._inject_keep = keep.cuda()
    xd, ead = x.cuda().requires_grad_(), ea.cuda().requires_grad_()
    params = list(conv.parameters()) + conv.mask_parameters()
    y = conv(xd, ei.cuda(), ead)
    grads = torch.autograd.grad(y, [xd, ead] + params, gy.cuda())
    w = restate.weights_from_module(conv)
    for t in w.tensors():
        t.requires_grad_()
    xr, ear = x.clone().requires_grad_(), ea.clone().requires_grad_()
    yr = restate.mmaconv_forward(w, xr, ei, ear, keep)
    gr = torch.autograd.grad(yr, [xr, ear], gy)
    close(y, yr, what="c2 y")
    close(grads[0], gr[0], what="c2 dx")
    close(grads[1], gr[1], what="c2 d edge_attr")
    # min / max selections: the layer's aggregate() on the oracle's own messages is bit-exact incl. the scaler blocks
    with torch.no_grad():
        xt = x.view(n, 1, 75).repeat(1, 5, 1)
        msg = restate.mmaconv_message(w, xt[ei[1]], xt[ei[0]], ea, keep)
        want = restate.mmaconv_aggregate(msg, ei[1], n, w.aggregators, w.scalers, w.avg_deg)
        got = conv.aggregate(msg.cuda(), ei[1].cuda(), n)
    assert torch.equal(got.cpu().view(torch.int32), want.view(torch.int32)), "min/max x scalers must be bit-exact"


@pytest.mark.parametrize("F,n_slices,p", [(128, 4, 0.5), (128, 2, 0.0), (64, 2, 0.5)])
def test_k1_column_windows_bit_equal_full_width(F, n_slices, p):
    """The column-window ABI (col0 / ncols, the virtual Q base `data_ptr - 4*col0` into a narrow copy with ldq = w, the
    per-window G / dP offsets, the std backward with Q given only as a window) == one full-width launch, bit for bit --
    forward outputs, arg indices, saved statistics, per-edge gradient rows and dP.  (mma_b200/parallel.py:_ShardedAggregate
    is exactly this loop plus the collectives.)"""
    import mma_b200
    from mma_b200 import functional as MF, _lib
    from mma_b200.parallel import _slices
    dev = torch.device("cuda")
    n, E = 3000, 45000
    g = torch.Generator().manual_seed(F + n_slices)
    src, dst = torch.randint(0, n, (E,), generator=g), torch.randint(0, n - 5, (E,), generator=g)
    graph = mma_b200.Graph(src.to(dev), dst.to(dev), n, sort_rows=True)
    P, Q = torch.randn(n, F, generator=g).to(dev), torch.randn(n, F, generator=g).to(dev)
    P[torch.rand(n, F, device=dev) < 0.3] = 0
    akinds = tuple(_lib.AGGR_KINDS[a] for a in ["mean", "sum", "min", "max", "std"])
    skinds = (0,)
    A = len(akinds)

    def buffers():
        mk = lambda dt, w: torch.full((n, w), -7, dtype=dt, device=dev)
        return (mk(torch.float32, A * F), mk(torch.int32, F), mk(torch.int32, F), mk(torch.float32, F),
                mk(torch.float32, F))
    Y0, amin0, amax0, mean0, var0 = buffers()
    MF.k1_forward(graph, P, Q, None, None, T=1, F_in=F, akinds=akinds, skinds=skinds, tab=None, p_drop=p, seed=5,
                  Y=Y0, arg_min=amin0, arg_max=amax0, mean=mean0, var=var0, local_args=True)
    Y1, amin1, amax1, mean1, var1 = buffers()
    sl = _slices(F, n_slices)
    assert len(sl) == n_slices
    windows = []
    for s_ in sl:
        w = s_.stop - s_.start
        Qk = Q[:, s_].contiguous()
        windows.append(Qk)
        MF.k1_forward(graph, P, None, None, None, T=1, F_in=F, akinds=akinds, skinds=skinds, tab=None, p_drop=p,
                      seed=5, Y=Y1, arg_min=amin1, arg_max=amax1, mean=mean1, var=var1, local_args=True,
                      col0=s_.start, ncols=w, q_ptr=Qk.data_ptr() - 4 * s_.start, ldq=w)
    for a, b, what in ((Y0, Y1, "Y"), (amin0, amin1, "argmin"), (amax0, amax1, "argmax"), (mean0, mean1, "mean"),
                       (var0, var1, "var")):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), f"forward {what}: windows != full width"
    # backward: destination pass per window, G and dP at full width in place
    dY = torch.randn(n, A * F, generator=g).to(dev)
    graph.build_transpose()
    G0, dP0 = torch.full((E, F), -7.0, device=dev), torch.full((n, F), -7.0, device=dev)
    MF.k1_backward_dst(graph, P, Q, None, None, T=1, F_in=F, akinds=akinds, skinds=skinds, tab=None, p_drop=p, seed=5,
                       dY=dY, arg_min=amin0, arg_max=amax0, mean=mean0, var=var0, gslot=graph.csr2csc, G=G0, ldg=F,
                       dP=dP0, lddp=F, local_args=True)
    G1, dP1 = torch.full((E, F), -7.0, device=dev), torch.full((n, F), -7.0, device=dev)
    for s_, Qk in zip(sl, windows):
        w = s_.stop - s_.start
        MF.k1_backward_dst(graph, P, None, None, None, T=1, F_in=F, akinds=akinds, skinds=skinds, tab=None, p_drop=p,
                           seed=5, dY=dY, arg_min=amin0, arg_max=amax0, mean=mean0, var=var0, gslot=graph.csr2csc,
                           G=G1, ldg=F, dP=dP1, lddp=F, local_args=True, col0=s_.start, ncols=w,
                           q_ptr=Qk.data_ptr() - 4 * s_.start, ldq=w)
    assert torch.equal(G0.view(torch.int32), G1.view(torch.int32)), "per-edge gradient rows: windows != full width"
    assert torch.equal(dP0.view(torch.int32), dP1.view(torch.int32)), "dP: windows != full width"
    # transpose pass per window == full width
    dQ0 = torch.empty(n, F, device=dev)
    _lib.check(_lib.lib().mma_segment_sum_rows(_lib.ptr(graph.colptr), None, None, n, _lib.ptr(G0), F, F,
                                               _lib.ptr(dQ0), F, _lib.stream_ptr(dev)), "segment_sum_rows")
    dQ1 = torch.empty(n, F, device=dev)
    for s_ in sl:
        w = s_.stop - s_.start
        part = torch.empty(n, w, device=dev)
        _lib.check(_lib.lib().mma_segment_sum_rows(_lib.ptr(graph.colptr), None, None, n, G0.data_ptr() + 4 * s_.start,
                                                   F, w, _lib.ptr(part), w, _lib.stream_ptr(dev)), "segment_sum_rows")
        dQ1[:, s_] = part
    assert torch.equal(dQ0.view(torch.int32), dQ1.view(torch.int32)), "dQ: windows != full width"


def test_graph_cache_survives_address_reuse():
    """a freed edge_index of the same shape hands its address to the next one; the cache must not
    return the old topology.  A per-step `torch.randint` loop alternates between two addresses."""
    import mma_b200
    from mma_b200.graph import cached_graph, clear_cache
    clear_cache()
    n, E = 200, 1500
    seen_ptrs = set()
    for step in range(6):
        g = torch.Generator(device="cuda").manual_seed(step)
        ei = torch.randint(0, n, (2, E), device="cuda", generator=g)
        seen_ptrs.add(ei.data_ptr())
        gr = cached_graph(ei, n)
        want = torch.bincount(ei[1], minlength=n)
        got = (gr.rowptr[1:] - gr.rowptr[:-1]).long()
        assert torch.equal(got, want), f"step {step}: stale graph served from the cache"
        assert cached_graph(ei, n) is gr                      # same tensor again: a hit
        assert cached_graph(ei[:, :], n) is gr                # a view of the same storage: a hit
        ei[1, 0] = (ei[1, 0] + 1) % n                         # in-place edit bumps the version: rebuilt
        gr2 = cached_graph(ei, n)
        assert gr2 is not gr
        assert torch.equal((gr2.rowptr[1:] - gr2.rowptr[:-1]).long(), torch.bincount(ei[1], minlength=n))
        del ei, gr, gr2
    # adjacency cache: same rule
    from mma_b200.graph import cached_adj
    for step in range(4):
        idx = torch.randint(0, n, (2, 900), device="cuda")
        adj = torch.sparse_coo_tensor(idx, torch.ones(900, device="cuda"), (n, n))
        s = cached_adj(adj)
        dense = adj.to_dense()
        assert torch.equal((s.rowptr[1:] - s.rowptr[:-1]).long(), (dense != 0).sum(1))
        del idx, adj, s, dense


def test_kernels_stay_inside_their_output_buffers():
    """compute-sanitizer is closed on this GPU pool (profiles/r2l_compute_sanitizer_closed.log), so the out-of-bounds
    check is done by hand: every output of the hot kernels is a window inside a larger allocation whose guard bands
    (before, after, and the padding columns of every row) hold a bit pattern that must survive the launch.  Odd sizes on
    purpose: rows not a multiple of the tile / chunk sizes, a ragged last tile, empty rows, a hub row."""
    import mma_b200
    from mma_b200 import functional as MF, tc_gemm as tg, _lib
    dev = torch.device("cuda")
    GUARD = 0x7FC0DEAD                      # a NaN payload no kernel produces

    def guarded(rows, cols, dtype, pad_cols=4, pad_rows=3):
        full = torch.full((rows + 2 * pad_rows, cols + pad_cols), GUARD, dtype=torch.int32, device=dev)
        view = full[pad_rows:pad_rows + rows, :cols].view(dtype)
        return full, view, (pad_rows, rows, cols)

    def intact(full, geom, what):
        pad_rows, rows, cols = geom
        assert bool((full[:pad_rows] == GUARD).all()) and bool((full[pad_rows + rows:] == GUARD).all()), f"{what}: rows outside"
        assert bool((full[:, cols:] == GUARD).all()), f"{what}: padding columns"

    n, E, F = 1237, 20011, 128
    g = torch.Generator().manual_seed(9)
    src = torch.randint(0, n, (E,), generator=g)
    dst = torch.randint(0, n - 7, (E,), generator=g)
    dst[:3000] = 5                                                  # a hub row
    graph = mma_b200.Graph(src.to(dev), dst.to(dev), n, sort_rows=True)
    graph.build_transpose()
    P, Q = torch.randn(n, F, generator=g).to(dev), torch.randn(n, F, generator=g).to(dev)
    akinds = (0, 2, 3, 5)
    A = len(akinds)
    Zf, Z, gz = guarded(n, A * F, torch.float32)
    amf, amin, ga = guarded(n, F, torch.int32, pad_cols=0)          # [n, F] contiguous by contract: row guards only
    axf, amax, gx = guarded(n, F, torch.int32, pad_cols=0)
    mf, mean, gm = guarded(n, F, torch.float32, pad_cols=0)
    vf, var, gv = guarded(n, F, torch.float32, pad_cols=0)
    MF.k1_forward(graph, P, Q, None, None, T=1, F_in=F, akinds=akinds, skinds=(0,), tab=None, p_drop=0.5, seed=3, Y=Z,
                  arg_min=amin, arg_max=amax, mean=mean, var=var, local_args=True)
    torch.cuda.synchronize()
    for full, geom, what in ((Zf, gz, "Z"), (amf, ga, "argmin"), (axf, gx, "argmax"), (mf, gm, "mean"), (vf, gv, "var")):
        intact(full, geom, "K1 forward " + what)
    assert bool(torch.isfinite(Z).all())
    dZ = torch.randn(n, A * F, generator=g).to(dev)
    Gf, G, gg = guarded(E, F, torch.float32)
    dPf, dP, gp = guarded(n, F, torch.float32)
    MF.k1_backward_dst(graph, P, Q, None, None, T=1, F_in=F, akinds=akinds, skinds=(0,), tab=None, p_drop=0.5, seed=3,
                       dY=dZ, arg_min=amin, arg_max=amax, mean=mean, var=var, gslot=graph.csr2csc, G=G, ldg=G.stride(0), dP=dP,
                       lddp=dP.stride(0), local_args=True)
    dQf, dQ, gq = guarded(n, F, torch.float32)
    _lib.check(_lib.lib().mma_segment_sum_rows(_lib.ptr(graph.colptr), None, None, n, _lib.ptr(G), G.stride(0), F,
                                               _lib.ptr(dQ), dQ.stride(0), _lib.stream_ptr(dev)), "segment_sum_rows")
    torch.cuda.synchronize()
    intact(Gf, gg, "K1 backward G"); intact(dPf, gp, "K1 backward dP"); intact(dQf, gq, "transpose pass dQ")
    assert bool(torch.isfinite(G).all()) and bool(torch.isfinite(dQ).all())
    # GEMMs: ragged last tile (M % 128 != 0), N not a multiple of 128, row scatter into a guarded output
    for M, N, K in ((1237, 132, 100), (300, 384, 128), (77, 20, 36)):
        Am = torch.randn(M, K, generator=g).to(dev)
        W = torch.randn(N, K, generator=g).to(dev)
        hi, lo = tg.split_weight(W)
        Cf, C, gc = guarded(M, N, torch.float32)
        perm = torch.randperm(M, generator=g).to(dev).int()
        tg.linear(Am, hi, lo, N, out=C, out_map=perm)
        torch.cuda.synchronize()
        intact(Cf, gc, f"linear {M}x{N}x{K}")
        ref = (Am.double() @ W.double().t())
        got = torch.empty_like(ref); got[perm.long()] = ref
        assert float((C.double() - got).abs().max() / got.abs().max()) < 1e-5
        Cf2, C2, gc2 = guarded(M, N, torch.float32)
        tg.linear(Am, hi, lo, N, out=C2)                              # plain output: the TMA tile-store path where it applies
        torch.cuda.synchronize()
        intact(Cf2, gc2, f"linear (plain) {M}x{N}x{K}")
