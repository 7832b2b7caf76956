"""CPU tests of the host-side work partition of the persistent K1 kernels (mma_b200/graph.py): the virtual rows that
cut long rows of a skewed graph (config 5) into segments, and the cost-balanced row chunks.  The Graph object itself
needs a GPU (its CSR is built by the library); here only its pure-torch planning methods run, on a hand-made CSR row
pointer -- the kernels that consume these tables are covered by tests/test_mmconv_gpu.py::
test_stream_kernels_long_rows_and_hubs and tests/test_fullsize_gpu.py."""
import types

import pytest
import torch

from mma_b200.graph import Graph


def fake_graph(deg, seg_len, monkeypatch, sms=148):
    deg = torch.as_tensor(deg, dtype=torch.int64)
    g = object.__new__(Graph)
    g.device = torch.device("cpu")
    g.n_dst = int(deg.numel())
    g.rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(deg, 0)]).to(torch.int32)
    g.E = int(deg.sum())
    g._max_deg = None
    g.K1_MAX_SEG = seg_len
    monkeypatch.setattr(torch.cuda, "get_device_properties",
                        lambda dev: types.SimpleNamespace(multi_processor_count=sms))
    return g, deg


def powerlaw_degrees(n, E, seed, alpha=2.1):
    r = torch.arange(1, n + 1, dtype=torch.float64)
    w = r.pow(-1.0 / (alpha - 1.0))
    deg = (w / w.sum() * E).round().long()
    return deg[torch.randperm(n, generator=torch.Generator().manual_seed(seed))]


@pytest.mark.parametrize("seg_len", [32, 64, 4096])
def test_virtual_rows_cut_long_rows_in_order(seg_len, monkeypatch):
    deg = powerlaw_degrees(5000, 200_000, seed=1)
    deg[17] = 0                                                  # an empty row stays one (empty) virtual row
    g, deg = fake_graph(deg, seg_len, monkeypatch)
    seg = g.k1_segments()
    if int(deg.max()) <= seg_len:
        assert seg is None
        return
    vr, tab, split = seg.vrowptr.long(), seg.seg_tab.long(), seg.split_tab.long()
    rp = g.rowptr.long()
    assert vr[0] == 0 and vr[-1] == g.E and seg.n_vrows == vr.numel() - 1 == tab.shape[0]
    vlen = vr[1:] - vr[:-1]
    assert int(vlen.min()) >= 0 and int(vlen.max()) <= seg_len
    real, pos0, slot = tab[:, 0], tab[:, 1], tab[:, 2]
    assert torch.equal(real, torch.sort(real).values), "virtual rows follow the order of the real rows"
    assert torch.equal(vr[:-1], rp[real] + pos0), "a segment starts pos0 edges into its row"
    assert bool((pos0 % seg_len == 0).all()), "segment starts keep the dropout stream's 32-edge alignment"
    # segments of one row tile it exactly: the lengths per real row add up to its degree
    per_row = torch.zeros(g.n_dst, dtype=torch.int64).index_add_(0, real, vlen)
    assert torch.equal(per_row, deg)
    n_seg = torch.bincount(real, minlength=g.n_dst)
    assert torch.equal(n_seg, ((deg + seg_len - 1) // seg_len).clamp(min=1))
    # partial slots: dense numbering over the segments of split rows only, in order
    is_split = n_seg[real] > 1
    assert bool((slot[~is_split] == -1).all())
    assert torch.equal(slot[is_split], torch.arange(int(is_split.sum())))
    assert seg.n_slots == int(is_split.sum()) and seg.n_split == int((n_seg > 1).sum()) == split.shape[0]
    # merge table: (row, first slot, number of segments)
    rows = split[:, 0]
    assert torch.equal(rows, torch.nonzero(n_seg > 1).flatten())
    assert torch.equal(split[:, 2], n_seg[rows])
    first = torch.cumsum(split[:, 2], 0) - split[:, 2]
    assert torch.equal(split[:, 1], first)


@pytest.mark.parametrize("kind", ["uniform", "powerlaw", "tiny"])
def test_row_chunks_cover_rows_and_balance_cost(kind, monkeypatch):
    if kind == "uniform":
        deg = torch.poisson(torch.full((200_000,), 16.0), generator=torch.Generator().manual_seed(0)).long()
    elif kind == "powerlaw":
        deg = powerlaw_degrees(200_000, 4_000_000, seed=2)
    else:
        deg = torch.tensor([3, 0, 5, 1, 0, 0, 2, 9, 4])
    g, deg = fake_graph(deg, 4096, monkeypatch)
    seg = g.k1_segments()
    ch = g.k1_chunks().long()
    n = g.n_dst if seg is None else seg.n_vrows
    rp = (g.rowptr if seg is None else seg.vrowptr).long()
    assert ch[0] == 0 and ch[-1] == n and bool((ch[1:] >= ch[:-1]).all()), "chunks tile the (virtual) rows in order"
    n_chunks = ch.numel() - 1
    assert n_chunks == max(1, min(148 * 16 * Graph.K1_CHUNKS_PER_WARP, n // 8))
    cost = (rp[ch[1:]] - rp[ch[:-1]]) + Graph.K1_ROW_COST * (ch[1:] - ch[:-1])
    total = g.E + Graph.K1_ROW_COST * n
    assert int(cost.sum()) == total
    biggest_row = int((rp[1:] - rp[:-1]).max()) + Graph.K1_ROW_COST
    # a chunk never exceeds its fair share by more than one row (rows are atomic; long rows were cut by k1_segments)
    assert int(cost.max()) <= total // n_chunks + 1 + biggest_row
    if seg is not None:
        assert biggest_row <= 4096 + Graph.K1_ROW_COST


def test_post_plan_tiles_cover_every_row_once_and_scalers_compound():
    """fused_layer.PostPlan: the degree ranges of a degree-sorted CSR -> tile / slab tables of the grouped GEMMs.
    Every CSR row is handled exactly once (a tile of a big range or the tail path), a tile's weight block is its
    range's, slabs never cross a range or exceed the accumulation-chain bound, and the per-range factors are the
    reference's CUMULATIVE scalers (mma_conv.py:181-195, Q4) evaluated at the range's degree."""
    from mma_b200 import tc_gemm as tg
    from mma_b200.fused_layer import PostPlan
    deg = torch.sort(powerlaw_degrees(30_000, 480_000, seed=3), descending=True).values
    vals, counts = torch.unique_consecutive(deg, return_counts=True)
    hi = torch.cumsum(counts, 0)
    g = types.SimpleNamespace(device=torch.device("cpu"),
                              buckets=[(int(d), int(h - c), int(h)) for d, c, h in zip(vals, counts, hi)],
                              row_map=torch.randperm(deg.numel(), generator=torch.Generator().manual_seed(4)).int())
    scalers = ["identity", "amplification", "attenuation", "linear"]
    avg = {"lin": float(deg.float().mean()), "log": float((deg.float() + 1).log().mean())}
    Fo, F, min_rows = 128, 128, 64
    akinds = (1, 0, 2, 3, 5)                                   # mean, sum, min, max, std
    plan = PostPlan(g, scalers, avg, min_rows, Fo, akinds, F)
    K = plan.K
    # the mean is not materialised: it rides on the sum block with coefficient 1 / deg
    assert plan.mat == (0, 2, 3, 5) and plan.block_of == (0, 0, 1, 2, 3) and plan.inv == (True, False, False, False, False)
    assert K == 4 * F
    n = deg.numel()
    seen = torch.zeros(n, dtype=torch.int64)
    big = {b: i for i, b in enumerate(plan.big)}
    for (r, h, woff, _), (rt, ht, wofft, _) in zip(plan.tile_tab.tolist(), plan.tile_tab_t.tolist()):
        b = next(bb for bb, (_, lo, hh) in enumerate(g.buckets) if lo <= r < hh)
        assert h == g.buckets[b][2] and (r - g.buckets[b][1]) % tg.BM == 0
        assert woff == big[b] * Fo and (rt, ht, wofft) == (r, h, big[b] * K)
        seen[r:min(h, r + tg.BM)] += 1
    if plan.tail_idx is not None:
        seen[plan.tail_idx] += 1
        assert torch.equal(plan.tail_nodes, g.row_map.long()[plan.tail_idx])
        assert all(g.buckets[b][2] - g.buckets[b][1] < min_rows for b in set(plan.tail_bucket.tolist()))
    assert bool((seen == 1).all()), "every degree-sorted row goes through exactly one tile or the tail path"
    slab_rows = torch.zeros(n, dtype=torch.int64)
    for i, (r0, r1, idx, _) in enumerate(plan.slabs.tolist()):
        assert idx == i and 0 < r1 - r0 <= tg.MAX_SLAB_ROWS
        assert len({int(deg[r0]), int(deg[r1 - 1])}) == 1, "a slab lies inside one degree range"
        slab_rows[r0:r1] += 1
    big_rows = torch.zeros(n, dtype=torch.bool)
    for b in plan.big:
        big_rows[g.buckets[b][1]:g.buckets[b][2]] = True
    assert torch.equal(slab_rows, big_rows.long())
    assert plan.seg_ptr.tolist()[0] == 0 and plan.seg_ptr.tolist()[-1] == plan.slabs.shape[0]
    # cumulative factors at each range's degree, the reference's expressions
    d = torch.tensor([max(b[0], 1) for b in g.buckets], dtype=torch.float32)
    amp = torch.log(d + 1) / avg["log"]
    want = torch.stack([torch.ones_like(d), amp, amp * (avg["log"] / torch.log(d + 1)),
                        amp * (avg["log"] / torch.log(d + 1)) * (d / avg["lin"])])
    assert torch.equal(plan.cum, want)
    # coef[b, s, a, m] = cum_s(d_b) * [a -> m] * (1 / d_b for the mean): contracting it with the reference's S x A blocks
    # Y[s, a] = cum_s * agg_a reproduces the literal formula from the materialised blocks (sum, min, max, std)
    B = len(plan.big)
    assert plan.coef_big.shape == (B, 4, 5, 4)
    gen = torch.Generator().manual_seed(0)
    zsum, zmin, zmax, zstd = (torch.randn(B, generator=gen) for _ in range(4))
    dbig = d[plan.big]
    lit = torch.stack([zsum / dbig, zsum, zmin, zmax, zstd], 1).unsqueeze(1) * plan.cum[:, plan.big].t().unsqueeze(2)   # [B,S,A]
    got = torch.einsum("bsam,bm->bsa", plan.coef_big, torch.stack([zsum, zmin, zmax, zstd], 1))
    assert torch.allclose(got, lit, rtol=1e-6, atol=0)
    # min_rows = None (the layer's default): at most MAX_BIG degree ranges get their own weight -- all of them on a graph
    # with few distinct degrees (no tail path at all), the largest ones on a skewed one
    auto = PostPlan(g, scalers, avg, None, Fo, akinds, F)
    assert len(auto.big) <= PostPlan.MAX_BIG and auto.min_rows >= 1
    if len(g.buckets) <= PostPlan.MAX_BIG:
        assert auto.tail_idx is None and len(auto.big) == len(g.buckets)
    few = types.SimpleNamespace(device=torch.device("cpu"), buckets=[(5, 0, 3), (4, 3, 1000), (0, 1000, 1001)],
                                row_map=torch.arange(1001).int())
    pf = PostPlan(few, scalers, avg, None, Fo, akinds, F)
    assert pf.tail_idx is None and pf.big == [0, 1, 2] and pf.min_rows == 1
    from mma_b200.fused_layer import fold_blocks, materialised_blocks
    assert fold_blocks((1, 2)) == ((0, 2), (0, 1), (True, False))          # a mean without a sum: a sum block is written
    assert materialised_blocks(["mean", "sum", "min", "max", "std"]) == 4 and materialised_blocks(["min", "max"]) == 2
