"""Multi-GPU path (mma_b200/parallel.py).

CPU (gloo, world_size 2): the host-side logic -- destination-range bounds, local edge filtering
with global edge ids, the padded all-gather layout and the reduce-scatter -- checked against a
single-process recomputation.  The aggregation kernels themselves have no CPU path.

GPU (needs >= 2 devices, NCCL): sharded K1 forward/backward == single-GPU result (min/max
bit-exact incl. Philox dropout keyed by global edge ids, the rest to fp32 rounding)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _cpu_worker(rank, world, port, n, E):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mma_b200 import parallel as par
        g = torch.Generator().manual_seed(0)
        src = torch.randint(0, n, (E,), generator=g)
        dst = (torch.rand(E, generator=g) ** 2 * n).long().clamp_(max=n - 1)       # skewed
        for balance in ("nodes", "edges"):
            sg = par.ShardedGraph(src, dst, n, rank, world, balance=balance)
            b = sg.bounds
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
            # every edge lands on exactly one rank, in original order, with its global id
            cnt = torch.tensor([sg.E]); dist.all_reduce(cnt); assert int(cnt) == E
            assert torch.equal(src[sg._gid], src[(dst >= sg.lo) & (dst < sg.hi)])
            assert torch.all(sg._gid[1:] > sg._gid[:-1])
            assert torch.equal(sg._dst_local + sg.lo, dst[sg._gid])
            # padded source addressing: gathering rank-major padded rows reproduces global rows
            F = 3
            X = torch.arange(n * F, dtype=torch.float32).view(n, F)
            allx = par.all_gather_rows(X[sg.lo:sg.hi], sg.max_rows)
            assert torch.equal(allx[sg.src_padded], X[src[sg._gid]])
            # reduce-scatter of partial per-source sums == global per-source sum of my rows
            part = torch.zeros(world * sg.max_rows, F).index_add_(0, sg.src_padded, torch.ones(sg.E, F))
            mine = par.reduce_scatter_rows(part, sg.max_rows, sg.rows)
            ref = torch.zeros(n, F).index_add_(0, src, torch.ones(E, F))[sg.lo:sg.hi]
            assert torch.equal(mine, ref)
            if balance == "edges":
                per = [int(((dst >= b[r]) & (dst < b[r + 1])).sum()) for r in range(world)]
                assert max(per) <= 0.75 * E, per                                    # balanced by edges, not nodes
        # data-parallel gradient all-reduce
        p = torch.nn.Parameter(torch.ones(4)); p.grad = torch.full((4,), float(rank + 1))
        par.allreduce_grads([p]); assert torch.equal(p.grad, torch.full((4,), float(sum(range(1, world + 1)))))
    finally:
        dist.destroy_process_group()


def test_partition_and_collectives_gloo_world2():
    mp.spawn(_cpu_worker, args=(2, _free_port(), 500, 6000), nprocs=2, join=True)


def test_partition_bounds_single_process():
    from mma_b200.parallel import partition_bounds, _slices
    dst = torch.cat([torch.zeros(900, dtype=torch.long), torch.arange(100)])
    assert partition_bounds(dst, 100, 4, "nodes") == [0, 25, 50, 75, 100]
    b = partition_bounds(dst, 100, 4, "edges")
    assert b[0] == 0 and b[-1] == 100 and b[1] == 1           # the hub row alone fills the first shard
    assert partition_bounds(torch.zeros(0, dtype=torch.long), 10, 3, "edges") == [0, 3, 6, 10]
    # no rank is ever left without rows: a hub holding more than 1/world of the edges, or fewer nodes per rank than
    # the ceiling split would deal out -- and the failure for N < world is the same ValueError on every rank
    hub = torch.cat([torch.full((1000,), 7, dtype=torch.long), torch.arange(8)])
    for mode in ("edges", "nodes"):
        for world in (2, 3, 4, 8):
            bb = partition_bounds(hub, 8, world, mode)
            assert bb[0] == 0 and bb[-1] == 8 and all(bb[i + 1] > bb[i] for i in range(world)), (mode, world, bb)
        assert partition_bounds(hub, 5, 4, mode) [-1] == 5
        with pytest.raises(ValueError):
            partition_bounds(hub, 3, 4, mode)
    sl = _slices(128, 4)
    assert [(s.start, s.stop) for s in sl] == [(0, 32), (32, 64), (64, 96), (96, 128)]
    assert [(s.start, s.stop) for s in _slices(75, 4)] == [(0, 75)]
    assert sum(s.stop - s.start for s in _slices(100, 3)) == 100


def test_split_graph_batch_covers_every_graph_once():
    """Scheme (ii), config 2: a batch of small graphs is split by whole graphs -- every node and edge lands on exactly
    one rank, renumbered from 0, in its original relative order."""
    from mma_b200.parallel import split_graph_batch
    from oracle import restate
    ei, batch = restate.zinc_like_batch(13, seed=1)
    n, E = batch.numel(), ei.shape[1]
    x = torch.arange(n).float().unsqueeze(1)
    ea = torch.arange(E)
    for world in (1, 2, 3, 5):
        seen_nodes, seen_edges = [], []
        for r in range(world):
            e_l, b_l, nodes, eids, (x_l,), (ea_l,) = split_graph_batch(ei, batch, r, world, x, edge_tensors=(ea,))
            assert torch.equal(x_l, x[nodes]) and torch.equal(ea_l, eids)
            assert torch.equal(e_l + nodes.start, ei[:, eids]) and torch.all(eids[1:] > eids[:-1])
            if b_l.numel():
                assert int(b_l.min()) == 0 and torch.equal(b_l + int(batch[nodes.start]), batch[nodes])
            seen_nodes.append(torch.arange(nodes.start, nodes.stop)); seen_edges.append(eids)
        assert torch.equal(torch.cat(seen_nodes), torch.arange(n))
        assert torch.equal(torch.sort(torch.cat(seen_edges)).values, torch.arange(E))
    with pytest.raises(ValueError):
        split_graph_batch(ei, batch.flip(0), 0, 2)
    bad = ei.clone(); bad[0, 0] = n - 1                         # an edge between the first and the last graph
    with pytest.raises(ValueError):
        split_graph_batch(bad, batch, 0, 2)


def test_config5_partition_balances_edges_and_rows():
    """config 5's generator (bench.py): power-law in-degrees dealt to random node ids.  Destination ranges balanced by
    in-edge count must then also hold similar row counts -- the all-gathered Q layout is padded to the largest range."""
    import bench
    from mma_b200.parallel import partition_bounds
    N, E, world = 40_000, 800_000, 8
    src, dst = bench.powerlaw_edges(N, E, torch.device("cpu"))
    deg = torch.bincount(dst, minlength=N)
    assert abs(dst.numel() - E) < 0.02 * E and int(deg.max()) == E // 200          # the capped hubs are there
    assert int(src.max()) < N and int(dst.max()) < N
    b = partition_bounds(dst, N, world, "edges")
    csum = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(deg, 0)])
    edges = [int(csum[b[r + 1]] - csum[b[r]]) for r in range(world)]
    rows = [b[r + 1] - b[r] for r in range(world)]
    assert max(edges) <= 1.05 * dst.numel() / world, edges
    assert max(rows) <= 1.35 * N / world, rows                                      # padding of the exchange stays small
    # by node count the shard that drew the largest hubs would carry visibly more edges
    bn = partition_bounds(dst, N, world, "nodes")
    en = [int(csum[bn[r + 1]] - csum[bn[r]]) for r in range(world)]
    assert max(en) > max(edges)


def _gpu_worker(rank, world, port, n, E, Fd, p_drop):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import mma_b200
        from mma_b200 import parallel as par
        from oracle import restate
        g = torch.Generator().manual_seed(1)
        src = torch.randint(0, n, (E,), generator=g)
        dst = torch.randint(0, n - 3, (E,), generator=g)
        P, Q = torch.randn(n, Fd, generator=g), torch.randn(n, Fd, generator=g)
        P[torch.rand(n, Fd, generator=g) < 0.3] = 0; Q[torch.rand(n, Fd, generator=g) < 0.3] = 0
        gy = torch.randn(n, 1, 4 * 5 * Fd, generator=g)
        aggr, scal = ["mean", "sum", "min", "max", "std"], ["identity", "amplification", "attenuation", "linear"]
        avg = restate.avg_deg_from_hist(torch.bincount(torch.bincount(dst, minlength=n)))
        # single-GPU reference on this rank's device
        full = mma_b200.Graph(src.to(dev), dst.to(dev), n)
        Pf, Qf = P.to(dev).requires_grad_(), Q.to(dev).requires_grad_()
        Yf = mma_b200.mmconv_aggregate(Pf, Qf, None, full, towers=1, F_in=Fd, aggregators=aggr, scalers=scal,
                                       avg_deg=avg, p_drop=p_drop, seed=77)
        gPf, gQf = torch.autograd.grad(Yf, [Pf, Qf], gy.to(dev))
        for balance, slices in (("nodes", 1), ("edges", 4)):
            sg = par.ShardedGraph(src.to(dev), dst.to(dev), n, rank, world, balance=balance)
            Pl = P[sg.lo:sg.hi].to(dev).requires_grad_(); Ql = Q[sg.lo:sg.hi].to(dev).requires_grad_()
            Yl = par.sharded_mmconv_aggregate(Pl, Ql, sg, F_in=Fd, aggregators=aggr, scalers=scal, avg_deg=avg,
                                              p_drop=p_drop, seed=77, n_slices=slices, max_deg=full.max_deg)
            gPl, gQl = torch.autograd.grad(Yl, [Pl, Ql], gy[sg.lo:sg.hi].to(dev))
            ref = Yf[sg.lo:sg.hi]
            Yv, Rv = Yl.view(sg.rows, 4, 5, Fd), ref.view(sg.rows, 4, 5, Fd)
            assert torch.equal(Yv[:, :, 2:4], Rv[:, :, 2:4]), "min/max must be bit-identical across shardings"
            assert torch.equal(Yv, Rv), "same kernel, same per-row order: sharded forward is bit-identical"
            assert torch.equal(gPl, gPf[sg.lo:sg.hi])
            scale = gQf.abs().max().item()
            assert (gQl - gQf[sg.lo:sg.hi]).abs().max().item() <= 1e-5 * scale      # reduce-scatter sum order
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("p_drop", [0.0, 0.5])
def test_sharded_aggregate_matches_single_gpu(p_drop):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_gpu_worker, args=(2, _free_port(), 4000, 60000, 128, p_drop), nprocs=2, join=True)


@pytest.mark.gpu
def test_sharded_aggregate_world1_runs_the_windowed_exchange_path():
    """One GPU is enough to drive the sharded autograd node end to end (NCCL communicator of one rank): the padded
    all-gathered layout, the column-window pipeline (4 windows: virtual Q base pointers, per-window partial dQ and
    reduce-scatter, the std backward on a re-gathered window) against the plain single-GPU call, bit for bit."""
    mp.spawn(_gpu_worker, args=(1, _free_port(), 3000, 45000, 128, 0.5), nprocs=1, join=True)


def _gpu_layer_worker(rank, world, port, n, E, Fd):
    """The whole drop-in layer on a destination-range shard (fused tcgen05 path) vs the single-GPU layer:
    same seed -> same dropout stream (keyed by global node id), outputs and all gradients must agree."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import mma_b200
        from mma_b200 import parallel as par
        g = torch.Generator().manual_seed(3)
        src = torch.randint(0, n, (E,), generator=g)
        dst = torch.randint(0, n - 3, (E,), generator=g)
        hist = torch.bincount(torch.bincount(dst, minlength=n))
        aggr, scal = ["mean", "sum", "min", "max", "std"], ["identity", "amplification", "attenuation", "linear"]
        torch.manual_seed(5)
        conv = mma_b200.MMAConv(Fd, Fd, aggr, scal, hist, towers=1, strict_reference=False).to(dev)
        conv.fold_min_rows = 64
        x = torch.randn(n, Fd, generator=g)
        gy = torch.randn(n, Fd, generator=g)
        params = list(conv.parameters()) + conv.mask_parameters()
        # single-GPU reference
        full = mma_b200.Graph(src.to(dev), dst.to(dev), n, sort_rows=True)
        xf = x.to(dev).requires_grad_()
        torch.manual_seed(11); conv._calls = 0
        yf = conv(xf, full)
        gf = torch.autograd.grad(yf, [xf] + params, gy.to(dev))
        from mma_b200 import fused_layer
        # ("edges", "1"): the backward's mask GEMMs cut into their dQ-independent part (before the wait for the
        # reduce-scatter) and their dQ part (after it) -- what bench.py's sharded step runs from 4 ranks up
        for balance, split in (("nodes", "auto"), ("edges", "1")):
            fused_layer.SPLIT_MASK_BWD = split
            sg = par.ShardedGraph(src.to(dev), dst.to(dev), n, rank, world, balance=balance)
            xl = x[sg.lo:sg.hi].to(dev).requires_grad_()
            torch.manual_seed(11); conv._calls = 0
            yl = conv(xl, sg)
            gl = torch.autograd.grad(yl, [xl] + params, gy[sg.lo:sg.hi].to(dev))
            tol = lambda a: 2e-5 * a.abs().max().item() + 1e-30
            assert (yl - yf[sg.lo:sg.hi]).abs().max().item() <= tol(yf), "sharded layer output"
            assert (gl[0] - gf[0][sg.lo:sg.hi]).abs().max().item() <= tol(gf[0]), "sharded dx"
            for a, b in zip(gl[1:], gf[1:]):                 # weight gradients: partial per rank -> all-reduce
                a = a.clone()
                dist.all_reduce(a)
                assert (a - b).abs().max().item() <= tol(b), "sharded weight gradient"
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_fused_layer_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_gpu_layer_worker, args=(2, _free_port(), 6000, 90000, 128), nprocs=2, join=True)


@pytest.mark.gpu
def test_sharded_fused_layer_world1():
    """The sharded branch of the fused layer (exchange of Q, partial dQ over the padded source layout, reduce-scatter,
    all-reduced weight gradients) on a one-rank communicator against the single-GPU branch."""
    mp.spawn(_gpu_layer_worker, args=(1, _free_port(), 5000, 70000, 128), nprocs=1, join=True)


def _gpu_dp_worker(rank, world, port):
    """Scheme (ii), config 2: a ZINC-shaped batch split by whole graphs, MMAConv (towers 5, edge features) per rank on
    its graphs, weight gradients all-reduced -- against the full batch on one GPU with the same injected masks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import mma_b200
        from mma_b200 import parallel as par
        from oracle import restate
        ei, batch = restate.zinc_like_batch(24, seed=2)
        n, E = batch.numel(), ei.shape[1]
        g = torch.Generator().manual_seed(4)
        x = torch.randn(n, 75, generator=g)
        ea = torch.randn(E, 50, generator=g)
        gy = torch.randn(n, 75, generator=g)
        keep = (torch.rand(E, 5, 75, generator=g) < 0.5).float() * 2
        torch.manual_seed(6)
        conv = mma_b200.MMAConv(75, 75, ["min", "max"], ["identity", "amplification", "linear"],
                                restate.degree_histogram(ei, n), edge_dim=50, towers=5).to(dev)
        params = list(conv.parameters()) + conv.mask_parameters()
        # the whole batch on this GPU
        xf = x.to(dev).requires_grad_()
        conv._inject_keep = keep.to(dev)
        yf = conv(xf, ei.to(dev), ea.to(dev))
        gf = torch.autograd.grad(yf, [xf] + params, gy.to(dev), allow_unused=True)
        # this rank's graphs
        e_l, b_l, nodes, eids, (x_l, gy_l), (ea_l, keep_l) = par.split_graph_batch(ei, batch, rank, world, x, gy,
                                                                                    edge_tensors=(ea, keep))
        xl = x_l.to(dev).requires_grad_()
        conv._inject_keep = keep_l.to(dev)
        yl = conv(xl, e_l.to(dev), ea_l.to(dev))
        gl = torch.autograd.grad(yl, [xl] + params, gy_l.to(dev), allow_unused=True)
        tol = lambda a: 1e-5 * a.abs().max().item() + 1e-30
        assert (yl - yf[nodes]).abs().max().item() <= tol(yf), "per-rank output rows"
        assert (gl[0] - gf[0][nodes]).abs().max().item() <= tol(gf[0]), "per-rank dx"
        for p_, a in zip(params, gl[1:]):
            p_.grad = None if a is None else a.clone()
        par.allreduce_grads(params)
        for p_, b in zip(params, gf[1:]):
            if b is not None:
                assert (p_.grad - b).abs().max().item() <= tol(b), "all-reduced weight gradient"
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_graph_batch_data_parallel_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_gpu_dp_worker, args=(2, _free_port()), nprocs=2, join=True)

