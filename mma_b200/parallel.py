"""Multi-GPU execution of the aggregation path: one process per GPU, torch.distributed (NCCL
over NVLink 5 / NVSwitch) for the plumbing.  The reference has no distributed code at all
(SURVEY.md 2.1); both schemes below are new and follow BASELINE.json's north_star:

(i)  ONE LARGE GRAPH -> partition by contiguous DESTINATION-node ranges (balanced by in-edge
     count or by node count).  Rank r owns rows [lo_r, hi_r) of the destination CSR, of P, Y,
     arg and the Q rows of its own nodes.  Per layer call there is exactly one exchange step per
     direction:
       forward : all-gather of the Q rows (the halo; for a uniform random graph ~all of Q),
       backward: reduce-scatter of the per-rank partial dQ (sum over ranks, fixed ring order).
     min/max arg indices and the Philox dropout are keyed by GLOBAL edge ids (edge_gid), so the
     sharded result equals the single-GPU result bit-for-bit for min/max and within fp32
     rounding for the rest.  Two carriers for the exchanges: the fused layer (fused_layer.py, what bench.py
     runs) moves them over NVLink peer memory on the copy engines (peer.py); the aggregate-only entry point
     of this module (`sharded_mmconv_aggregate`) uses NCCL collectives pipelined over FEATURE SLICES (columns
     are independent under every aggregator: the collective of slice k+1 overlaps the aggregation kernel of
     slice k on a side stream) -- it is also the gloo-testable statement of the protocol.
(ii) BATCHES OF SMALL GRAPHS (ZINC-like) -> plain data parallel: split the block-diagonal batch
     by graphs (no halo), all-reduce weight gradients only (`allreduce_grads`).

Host-side logic (bounds, edge filtering, collectives) also runs on CPU tensors with the gloo
backend so it can be tested without GPUs; the aggregation itself has no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib
from . import functional as MF
from .graph import Graph


# ------------------------------------------------------------------------------------------
# partitioning (pure index arithmetic: works on CPU and CUDA tensors)
# ------------------------------------------------------------------------------------------
def partition_bounds(dst: Tensor, num_nodes: int, world: int, balance: str = "edges") -> List[int]:
    """world+1 node boundaries of contiguous destination ranges.  balance='nodes': equal node
    counts; 'edges': equal in-edge counts (essential for skewed degree distributions)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if balance == "nodes" or dst.numel() == 0:
        if 0 < num_nodes < world:
            raise ValueError(f"cannot give {world} ranks a non-empty destination range of {num_nodes} nodes")
        return [(num_nodes * r) // world for r in range(world)] + [num_nodes]
    if balance != "edges":
        raise ValueError(f"unknown balance mode {balance!r}")
    deg = torch.bincount(dst, minlength=num_nodes)
    csum = torch.cumsum(deg, 0)
    total = int(csum[-1])
    targets = torch.tensor([(total * r) // world for r in range(1, world)], device=dst.device, dtype=csum.dtype)
    cuts = torch.searchsorted(csum, targets, right=False).tolist() if world > 1 else []
    bounds = [0] + [min(int(c) + 1, num_nodes) for c in cuts] + [num_nodes]
    # every rank gets at least one row (a hub holding >= 1/world of the edges would otherwise leave an EMPTY range,
    # and a rank without rows cannot enter the layer's GEMMs / collectives): the cuts are pushed apart.  The same
    # arithmetic runs on every rank, so either all ranks get bounds or all raise.
    if num_nodes < world:
        raise ValueError(f"cannot give {world} ranks a non-empty destination range of {num_nodes} nodes")
    for i in range(1, world):
        bounds[i] = min(max(bounds[i], bounds[i - 1] + 1), num_nodes - (world - i))
    return bounds


def local_edges(src: Tensor, dst: Tensor, lo: int, hi: int):
    """Edges whose destination lies in [lo, hi), in their original relative order, plus their
    global edge ids."""
    sel = (dst >= lo) & (dst < hi)
    gid = torch.nonzero(sel, as_tuple=False).flatten()
    return src.index_select(0, gid), dst.index_select(0, gid) - lo, gid


class ShardedGraph:
    """This rank's shard of a destination-range partitioned graph."""

    def __init__(self, src: Tensor, dst: Tensor, num_nodes: int, rank: int, world: int,
                 balance: str = "nodes", bounds: Optional[Sequence[int]] = None, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.num_nodes = int(num_nodes)
        self.bounds = list(bounds) if bounds is not None else partition_bounds(dst, num_nodes, world, balance)
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.rows = self.hi - self.lo
        self.max_rows = max(self.bounds[r + 1] - self.bounds[r] for r in range(world))
        self.E_total = int(dst.numel())
        s, d, gid = local_edges(src, dst, self.lo, self.hi)
        self.E = int(gid.numel())
        # sources are addressed in the all-gathered, per-rank padded layout [world * max_rows]
        self.src_padded = self.to_padded(s)
        self.local: Optional[Graph] = None
        if src.is_cuda:
            # local rows in order of decreasing in-degree (row_map addresses P / dP): equal-degree
            # rows are contiguous for the scaler-folded post GEMM, and the 4 or 8 rows that share a
            # warp in a narrow feature window have (almost) equal length
            g = Graph(self.src_padded, d, self.rows, world * self.max_rows, need_transpose=False,
                      sort_rows=True)
            g.gid = gid.to(torch.int32).contiguous()
            g.E_total = self.E_total
            g.rng_row0 = self.lo                      # dropout stream keyed by the GLOBAL destination node id
            self.local = g
        self._dst_local, self._gid = d, gid

    def to_padded(self, node: Tensor) -> Tensor:
        """global node id -> row in the all-gathered [world, max_rows] layout."""
        b = torch.tensor(self.bounds, device=node.device, dtype=node.dtype)
        owner = torch.searchsorted(b, node, right=True) - 1
        owner = owner.clamp_(0, self.world - 1)
        return owner * self.max_rows + (node - b[owner])

    @property
    def max_deg(self) -> int:
        return self.local.max_deg


# ------------------------------------------------------------------------------------------
# collectives (NCCL on GPUs; gloo-compatible so the plumbing is testable on CPU)
# ------------------------------------------------------------------------------------------
def all_gather_rows(x_loc: Tensor, max_rows: int, group=None, out: Optional[Tensor] = None) -> Tensor:
    """[rows_r, F] per rank (rows_r <= max_rows) -> [world * max_rows, F] (rank-major, zero padded)."""
    world = dist.get_world_size(group)
    Fd = x_loc.shape[1]
    if x_loc.shape[0] != max_rows:
        pad = torch.zeros((max_rows, Fd), dtype=x_loc.dtype, device=x_loc.device)
        pad[: x_loc.shape[0]] = x_loc
        x_loc = pad
    x_loc = x_loc.contiguous()
    if out is None:
        out = torch.empty((world * max_rows, Fd), dtype=x_loc.dtype, device=x_loc.device)
    dist.all_gather_into_tensor(out, x_loc, group=group)
    return out


def reduce_scatter_rows(x_all: Tensor, max_rows: int, rows: int, group=None) -> Tensor:
    """[world * max_rows, F] partial sums per rank -> this rank's [rows, F] total."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    Fd = x_all.shape[1]
    x_all = x_all.contiguous()
    if dist.get_backend(group) == "gloo":       # gloo has no reduce_scatter: test-only plumbing path
        tmp = x_all.clone()
        dist.all_reduce(tmp, group=group)
        return tmp[rank * max_rows: rank * max_rows + rows].contiguous()
    out = torch.empty((max_rows, Fd), dtype=x_all.dtype, device=x_all.device)
    dist.reduce_scatter_tensor(out, x_all, group=group)
    return out[:rows]


def allreduce_grads(params, group=None, average: bool = False) -> None:
    """Data-parallel step (ii): SUM weight gradients over ranks in one flat bucket (`average=True` divides by the
    world size afterwards; the layers' losses here are sums over nodes / graphs, so the sum is the full-batch
    gradient).  The bucket layout is rank-independent: every parameter of `params` takes part, a rank whose
    gradient is None (e.g. it received no graph of the batch) contributes zeros and receives the total."""
    params = list(params)
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        g = flat[off: off + p.numel()].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += p.numel()


def split_graph_batch(edge_index: Tensor, batch: Tensor, rank: int, world: int, *node_tensors: Tensor,
                      edge_tensors: Sequence[Tensor] = ()):
    """Scheme (ii) for a PyG-style batch of small graphs (ZINC, graph_regression/mma.py:52-54): the batch is block
    diagonal, so rank r simply takes a contiguous run of whole graphs -- no halo, no exchange in the layer; only the
    weight gradients are summed afterwards (`allreduce_grads`).  `batch` [N] is the sorted graph id of every node.
    Returns (edge_index_local [2, E_r] renumbered from 0, batch_local [N_r] renumbered from 0, node_slice, edge_ids,
    node tensors sliced to the rank's nodes, edge tensors sliced to its edges).  Pure index arithmetic (CPU or CUDA)."""
    if batch.numel() and bool((batch[1:] < batch[:-1]).any()):
        raise ValueError("split_graph_batch expects nodes numbered graph by graph (sorted `batch`)")
    n_graphs = int(batch.max()) + 1 if batch.numel() else 0
    per = (n_graphs + world - 1) // world
    g_lo, g_hi = min(rank * per, n_graphs), min((rank + 1) * per, n_graphs)
    counts = torch.bincount(batch, minlength=n_graphs)
    ptr = torch.zeros(n_graphs + 1, dtype=torch.long, device=batch.device)
    ptr[1:] = torch.cumsum(counts, 0)
    n_lo, n_hi = int(ptr[g_lo]), int(ptr[g_hi])
    dst = edge_index[1]
    eids = torch.nonzero((dst >= n_lo) & (dst < n_hi), as_tuple=False).flatten()      # original relative order
    ei = edge_index.index_select(1, eids) - n_lo
    if ei.numel() and (int(ei.min()) < 0 or int(ei.max()) >= n_hi - n_lo):
        raise ValueError("an edge connects two graphs of the batch: it is not block diagonal")
    nodes = slice(n_lo, n_hi)
    return (ei, batch[nodes] - g_lo, nodes, eids, tuple(t[nodes] for t in node_tensors),
            tuple(t.index_select(0, eids) for t in edge_tensors))


# ------------------------------------------------------------------------------------------
# sharded fused aggregate
# ------------------------------------------------------------------------------------------
def _slices(F_in: int, n_slices: int) -> List[slice]:
    """Column slices of one tower's F_in (multiples of 4 columns so 128-bit access survives)."""
    n_slices = max(1, min(n_slices, F_in // 4 if F_in % 4 == 0 else 1))
    step = -(-F_in // n_slices)
    step = -(-step // 4) * 4 if F_in % 4 == 0 else F_in
    return [slice(lo, min(lo + step, F_in)) for lo in range(0, F_in, step)]


class _ShardedAggregate(torch.autograd.Function):
    """K1 over this rank's destination rows with the source rows all-gathered.  Pipelined over
    column windows (mmconv_aggregate_fwd's col0/ncols): the all-gather (forward) / reduce-scatter
    (backward) of one window runs on a communication stream while the kernel of the neighbouring
    window runs on the compute stream.  Every full-width tensor (P, Y, arg, stats, dY, dP) is
    addressed in place; only the gathered Q window is a narrow [world*max_rows, w] buffer, passed
    with the base pointer shifted back by col0 columns so that global column indexing lands in it."""

    @staticmethod
    def forward(ctx, P, Q, sg: ShardedGraph, F_in: int, akinds, skinds, tab, p_drop, seed, n_slices):
        g = sg.local
        dev = P.device
        P = P if P.stride(-1) == 1 else P.contiguous()
        A, S = len(akinds), len(skinds)
        sl = _slices(F_in, n_slices)
        n = sg.rows
        Y = torch.empty((n, 1, S * A * F_in), dtype=torch.float32, device=dev)
        has_min, has_max = 2 in akinds, 3 in akinds
        need_sq = 4 in akinds or 5 in akinds
        mk = lambda dt: torch.empty((n, F_in), dtype=dt, device=dev)
        arg_min = mk(torch.int32) if has_min else None
        arg_max = mk(torch.int32) if has_max else None
        mean = mk(torch.float32) if need_sq else None
        var = mk(torch.float32) if need_sq else None
        comm = _comm_stream(dev)
        cur = torch.cuda.current_stream(dev)
        comm.wait_stream(cur)
        gathered, events = [], []
        for s_ in sl:                                   # enqueue every window's all-gather up front
            with torch.cuda.stream(comm):
                qa = all_gather_rows(Q[:, s_].contiguous(), sg.max_rows, sg.group)
                ev = torch.cuda.Event(); ev.record(comm)
            gathered.append(qa); events.append(ev)
        ak, sk = _lib.i32_array(akinds), _lib.i32_array(skinds)
        for k, s_ in enumerate(sl):
            cur.wait_event(events[k])
            w = s_.stop - s_.start
            Qk = gathered[k]
            q_base = Qk.data_ptr() - 4 * s_.start       # virtual base: column c of the window's row j
            MF.k1_forward(g, P, None, None, None, T=1, F_in=F_in, akinds=akinds, skinds=skinds, tab=tab,
                          p_drop=p_drop, seed=seed, Y=Y, arg_min=arg_min, arg_max=arg_max, mean=mean, var=var,
                          col0=s_.start, ncols=w, q_ptr=q_base, ldq=w)
            Qk.record_stream(cur)
        ctx.sg, ctx.cfg = sg, (F_in, akinds, skinds, p_drop, seed, n_slices)
        ctx.save_for_backward(P, Q, tab, arg_min, arg_max, mean, var)
        return Y

    @staticmethod
    def backward(ctx, dY):
        P, Q, tab, arg_min, arg_max, mean, var = ctx.saved_tensors
        sg: ShardedGraph = ctx.sg
        F_in, akinds, skinds, p_drop, seed, n_slices = ctx.cfg
        g = sg.local
        g.build_transpose()
        dev = dY.device
        n, E = sg.rows, g.E
        A, S = len(akinds), len(skinds)
        sl = _slices(F_in, n_slices)
        need_sq = 4 in akinds or 5 in akinds
        dY = dY.contiguous().view(n, S * A * F_in)
        dP = torch.empty((n, F_in), dtype=torch.float32, device=dev)
        dQ = torch.empty((n, F_in), dtype=torch.float32, device=dev)
        G = torch.empty((E, F_in), dtype=torch.float32, device=dev)       # CSC order, full width
        comm = _comm_stream(dev)
        cur = torch.cuda.current_stream(dev)
        ak, sk = _lib.i32_array(akinds), _lib.i32_array(skinds)
        l = _lib.lib()
        pending = []
        for s_ in sl:
            w = s_.stop - s_.start
            q_base, Qk = None, None
            if need_sq:                                 # std/var backward needs m_e again: re-gather Q
                with torch.cuda.stream(comm):
                    comm.wait_stream(cur)
                    Qk = all_gather_rows(Q[:, s_].contiguous(), sg.max_rows, sg.group)
                cur.wait_stream(comm)
                Qk.record_stream(cur)
                q_base = Qk.data_ptr() - 4 * s_.start
            MF.k1_backward_dst(g, P, None, None, None, T=1, F_in=F_in, akinds=akinds, skinds=skinds, tab=tab,
                               p_drop=p_drop, seed=seed, dY=dY, arg_min=arg_min, arg_max=arg_max, mean=mean, var=var,
                               gslot=g.csr2csc, G=G, ldg=F_in, dP=dP, lddp=F_in, col0=s_.start, ncols=w,
                               q_ptr=q_base, ldq=w)
            part = torch.empty((g.n_src, w), dtype=torch.float32, device=dev)
            with _lib.kernel_scope("mma_segment_sum_rows", dev):
                _lib.check(l.mma_segment_sum_rows(_lib.ptr(g.colptr), None, None, g.n_src,
                                                  G.data_ptr() + 4 * s_.start, F_in, w,
                                                  _lib.ptr(part), w, _lib.stream_ptr(dev)), "mma_segment_sum_rows")
            ev = torch.cuda.Event(); ev.record(cur)
            with torch.cuda.stream(comm):               # reduce-scatter of this window overlaps the next window's kernels
                comm.wait_event(ev)
                dQk = reduce_scatter_rows(part, sg.max_rows, n, sg.group)
                done = torch.cuda.Event(); done.record(comm)
            part.record_stream(comm)
            pending.append((s_, dQk, done))
        for s_, dQk, done in pending:
            cur.wait_event(done)
            dQ[:, s_] = dQk
            dQk.record_stream(cur)
        return dP, dQ, None, None, None, None, None, None, None, None


_COMM_STREAMS: Dict[str, torch.cuda.Stream] = {}


def _comm_stream(dev) -> torch.cuda.Stream:
    k = str(dev)
    if k not in _COMM_STREAMS:
        _COMM_STREAMS[k] = torch.cuda.Stream(device=dev)
    return _COMM_STREAMS[k]


def sharded_mmconv_aggregate(P_loc: Tensor, Q_loc: Tensor, sg: ShardedGraph, *, F_in: int,
                             aggregators: Sequence[str], scalers: Sequence[str],
                             avg_deg: Optional[Dict[str, float]] = None, p_drop: float = 0.0, seed: int = 0,
                             n_slices: int = 2, max_deg: Optional[int] = None,
                             sorted_rows: bool = False) -> Tensor:
    """Sharded K1 (single tower): P_loc/Q_loc are this rank's rows [rows, F_in]; returns this
    rank's Y [rows, 1, S*A*F_in] in node order (sorted_rows=True: in the shard's degree-sorted row
    order, sg.local.row_map, without the extra un-permute).  `max_deg` must be the GLOBAL maximum
    in-degree when scalers are used (the lookup table is then identical on all ranks); default:
    all-reduce(max)."""
    for a in aggregators:
        if a not in _lib.AGGR_KINDS:
            raise ValueError(f'Unknown aggregator "{a}".')
    for s in scalers:
        if s not in _lib.SCALER_KINDS:
            raise ValueError(f'Unknown scaler "{s}".')
    akinds = tuple(_lib.AGGR_KINDS[a] for a in aggregators)
    skinds = tuple(_lib.SCALER_KINDS[s] for s in scalers)
    tab = None
    if any(k != 0 for k in skinds):
        if max_deg is None:
            t = torch.tensor([sg.max_deg], device=P_loc.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=sg.group)
            max_deg = int(t.item())
        tab = MF.scale_table(avg_deg, max(max_deg, 1), P_loc.device)
    Y = _ShardedAggregate.apply(P_loc, Q_loc, sg, F_in, akinds, skinds, tab, float(p_drop), int(seed),
                                int(n_slices))
    if sorted_rows or sg.local.row_rank is None:
        return Y
    return Y.index_select(0, sg.local.row_rank)
