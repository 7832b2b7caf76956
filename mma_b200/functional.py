"""Autograd-aware wrappers over the C-ABI kernels (include/mma_b200.h).

PyTorch is used for device memory, streams and autograd plumbing only; every
aggregation op below runs in libmma_b200.so.  No CPU path exists.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import ctypes as C

import torch
from torch import Tensor

from . import _lib
from .graph import Graph

_SCALE_TAB_CACHE: Dict[tuple, Tensor] = {}


def scale_table(avg_deg: Dict[str, float], max_deg: int, device) -> Tensor:
    """[4, max_deg+1] lookup of the degree scalers, evaluated with the reference's own fp32
    expressions on the CPU (graph_regression/mma_conv.py:185-191: `deg` is the clamped
    in-degree as float32, avg_deg[...] a Python float) so the factors are bit-identical
    to the reference's; row k-1 holds scaler kind k (include/mma_b200.h)."""
    key = (float(avg_deg["log"]), float(avg_deg["lin"]), int(max_deg), str(device))
    tab = _SCALE_TAB_CACHE.get(key)
    if tab is None:
        deg = torch.arange(0, max_deg + 1, dtype=torch.float32).clamp_(1)
        amp = torch.log(deg + 1) / avg_deg["log"]
        att = avg_deg["log"] / torch.log(deg + 1)
        lin = deg / avg_deg["lin"]
        inv = avg_deg["lin"] / deg
        tab = torch.stack([amp, att, lin, inv]).contiguous().to(device)
        if len(_SCALE_TAB_CACHE) > 64:
            _SCALE_TAB_CACHE.clear()
        _SCALE_TAB_CACHE[key] = tab
    return tab


def _c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise RuntimeError(f"mma_b200 computes in fp32; got {t.dtype}")
    return t if t.stride(-1) == 1 and t.dim() == 2 else t.contiguous()


def _ld(t: Optional[Tensor]) -> int:
    return 0 if t is None else t.stride(0)

def k1_forward(graph: Graph, P, Q, R, keep, *, T: int, F_in: int, akinds, skinds, tab, p_drop: float, seed: int,
               Y: Tensor, arg_min, arg_max, mean, var, col0: int = 0, ncols: int = 0, local_args: bool = False,
               q_ptr: Optional[int] = None, ldq: Optional[int] = None, seed_dev: Optional[Tensor] = None) -> None:
    """One launch of mmconv_aggregate_fwd (include/mma_b200.h) on `graph`'s destination CSR.
    q_ptr / ldq override Q's base pointer and leading dimension (column-window pipeline of the
    sharded path: a narrow gathered window addressed with global column indices).  seed_dev: int64
    device tensor holding the dropout seed (read at run time: CUDA-graph replays draw fresh masks)."""
    dev = graph.device
    seg = graph.k1_segments()
    seg_ws = (torch.empty((seg.n_slots, 6, T * F_in), dtype=torch.float32, device=dev)
              if seg is not None and seg.n_slots else None)
    a, keepalive = _k1_args(graph, seg, seg_ws, P, Q, R, keep, T, F_in, akinds, skinds, tab, p_drop, seed, seed_dev, q_ptr,
                            ldq, col0, ncols, local_args)
    a.Y, a.ldy = _lib.ptr(Y), Y.stride(0)
    a.arg_min, a.arg_max, a.stat_mean, a.stat_var = _lib.ptr(arg_min), _lib.ptr(arg_max), _lib.ptr(mean), _lib.ptr(var)
    with _lib.kernel_scope("mmconv_aggregate_fwd", dev):
        _lib.check(_lib.lib().mmconv_aggregate_fwd_args(C.byref(a), _lib.stream_ptr(dev)), "mmconv_aggregate_fwd")


def _k1_args(graph: Graph, seg, seg_ws, P, Q, R, keep, T, F_in, akinds, skinds, tab, p_drop, seed, seed_dev, q_ptr, ldq,
             col0, ncols, local_args):
    """The versioned argument block (include/mma_b200.h: mma_k1_args_t) for `graph`'s destination CSR and these
    operands; the second value keeps the host arrays it points to alive until the call has been made."""
    ak, sk = _lib.i32_array(akinds), _lib.i32_array(skinds)
    chunks = graph.k1_chunks()
    a = _lib.K1Args()
    a.struct_size = C.sizeof(_lib.K1Args)
    a.flags = _lib.K1_ARGS_LOCAL if local_args else 0
    a.rowptr, a.col, a.perm, a.edge_gid = _lib.ptr(graph.rowptr), _lib.ptr(graph.col), _lib.ptr(graph.perm), _lib.ptr(graph.gid)
    a.E_total = graph.E_total
    a.row_map, a.rng_row, a.rng_row0 = _lib.ptr(graph.row_map), _lib.ptr(graph.rng_row), int(graph.rng_row0)
    a.row_chunks, a.n_chunks = _lib.ptr(chunks), 0 if chunks is None else chunks.numel() - 1
    if seg is not None:
        a.vrowptr, a.n_vrows = _lib.ptr(seg.vrowptr), seg.n_vrows
        a.seg_tab, a.split_tab, a.n_split = _lib.ptr(seg.seg_tab), _lib.ptr(seg.split_tab), seg.n_split
    a.seg_ws = _lib.ptr(seg_ws)
    a.n_rows, a.E = graph.n_dst, graph.E
    a.P, a.ldp = _lib.ptr(P), _ld(P)
    a.Q, a.ldq = (_lib.ptr(Q) if q_ptr is None else q_ptr), (_ld(Q) if ldq is None else ldq)
    a.R, a.ldr, a.keep, a.ldk = _lib.ptr(R), _ld(R), _lib.ptr(keep), _ld(keep)
    a.p_drop, a.seed, a.seed_dev = float(p_drop), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_dev)
    a.T, a.F_in, a.A, a.S = T, F_in, len(akinds), len(skinds)
    a.aggr_kinds, a.scaler_kinds = C.cast(ak, C.c_void_p), C.cast(sk, C.c_void_p)
    a.scale_tab, a.tab_stride = _lib.ptr(tab), 0 if tab is None else tab.shape[1]
    a.col0, a.ncols = col0, ncols
    return a, (ak, sk, chunks)


def k1_backward_dst(graph: Graph, P, Q, R, keep, *, T: int, F_in: int, akinds, skinds, tab, p_drop: float,
                    seed: int, dY: Tensor, arg_min, arg_max, mean, var, gslot, G, ldg: int, dP, lddp: int,
                    col0: int = 0, ncols: int = 0, local_args: bool = False, q_ptr: Optional[int] = None,
                    ldq: Optional[int] = None, seed_dev: Optional[Tensor] = None) -> None:
    """One launch of mmconv_aggregate_bwd_dst (destination pass of K1's backward)."""
    dev = graph.device
    seg = graph.k1_segments()
    seg_ws = (torch.empty((seg.n_slots, T * F_in), dtype=torch.float32, device=dev)
              if seg is not None and seg.n_slots else None)
    a, keepalive = _k1_args(graph, seg, seg_ws, P, Q, R, keep, T, F_in, akinds, skinds, tab, p_drop, seed, seed_dev, q_ptr,
                            ldq, col0, ncols, local_args)
    a.Y, a.ldy = _lib.ptr(dY), dY.stride(0)
    a.arg_min, a.arg_max, a.stat_mean, a.stat_var = _lib.ptr(arg_min), _lib.ptr(arg_max), _lib.ptr(mean), _lib.ptr(var)
    a.gslot, a.G, a.ldg, a.dP, a.lddp = _lib.ptr(gslot), _lib.ptr(G), ldg, _lib.ptr(dP), lddp
    with _lib.kernel_scope("mmconv_aggregate_bwd_dst", dev):
        _lib.check(_lib.lib().mmconv_aggregate_bwd_dst_args(C.byref(a), _lib.stream_ptr(dev)), "mmconv_aggregate_bwd_dst")


class _MMConvAggregate(torch.autograd.Function):
    """Y = fused gather + mask-add + dropout + multi-aggregate + cumulative scalers (K1)."""

    @staticmethod
    def forward(ctx, P, Q, R, keep, graph: Graph, T: int, F_in: int, akinds: tuple, skinds: tuple,
                tab: Optional[Tensor], p_drop: float, seed: int):
        dev = _lib.require_cuda(P, Q, R, keep, graph.rowptr)
        P, Q, R, keep = _c(P), _c(Q), _c(R), _c(keep)
        F = T * F_in
        for name, t, rows in (("P", P, graph.n_dst), ("Q", Q, None), ("R", R, graph.E), ("keep", keep, graph.E)):
            if t is not None and (t.shape[1] != F or (rows is not None and t.shape[0] != rows)):
                raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected [{rows}, {F}]")
        A, S = len(akinds), len(skinds)
        n = graph.n_dst
        Y = torch.empty((n, T, S * A * F_in), dtype=torch.float32, device=dev)
        has_min, has_max = 2 in akinds, 3 in akinds
        need_sq = 4 in akinds or 5 in akinds
        arg_min = torch.empty((n, F), dtype=torch.int32, device=dev) if has_min else None
        arg_max = torch.empty((n, F), dtype=torch.int32, device=dev) if has_max else None
        stat_mean = torch.empty((n, F), dtype=torch.float32, device=dev) if need_sq else None
        stat_var = torch.empty((n, F), dtype=torch.float32, device=dev) if need_sq else None
        k1_forward(graph, P, Q, R, keep, T=T, F_in=F_in, akinds=akinds, skinds=skinds, tab=tab, p_drop=p_drop,
                   seed=seed, Y=Y, arg_min=arg_min, arg_max=arg_max, mean=stat_mean, var=stat_var)
        ctx.graph, ctx.cfg = graph, (T, F_in, akinds, skinds, p_drop, seed)
        ctx.has = (P is not None, Q is not None, R is not None)
        ctx.save_for_backward(P, Q, R, keep, tab, arg_min, arg_max, stat_mean, stat_var)
        ctx.mark_non_differentiable(*[t for t in (arg_min, arg_max) if t is not None])
        E_t = torch.empty(0, dtype=torch.int32, device=dev)
        return Y, (arg_min if has_min else E_t), (arg_max if has_max else E_t)

    @staticmethod
    def backward(ctx, dY, _ga, _gb):
        P, Q, R, keep, tab, arg_min, arg_max, stat_mean, stat_var = ctx.saved_tensors
        graph: Graph = ctx.graph
        T, F_in, akinds, skinds, p_drop, seed = ctx.cfg
        need_P = P is not None and ctx.needs_input_grad[0]
        need_Q = Q is not None and ctx.needs_input_grad[1]
        need_R = R is not None and ctx.needs_input_grad[2]
        if not (need_P or need_Q or need_R):
            return (None,) * 12
        dev = dY.device
        F = T * F_in
        n, E = graph.n_dst, graph.E
        dY = dY.contiguous().view(n, -1)
        A, S = len(akinds), len(skinds)
        if E == 0:
            z = lambda t, need: torch.zeros_like(t) if need else None
            return (z(P, need_P), z(Q, need_Q), z(R, need_R)) + (None,) * 9
        dP = torch.empty((n, F), dtype=torch.float32, device=dev) if need_P else None
        G = gslot = None
        if need_R:                      # G in original edge order IS dL/dR
            G = torch.empty((E, F), dtype=torch.float32, device=dev)
            gslot = graph.perm
        elif need_Q:                    # G in CSC order: the source pass streams it sequentially
            graph.build_transpose()
            G = torch.empty((E, F), dtype=torch.float32, device=dev)
            gslot = graph.csr2csc
        l = _lib.lib()
        k1_backward_dst(graph, P, Q, R, keep, T=T, F_in=F_in, akinds=akinds, skinds=skinds, tab=tab, p_drop=p_drop,
                        seed=seed, dY=dY, arg_min=arg_min, arg_max=arg_max, mean=stat_mean, var=stat_var,
                        gslot=gslot, G=G, ldg=F, dP=dP, lddp=F)
        dQ = None
        if need_Q:
            graph.build_transpose()
            dQ = torch.empty((Q.shape[0], F), dtype=torch.float32, device=dev)
            if Q.shape[0] != graph.n_src:
                raise RuntimeError("Q rows != number of source nodes of the graph")
            idx = graph.perm_t if need_R else None
            with _lib.kernel_scope("mma_segment_sum_rows", dev):
                _lib.check(l.mma_segment_sum_rows(_lib.ptr(graph.colptr), _lib.ptr(idx), None, graph.n_src,
                                                  _lib.ptr(G), F, F, _lib.ptr(dQ), F,
                                                  _lib.stream_ptr(dev)), "mma_segment_sum_rows")
        return (dP, dQ, G if need_R else None) + (None,) * 9


def mmconv_aggregate(P: Optional[Tensor], Q: Optional[Tensor], R: Optional[Tensor], graph: Graph, *,
                     towers: int, F_in: int, aggregators: Sequence[str], scalers: Sequence[str],
                     avg_deg: Optional[Dict[str, float]] = None, keep: Optional[Tensor] = None,
                     p_drop: float = 0.0, seed: int = 0, return_args: bool = False):
    """Fused MultiMaskConv aggregate (K1).  m_e = ((P[dst]+Q[src])+R[e]) * keepscale, then the
    reductions + cumulative scalers of MMAConv.aggregate (graph_regression/mma_conv.py:159-196).
    Returns Y [n_dst, towers, S*A*F_in] (and original-edge-id arg_min/arg_max [n_dst, towers*F_in])."""
    for a in aggregators:
        if a not in _lib.AGGR_KINDS:
            raise ValueError(f'Unknown aggregator "{a}".')
    for s in scalers:
        if s not in _lib.SCALER_KINDS:
            raise ValueError(f'Unknown scaler "{s}".')
    if len(aggregators) > _lib.MAX_AGGR or len(scalers) > _lib.MAX_SCALER:
        raise _lib.MMAError("more than 8 aggregators or scalers in one call")
    akinds = tuple(_lib.AGGR_KINDS[a] for a in aggregators)
    skinds = tuple(_lib.SCALER_KINDS[s] for s in scalers)
    tab = None
    if any(k != 0 for k in skinds):
        if avg_deg is None:
            raise ValueError("avg_deg is required for non-identity scalers")
        tab = scale_table(avg_deg, max(graph.max_deg, 1), graph.device)
    Y, amin, amax = _MMConvAggregate.apply(P, Q, R, keep, graph, towers, F_in, akinds, skinds, tab,
                                           float(p_drop), int(seed))
    if return_args:
        return Y, (amin if amin.numel() else None), (amax if amax.numel() else None)
    return Y


class _SegmentSumRows(torch.autograd.Function):
    """out[i] = sum_k val[k] * src[idx[k]] over CSR segments (K3); backward on the transposed CSR."""

    @staticmethod
    def forward(ctx, src, ptr, idx, val, n_rows, ptr_t, idx_t, val_t):
        dev = _lib.require_cuda(src, ptr)
        src = _c(src)
        F = src.shape[1]
        out = torch.empty((n_rows, F), dtype=torch.float32, device=dev)
        with _lib.kernel_scope("mma_segment_sum_rows", dev):
            _lib.check(_lib.lib().mma_segment_sum_rows(_lib.ptr(ptr), _lib.ptr(idx), _lib.ptr(val), n_rows,
                                                       _lib.ptr(src), src.stride(0), F, _lib.ptr(out), F,
                                                       _lib.stream_ptr(dev)), "mma_segment_sum_rows")
        ctx.t = (ptr_t, idx_t, val_t, src.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        ptr_t, idx_t, val_t, n_src = ctx.t
        if ptr_t is None:
            raise RuntimeError("segment_sum_rows: transposed structure not provided, cannot backpropagate")
        g = g.contiguous()
        dev, F = g.device, g.shape[1]
        d = torch.empty((n_src, F), dtype=torch.float32, device=dev)
        with _lib.kernel_scope("mma_segment_sum_rows", dev):
            _lib.check(_lib.lib().mma_segment_sum_rows(_lib.ptr(ptr_t), _lib.ptr(idx_t), _lib.ptr(val_t), n_src,
                                                       _lib.ptr(g), F, F, _lib.ptr(d), F,
                                                       _lib.stream_ptr(dev)), "mma_segment_sum_rows")
        return (d,) + (None,) * 7


def segment_sum_rows(src: Tensor, ptr: Tensor, idx: Optional[Tensor], val: Optional[Tensor], n_rows: int,
                     ptr_t: Optional[Tensor] = None, idx_t: Optional[Tensor] = None,
                     val_t: Optional[Tensor] = None) -> Tensor:
    return _SegmentSumRows.apply(src, ptr, idx, val, n_rows, ptr_t, idx_t, val_t)


def dropout_keep_scale(p: float, seed: int, E: int, F: int, device, stream_id: int = 0,
                       graph: Optional[Graph] = None) -> Tensor:
    """The keep-scale tensor [E,F] (0 or 1/(1-p), original edge order) that the kernels generate on
    the fly -- for injecting the identical dropout into the CPU oracle in tests.  With `graph`: K1's
    stream (keyed by destination row and in-row position); without: K2's stream for aggregator slot
    `stream_id` (keyed by edge id)."""
    out = torch.empty((E, F), dtype=torch.float32, device=device)
    if graph is not None:
        if graph.E != E:
            raise RuntimeError("graph has a different number of edges")
        with torch.cuda.device(device):
            _lib.check(_lib.lib().mma_dropout_keep_scale_rows(
                _lib.ptr(graph.rowptr), _lib.ptr(graph.perm), _lib.ptr(graph.rng_row), int(graph.rng_row0),
                graph.n_dst, E, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, F, _lib.ptr(out), F,
                _lib.stream_ptr(out.device)), "mma_dropout_keep_scale_rows")
        return out
    with torch.cuda.device(device):
        _lib.check(_lib.lib().mma_dropout_keep_scale(float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id),
                                                     E, F, _lib.ptr(out), F, _lib.stream_ptr(out.device)),
                   "mma_dropout_keep_scale")
    return out


class _NcAggregate(torch.autograd.Function):
    """All A masked neighbour sums + combine of node_classification/layers.py:201-728 (K2)."""

    @staticmethod
    def forward(ctx, X, PA, QA, nbr, acts: tuple, combs: tuple, keep, p_drop: float, seed: int, seed_dev=None):
        dev = _lib.require_cuda(X, PA, QA, nbr.rowptr)
        if seed_dev is not None:
            seed_dev = seed_dev.clone()      # this call's seed: the backward must see the same value
        X, PA, QA = _c(X), _c(PA), _c(QA)
        N, F = X.shape
        A = len(acts)
        if PA.shape != (N, A * F) or QA.shape != (N, A * F):
            raise RuntimeError(f"PA/QA must be [{N}, {A * F}], got {tuple(PA.shape)} / {tuple(QA.shape)}")
        if keep is not None:
            keep = keep.contiguous()
            if keep.shape != (A, nbr.E, F):
                raise RuntimeError(f"keep must be [{A}, {nbr.E}, {F}]")
        OUT = torch.empty((A, N, F), dtype=torch.float32, device=dev)
        S = torch.empty((A, N, F), dtype=torch.float32, device=dev)
        ak, ck = _lib.i32_array(acts), _lib.i32_array(combs)
        with _lib.kernel_scope("mma_nc_aggregate_fwd", dev):
            _lib.check(_lib.lib().mma_nc_aggregate_fwd(
                _lib.ptr(nbr.rowptr), _lib.ptr(nbr.col), N, nbr.E, _lib.ptr(X), X.stride(0),
                _lib.ptr(PA), PA.stride(0), _lib.ptr(QA), QA.stride(0), F, A, ak, ck, _lib.ptr(keep),
                float(p_drop), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_dev), _lib.ptr(OUT), _lib.ptr(S),
                _lib.stream_ptr(dev)), "mma_nc_aggregate_fwd")
        ctx.nbr, ctx.cfg = nbr, (acts, combs, p_drop, seed)
        ctx.save_for_backward(X, PA, QA, keep, S, seed_dev)
        return OUT

    @staticmethod
    def backward(ctx, dOUT):
        X, PA, QA, keep, S, seed_dev = ctx.saved_tensors
        nbr = ctx.nbr
        acts, combs, p_drop, seed = ctx.cfg
        dev = dOUT.device
        N, F = X.shape
        A = len(acts)
        dOUT = dOUT.contiguous()
        gS = torch.empty((N, A * F), dtype=torch.float32, device=dev)
        dXdir = torch.empty((N, F), dtype=torch.float32, device=dev)
        dPA = torch.empty((N, A * F), dtype=torch.float32, device=dev)
        dQA = torch.empty((N, A * F), dtype=torch.float32, device=dev)
        dXn = torch.empty((N, F), dtype=torch.float32, device=dev)
        ak, ck = _lib.i32_array(acts), _lib.i32_array(combs)
        nbr.build_transpose()
        l = _lib.lib()
        sd = int(seed) & 0xFFFFFFFFFFFFFFFF
        with _lib.kernel_scope("mma_nc_aggregate_bwd_dst", dev):
            _lib.check(l.mma_nc_aggregate_bwd_dst(
                _lib.ptr(nbr.rowptr), _lib.ptr(nbr.col), N, nbr.E, _lib.ptr(X), X.stride(0),
                _lib.ptr(PA), PA.stride(0), _lib.ptr(QA), QA.stride(0), F, A, ak, ck, _lib.ptr(keep),
                float(p_drop), sd, _lib.ptr(seed_dev), _lib.ptr(S), _lib.ptr(dOUT), _lib.ptr(gS), _lib.ptr(dXdir),
                _lib.ptr(dPA), A * F, _lib.stream_ptr(dev)), "mma_nc_aggregate_bwd_dst")
        with _lib.kernel_scope("mma_nc_aggregate_bwd_src", dev):
            _lib.check(l.mma_nc_aggregate_bwd_src(
                _lib.ptr(nbr.colptr), _lib.ptr(nbr.row_t), _lib.ptr(nbr.perm_t), N, nbr.E,
                _lib.ptr(X), X.stride(0), _lib.ptr(PA), PA.stride(0), _lib.ptr(QA), QA.stride(0), F, A, ak,
                _lib.ptr(keep), float(p_drop), sd, _lib.ptr(seed_dev), _lib.ptr(gS), _lib.ptr(dQA), A * F, _lib.ptr(dXn), F,
                _lib.stream_ptr(dev)), "mma_nc_aggregate_bwd_src")
        return dXdir + dXn, dPA, dQA, None, None, None, None, None, None, None


def nc_aggregate(X: Tensor, PA: Tensor, QA: Tensor, nbr, acts: Sequence[int], combs: Sequence[int],
                 keep: Optional[Tensor] = None, p_drop: float = 0.0, seed: int = 0,
                 seed_dev: Optional[Tensor] = None) -> Tensor:
    """OUT[a] = combine_a(x_i, sum_j act_a(PA[i,a] + QA[j,a]) * keepscale * x_j)  -> [A, N, F].
    `nbr` is a NeighbourLists (CSR of add_all; edge id == CSR position).  `seed_dev` (int64 [1] on the device): the
    dropout key is read from it at run time instead of `seed` (CUDA-graph replays then draw fresh masks)."""
    if len(acts) > _lib.MAX_AGGR:
        raise _lib.MMAError("more than 8 aggregators in one call")
    return _NcAggregate.apply(X, PA, QA, nbr, tuple(int(a) for a in acts), tuple(int(c) for c in combs),
                              keep, float(p_drop), int(seed), seed_dev)


# ------------------------------------------------------------------------------------------
# post-transform with the degree scalers folded into the weights
# ------------------------------------------------------------------------------------------
def cumulative_scale_factors(scalers: Sequence[str], avg_deg: Dict[str, float], degrees: Sequence[int]) -> Tensor:
    """[S, len(degrees)] fp32: block s of MMAConv.aggregate's output is the raw aggregate times
    prod_{s'<=s} f_{s'}(deg) (cumulative re-assignment, graph_regression/mma_conv.py:181-195, Q4);
    evaluated on the CPU with the reference's expressions."""
    deg = torch.tensor(list(degrees), dtype=torch.float32).clamp_(1)
    cur = torch.ones_like(deg)
    rows = []
    for sc in scalers:
        if sc == "identity":
            pass
        elif sc == "amplification":
            cur = cur * (torch.log(deg + 1) / avg_deg["log"])
        elif sc == "attenuation":
            cur = cur * (avg_deg["log"] / torch.log(deg + 1))
        elif sc == "linear":
            cur = cur * (deg / avg_deg["lin"])
        elif sc == "inverse_linear":
            cur = cur * (avg_deg["lin"] / deg)
        else:
            raise ValueError(f'Unknown scaler "{sc}".')
        rows.append(cur)
    return torch.stack(rows)


class _ScaledPost(torch.autograd.Function):
    """H[r] = sum_s cum_s(deg_r) * (Z[r] @ W_s^T)  for degree-sorted rows r.

    The reference materialises Y = cat_s(Z * cum_s) [N, S*A*F] and multiplies by the post weight
    (mma_conv.py:181-196, 132-133).  Rows of equal degree share their S factors, so for a
    contiguous range of equal-degree rows this is ONE GEMM with the effective weight
    W_eff(d) = sum_s cum_s(d) W_s: S times fewer flops and no [N, S*A*F] tensor.  Degrees
    with fewer than `min_rows` nodes (the tail of a skewed distribution) are handled together by
    the literal formula."""

    @staticmethod
    def forward(ctx, Z, Wy, cum, ranges, tail_idx, tail_bucket, S: int):
        n, K = Z.shape
        Fo = Wy.shape[0]
        Ws = Wy.view(Fo, S, K)
        H = torch.empty((n, Fo), dtype=Z.dtype, device=Z.device)
        big = [b for b, _, _ in ranges]
        Weff = None
        if big:
            cb = cum[:, big]                                                      # [S, B]
            Weff = torch.einsum("sb,osk->bok", cb, Ws).contiguous()               # [B, Fo, K]
            for i, (_, lo, hi) in enumerate(ranges):
                torch.mm(Z[lo:hi], Weff[i].t(), out=H[lo:hi])
        if tail_idx is not None and tail_idx.numel() > 0:
            Zt = Z.index_select(0, tail_idx)
            ct = cum[:, tail_bucket].t()                                          # [nt, S]
            Yt = (Zt.unsqueeze(1) * ct.unsqueeze(2)).reshape(Zt.shape[0], S * K)
            H.index_copy_(0, tail_idx, Yt @ Wy.t())
        ctx.save_for_backward(Z, Wy, cum, Weff, tail_idx, tail_bucket)
        ctx.ranges, ctx.S = ranges, S
        return H

    @staticmethod
    def backward(ctx, dH):
        Z, Wy, cum, Weff, tail_idx, tail_bucket = ctx.saved_tensors
        ranges, S = ctx.ranges, ctx.S
        n, K = Z.shape
        Fo = Wy.shape[0]
        dH = dH.contiguous()
        dZ = torch.empty_like(Z) if ctx.needs_input_grad[0] else None
        dWy = None
        if ranges:
            dWeff = torch.empty_like(Weff) if ctx.needs_input_grad[1] else None
            for i, (_, lo, hi) in enumerate(ranges):
                if dZ is not None:
                    torch.mm(dH[lo:hi], Weff[i], out=dZ[lo:hi])
                if dWeff is not None:
                    torch.mm(dH[lo:hi].t(), Z[lo:hi], out=dWeff[i])
            if dWeff is not None:
                cb = cum[:, [b for b, _, _ in ranges]]
                dWy = torch.einsum("sb,bok->osk", cb, dWeff).reshape(Fo, S * K)
        if tail_idx is not None and tail_idx.numel() > 0:
            Zt = Z.index_select(0, tail_idx)
            ct = cum[:, tail_bucket].t()
            dHt = dH.index_select(0, tail_idx)
            if dZ is not None:
                dYt = (dHt @ Wy).view(-1, S, K)
                dZ.index_copy_(0, tail_idx, (dYt * ct.unsqueeze(2)).sum(dim=1))
            if ctx.needs_input_grad[1]:
                Yt = (Zt.unsqueeze(1) * ct.unsqueeze(2)).reshape(Zt.shape[0], S * K)
                g = dHt.t() @ Yt
                dWy = g if dWy is None else dWy + g
        return dZ, dWy, None, None, None, None, None


def scaled_post(Z: Tensor, Wy: Tensor, graph: Graph, scalers: Sequence[str], avg_deg: Dict[str, float],
                min_rows: int = 512) -> Tensor:
    """Z [n, A*F] (rows in graph.row_map order, raw aggregates) -> H [n, F_out] (same row order)
    == cat_s(Z * cum_s(deg)) @ Wy^T.  `graph` must have degree-sorted rows (graph.buckets)."""
    if graph.buckets is None:
        raise RuntimeError("scaled_post needs a Graph built with sort_rows=True")
    S = len(scalers)
    key = (tuple(scalers), float(avg_deg["log"]), float(avg_deg["lin"]), int(min_rows), str(Z.device))
    plans = graph.__dict__.setdefault("_post_plans", {})
    plan = plans.get(key)
    if plan is None:
        degs = [d for d, _, _ in graph.buckets]
        cum = cumulative_scale_factors(scalers, avg_deg, degs).to(Z.device)       # [S, n_buckets]
        ranges, tail_rows, tail_b = [], [], []
        for b, (d, lo, hi) in enumerate(graph.buckets):
            if hi - lo >= min_rows:
                ranges.append((b, lo, hi))
            else:
                tail_rows.append(torch.arange(lo, hi, dtype=torch.int64))
                tail_b.append(torch.full((hi - lo,), b, dtype=torch.int64))
        tail_idx = torch.cat(tail_rows).to(Z.device) if tail_rows else None
        tail_bucket = torch.cat(tail_b).to(Z.device) if tail_b else None
        plan = (cum, tuple(ranges), tail_idx, tail_bucket)
        plans[key] = plan
    cum, ranges, tail_idx, tail_bucket = plan
    return _ScaledPost.apply(Z, Wy, cum, ranges, tail_idx, tail_bucket, S)
