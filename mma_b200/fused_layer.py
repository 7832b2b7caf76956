"""The whole MultiMaskConv layer (towers == 1) as ONE autograd node over the C-ABI kernels.

reference (graph_regression/mma_conv.py)                  here (CSR rows degree-sorted, everything else in node order)
--------------------------------------------------------  ------------------------------------------------
x.view(-1,1,F).repeat(1,T,1)                (:128)        never materialised
mask Linear over cat([x_i,x_j])   (:146-152, mask_aggr)   [P|Q|XW] = x [W_i;W_j;W_lin W_x]^T + [b;0;b']  (G1, tcgen05)
dropout + A scatters + degree     (:157-179)              Z = K1(P, Q)      raw aggregates [N, A*F]
S cumulative scalers, cat, post Linear (:181-196,132-133) out = Z W_c(deg)^T + XW   (G2: grouped tcgen05 GEMM,
                                                            one effective weight per equal-degree row range)
lin                                (:136)                  composed into G1 / G2 (no nonlinearity in between): W_lin W_eff(d),
                                                            W_lin W_x; G2's epilogue scatters rows back to node order
autograd backward                  (mma.py:157)            dgrad = same GEMM kernel on transposed weights, wgrad =
                                                            split-K tcgen05 kernel + fixed-order slab reduction,
                                                            K1 destination pass + transpose-CSR pass (no atomics)

cat([x, out]) @ W_post^T is split as x W_x^T + Y W_y^T, and x W_x^T rides along in G1 (x is read once).
The GEMMs are 3xTF32 (fp32-accurate, see csrc/gemm_tf32x3.cu).  No CPU path.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from . import tc_gemm as tg
from . import functional as MF
from .functional import cumulative_scale_factors
from .graph import Graph

# Run the backward's reduce-scatter on a side stream under the weight-gradient GEMM.  Off by default: NCCL's
# CTAs take SMs away from the persistent one-CTA-per-SM GEMM, whose displaced CTAs then run as a second wave
# (measured: the GEMM took 2x as long, cancelling the overlap).
OVERLAP_EXCHANGE = os.environ.get("MMA_OVERLAP_EXCHANGE", "0") == "1"
# How the sharded layer's two exchanges travel: "peer" = copy engines over NVLink peer memory, pipelined per feature
# window (mma_b200/peer.py, the default); "nccl" = one all-gather / reduce-scatter (the round-1 path, kept for A/B runs).
EXCHANGE = os.environ.get("MMA_EXCHANGE", "peer")
G_ORDER = os.environ.get("MMA_G_ORDER", "csc")       # order of the per-edge gradient rows G (see _k1_bwd)
# Weight-space algebra (composition of W_lin, W_post and the scaler coefficients, and its gradient): "1" = the library's
# small GEMM + composition kernels (csrc/weight_prep.cu, ~15 launches per step); "0" = the torch formulation (~60 launches,
# kept as the readable statement of the algebra and for A/B runs).
WPREP = os.environ.get("MMA_WPREP", "1") == "1"
# Sharded backward: the mask GEMMs' dQ-independent two thirds (dx = dP W_i + dOut W_cx, the dP / dOut rows of dW1, the
# bias gradient) run BEFORE the wait for the reduce-scatter and only the dQ third after it.  "auto": from 4 ranks up
# (there the reduce-scatter is exposed; at 2 ranks it hides completely and the split only adds a pass over dx and x).
SPLIT_MASK_BWD = os.environ.get("MMA_SPLIT_MASK_BWD", "auto")

# Parity hook (bench.py --verify, tests): when set to a dict, every forward leaves references to its raw aggregates
# and arg indices there (Z in CSR-row order, arg_min / arg_max as CSR slots), so a sharded run can be compared bit for
# bit with the single-GPU run.  None = off (the default; nothing is retained).
KEEP_LAST: Optional[dict] = None


def fold_blocks(akinds: Sequence[int]):
    """Which aggregate blocks K1 has to MATERIALISE for the aggregator kinds `akinds` (sum 0, mean 1, min 2, max 3,
    var 4, std 5).  `mean` is `sum / deg`, and the post weight is already one matrix per in-degree, so a mean block is
    never written: its weight block, divided by the degree, is added onto the sum block's (W_sum + W_mean / d).
    Returns (kinds of the materialised blocks, block of each aggregator, whether its coefficient is 1 / deg)."""
    mat, block_of, inv = [], [], []
    for k in akinds:
        km = 0 if k == 1 else k
        if km not in mat:
            mat.append(km)
        block_of.append(mat.index(km))
        inv.append(k == 1)
    return tuple(mat), tuple(block_of), tuple(inv)


def materialised_blocks(aggregators: Sequence[str]) -> int:
    """Number of [N, F] aggregate blocks the fused layer's K1 writes for these aggregator names (bench.py's byte count)."""
    return len(fold_blocks(tuple(_lib.AGGR_KINDS[a] for a in aggregators))[0])


class PostPlan:
    """Row ranges of equal in-degree (graph.buckets) -> tile / slab tables of the grouped GEMMs, and the coefficients
    that turn the S x A blocks of the post weight into ONE effective weight per degree over the materialised blocks:
    coef[b, s, a, m] = cum_s(d_b) * [block_of(a) == m] * (1 / max(d_b, 1) if a is a mean else 1)."""

    MAX_BIG = 256

    def __init__(self, graph: Graph, scalers: Sequence[str], avg_deg: Dict[str, float], min_rows: Optional[int], Fo: int,
                 akinds: Sequence[int], F: int):
        dev = graph.device
        degs = [d for d, _, _ in graph.buckets]
        self.S = len(scalers)
        self.akinds = tuple(akinds)
        self.mat, self.block_of, self.inv = fold_blocks(akinds)
        A, Am = len(akinds), len(self.mat)
        K = Am * F                                                                       # width of Z / dZ
        self.K = K
        self.cum = cumulative_scale_factors(scalers, avg_deg, degs).to(dev)           # [S, n_buckets]
        degf = torch.tensor(degs, dtype=torch.float32).clamp_(min=1)
        fold = torch.zeros((len(degs), A, Am), dtype=torch.float32)                    # [n_buckets, A, Am]
        for a in range(A):
            fold[:, a, self.block_of[a]] = (1.0 / degf) if self.inv[a] else 1.0
        self.fold = fold.to(dev)
        self.block_of_t = torch.tensor(self.block_of, dtype=torch.int32, device=dev)
        big, tiles, tiles_t, slabs, seg_ptr, tail_rows, tail_b = [], [], [], [], [0], [], []
        if min_rows is None or min_rows <= 0:
            # auto: every degree range gets its own effective weight (no tail path at all) as long as there are at most
            # MAX_BIG of them -- a graph with few distinct degrees, e.g. config 4; on a skewed degree distribution the
            # MAX_BIG largest ranges do, and the many tiny hub ranges beyond them take the literal formula
            sizes = sorted((hi - lo for _, lo, hi in graph.buckets), reverse=True)
            min_rows = 1
            if len(sizes) > self.MAX_BIG:
                min_rows = sizes[self.MAX_BIG - 1]
                if sizes[self.MAX_BIG] == min_rows:          # ties at the cut: keep the count bounded
                    min_rows += 1
        self.min_rows = int(min_rows)
        for b, (_, lo, hi) in enumerate(graph.buckets):
            if hi - lo >= min_rows:
                i = len(big)
                big.append(b)
                for r in range(lo, hi, tg.BM):
                    tiles.append((r, hi, i * Fo, 0))
                    tiles_t.append((r, hi, i * K, 0))
                for r in range(lo, hi, tg.MAX_SLAB_ROWS):
                    slabs.append((r, min(hi, r + tg.MAX_SLAB_ROWS), len(slabs), 0))
                seg_ptr.append(len(slabs))
            else:
                tail_rows.append(torch.arange(lo, hi, dtype=torch.int64))
                tail_b.append(torch.full((hi - lo,), b, dtype=torch.int64))
        i32 = lambda rows: torch.tensor(rows, dtype=torch.int32, device=dev).reshape(-1, 4)
        self.big = big
        self.cum_big = self.cum[:, big].contiguous() if big else None                 # [S, B]
        # [B, S, A, Am]
        self.coef_big = (self.cum_big.t().reshape(len(big), self.S, 1, 1) * self.fold[big].unsqueeze(1)).contiguous() \
            if big else None
        self.tile_tab, self.tile_tab_t, self.slabs = i32(tiles), i32(tiles_t), i32(slabs)
        self.seg_ptr = torch.tensor(seg_ptr, dtype=torch.int32, device=dev)
        self.tail_idx = torch.cat(tail_rows).to(dev) if tail_rows else None           # CSR rows of the small buckets
        self.tail_bucket = torch.cat(tail_b).to(dev) if tail_b else None
        self.tail_nodes = (graph.row_map.to(torch.int64).index_select(0, self.tail_idx)
                           if tail_rows else None)                                      # their node ids
        # tail rows (literal formula): Y_t[n, s, a, :] = coef_t[n, s, a, m] * Z_t[n, m, :]
        self.tail_coef = ((self.cum[:, self.tail_bucket].t().reshape(-1, self.S, 1, 1)
                           * self.fold[self.tail_bucket].unsqueeze(1)).contiguous() if tail_rows else None)


def post_plan(graph: Graph, scalers, avg_deg, min_rows: Optional[int], Fo: int, akinds, F: int) -> PostPlan:
    key = ("tc", tuple(scalers), float(avg_deg["log"]), float(avg_deg["lin"]), int(min_rows or 0), int(Fo), tuple(akinds), int(F))
    plans = graph.__dict__.setdefault("_post_plans", {})
    p = plans.get(key)
    if p is None:
        p = plans[key] = PostPlan(graph, scalers, avg_deg, min_rows, Fo, akinds, F)
    return p


def _k1_fwd(graph: Graph, P: Tensor, Q: Tensor, R: Optional[Tensor], keep: Optional[Tensor], F: int,
            akinds: Tuple[int, ...], p_drop: float, seed: int, seed_dev: Optional[Tensor] = None):
    dev = P.device
    n, A = graph.n_dst, len(akinds)
    Z = torch.empty((n, A * F), dtype=torch.float32, device=dev)
    has_min, has_max = 2 in akinds, 3 in akinds
    need_sq = 4 in akinds or 5 in akinds
    arg_min = torch.empty((n, F), dtype=torch.int32, device=dev) if has_min else None
    arg_max = torch.empty((n, F), dtype=torch.int32, device=dev) if has_max else None
    mean = torch.empty((n, F), dtype=torch.float32, device=dev) if need_sq else None
    var = torch.empty((n, F), dtype=torch.float32, device=dev) if need_sq else None
    MF.k1_forward(graph, P, Q, R, keep, T=1, F_in=F, akinds=akinds, skinds=(0,), tab=None, p_drop=p_drop, seed=seed,
                  Y=Z, arg_min=arg_min, arg_max=arg_max, mean=mean, var=var, local_args=True, seed_dev=seed_dev)
    return Z, arg_min, arg_max, mean, var


def _k1_bwd(graph: Graph, P, Q, R, keep, F, akinds, p_drop, seed, dZ, arg_min, arg_max, mean, var,
            dP: Tensor, dQ: Tensor, need_R: bool, seed_dev: Optional[Tensor] = None):
    """Destination pass (per-edge gradient rows G, dP) + transpose pass (dQ[j] = sum of G over j's out-edges).
    dP [n_dst, F] and dQ [n_src, F] may be strided views.  Returns dR (or None)."""
    dev = dZ.device
    E = graph.E
    if E == 0:
        dP.zero_(); dQ.zero_()
        return torch.zeros_like(R) if need_R else None
    graph.build_transpose()
    G = torch.empty((E, F), dtype=torch.float32, device=dev)
    gslot = graph.perm if need_R else graph.csr2csc
    idx = graph.perm_t if need_R else None
    if G_ORDER == "csr" and not need_R:
        # A/B variant (MMA_G_ORDER=csr): G in CSR order -- sequential row stores in the destination pass, random row
        # gathers in the transpose pass; same bytes as the default (scattered stores, sequential reads).  Measured on
        # config 4: destination pass 7.10 vs 7.14 ms, transpose pass 2.57 vs 2.47 ms -- no gain (profiles/README.md,
        # round 2); kept for re-measurement only.
        if "_csc2csr" not in graph.__dict__:
            from .graph import invert_perm
            graph.__dict__["_csc2csr"] = invert_perm(graph.csr2csc)
        gslot, idx = None, graph.__dict__["_csc2csr"]
    MF.k1_backward_dst(graph, P, Q, R, keep, T=1, F_in=F, akinds=akinds, skinds=(0,), tab=None, p_drop=p_drop,
                       seed=seed, dY=dZ, arg_min=arg_min, arg_max=arg_max, mean=mean, var=var, gslot=gslot, G=G,
                       ldg=F, dP=dP, lddp=dP.stride(0), local_args=True, seed_dev=seed_dev)
    with _lib.kernel_scope("mma_segment_sum_rows", dev):
        _lib.check(_lib.lib().mma_segment_sum_rows(_lib.ptr(graph.colptr), _lib.ptr(idx), None, graph.n_src, _lib.ptr(G),
                                                   F, F, _lib.ptr(dQ), dQ.stride(0), _lib.stream_ptr(dev)),
                   "mma_segment_sum_rows")
    return G if need_R else None


def _k1_fwd_windows(graph: Graph, P: Tensor, ex, F: int, akinds, p_drop: float, seed: int, seed_dev):
    """K1 over this rank's rows, one launch per feature window of the peer exchange, each as soon as that window of
    the gathered Q has arrived from every rank (column-window ABI: col0 / ncols on the full-width gathered buffer)."""
    dev = P.device
    n, A = graph.n_dst, len(akinds)
    Z = torch.empty((n, A * F), dtype=torch.float32, device=dev)
    has_min, has_max = 2 in akinds, 3 in akinds
    need_sq = 4 in akinds or 5 in akinds
    arg_min = torch.empty((n, F), dtype=torch.int32, device=dev) if has_min else None
    arg_max = torch.empty((n, F), dtype=torch.int32, device=dev) if has_max else None
    mean = torch.empty((n, F), dtype=torch.float32, device=dev) if need_sq else None
    var = torch.empty((n, F), dtype=torch.float32, device=dev) if need_sq else None
    for k, s_ in enumerate(ex.windows):
        ex.wait(k)
        MF.k1_forward(graph, P, ex.recv_q, None, None, T=1, F_in=F, akinds=akinds, skinds=(0,), tab=None, p_drop=p_drop,
                      seed=seed, Y=Z, arg_min=arg_min, arg_max=arg_max, mean=mean, var=var, local_args=True,
                      seed_dev=seed_dev, col0=s_.start, ncols=s_.stop - s_.start)
    return Z, arg_min, arg_max, mean, var


def _k1_bwd_sharded(graph: Graph, P, ex, F, akinds, p_drop, seed, dZ, arg_min, arg_max, mean, var, dP: Tensor,
                    need_sq: bool, seed_dev) -> None:
    """Destination pass at full width, then the transpose pass OWNER BY OWNER (the partial dQ is laid out rank-major
    over all sources): each owner's slice leaves on the copy engine while the next owner's rows are being summed."""
    dev = dZ.device
    graph.build_transpose()
    part = torch.empty((graph.n_src, F), dtype=torch.float32, device=dev)
    if graph.E == 0:
        part.zero_(); dP.zero_()
        ex.push_partial(part)
        return
    G = torch.empty((graph.E, F), dtype=torch.float32, device=dev)       # CSC order
    MF.k1_backward_dst(graph, P, ex.recv_q if need_sq else None, None, None, T=1, F_in=F, akinds=akinds, skinds=(0,),
                       tab=None, p_drop=p_drop, seed=seed, dY=dZ, arg_min=arg_min, arg_max=arg_max, mean=mean, var=var,
                       gslot=graph.csr2csc, G=G, ldg=F, dP=dP, lddp=dP.stride(0), local_args=True, seed_dev=seed_dev)
    mr = ex.max_rows
    for o in ex.owner_order():
        with _lib.kernel_scope("mma_segment_sum_rows", dev):
            _lib.check(_lib.lib().mma_segment_sum_rows(graph.colptr.data_ptr() + 4 * o * mr, None, None, mr,
                                                       _lib.ptr(G), F, F, part.data_ptr() + 4 * o * mr * F, F,
                                                       _lib.stream_ptr(dev)), "mma_segment_sum_rows")
        ex.push_partial_block(o, part)


class _FusedMMAConv(torch.autograd.Function):
    """post_layers == 1: there is no nonlinearity between the post Linear (mma_conv.py:132-133) and `lin`
    (:136), so the two compose:  out = Z (W_lin W_eff(d))^T + x (W_lin W_x)^T + (W_lin b_post + b_lin).
    The composed x-part rides along in G1, H is never materialised, and `lin` costs no GEMM of its own
    in either direction; the gradients of W_lin / W_post are recovered from those of the composed
    weights by small [F_out x K] matrix products.

    Row orders: x, P, Q, XW, dP, dQ, dx and the layer output live in NODE order; only the CSR rows (hence
    Z, dZ and the row tiles of the grouped GEMMs) are degree-sorted.  K1 reaches P / dP through
    graph.row_map, G2's epilogue scatters its rows to node order, and the single row permutation left
    is dOut[row_map] in the backward (one gather pass that also yields the bias gradient)."""

    @staticmethod
    def forward(ctx, x, Wm, bm, Wp, bp, Wl, bl, R, keep, graph, plan: PostPlan, cfg):
        F, akinds, p_drop, seed, seed_dev = cfg[:5]
        n_win, owner = cfg[5:7]
        if seed_dev is not None:
            seed_dev = seed_dev.clone()      # this call's seed: the backward must see the same value
        dev = _lib.require_cuda(x, Wm, Wp, Wl)
        sg = None if isinstance(graph, Graph) else graph          # ShardedGraph: x holds this rank's rows only
        if sg is not None:
            graph = sg.local
        n = graph.n_dst
        A, S = len(akinds), plan.S
        Am, K = len(plan.mat), plan.K                # materialised aggregate blocks (a mean rides on the sum block)
        Fo, Co = Wp.shape[0], Wl.shape[0]
        zeros = lambda k: torch.zeros(k, dtype=torch.float32, device=dev)
        Wx, Wy = Wp[:, :F], Wp[:, F:]
        wprep = WPREP and Wp.stride(1) == 1 and Wl.stride(1) == 1
        if wprep:
            WlW = tg.small_gemm(Wl, Wp)                              # [Co, (S*A+1)*F] = W_lin [W_x | W_{s,a} ...]: one launch
            Wcx = WlW[:, :F]
        else:
            Wy = Wy.contiguous()
            Wcx = Wl @ Wx                                                                           # [Co, F]
        bc = (Wl @ bp if bp is not None else zeros(Co)) + (bl if bl is not None else zeros(Co))
        W1 = torch.cat([Wm[:, :F], Wm[:, F:2 * F], Wcx], dim=0)                                     # [2F+Co, F]
        b1 = torch.cat([bm if bm is not None else zeros(F), zeros(F), bc])
        W1hi, W1lo = tg.split_weight(W1)
        PQX = tg.linear(x, W1hi, W1lo, 2 * F + Co, bias=b1, name="gemm_mask_proj")                  # [n, 2F+Co], node order
        P, Q, XW = PQX[:, :F], PQX[:, F:2 * F], PQX[:, 2 * F:]
        gathered = ex = None
        if sg is not None and EXCHANGE == "peer":
            # the one exchange of the forward, over peer memory: every rank pushes its Q rows, feature window by
            # feature window, into every peer's gathered buffer with the copy engines (mma_b200/peer.py).  The pushes
            # start here and run under the weight-space algebra and under K1 of the earlier windows.
            from .parallel import _slices
            from .peer import exchange_for
            ex = exchange_for(sg, F, _slices(F, n_win), dev, owner)
            ex.begin_call()
            ex.fwd_calls += 1
            ex.push_q(Q, n)
        elif sg is not None:
            # NCCL variant (MMA_EXCHANGE=nccl): one all-gather of Q on the communication stream under the
            # weight-space algebra below; kept for A/B measurements against the copy-engine exchange
            from .parallel import all_gather_rows, _comm_stream
            comm, cur = _comm_stream(dev), torch.cuda.current_stream(dev)
            Qc = Q.contiguous()
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                Qg = all_gather_rows(Qc, sg.max_rows, sg.group)
                done = torch.cuda.Event(); done.record(comm)
            Qc.record_stream(comm)
            gathered = (Qg, done)
        # composed weight of every degree range over the materialised blocks:
        #   W_c(d)[:, m] = sum_s sum_{a -> m} c_s(d) coef_a(d) (W_lin W_{s,a})      (coef = 1 / d for a mean, else 1)
        # S small products, then one contraction with the per-range coefficients (no [B, F_out, K] batched product)
        Wc = hi = lo = None
        bw = None
        if wprep:
            # one composition kernel writes W_c already split for the 3xTF32 GEMMs, and its transpose (the dgrad GEMM's
            # weight) with it; the backward's split W1^T is made here too -- a sharded rank is waiting for the exchange at
            # this point, and the backward's critical path (dgrad -> K1 -> push; slices -> dgrad / wgrad) gets shorter
            wct = None
            if plan.big:
                hi3, lo3, hiT, loT = tg.compose_post_weight(plan.coef_big, WlW, F, F, True)
                hi, lo, wct = hi3.view(-1, K), lo3.view(-1, K), (hiT.view(-1, Co), loT.view(-1, Co))
            bw = (wct, tg.split_weight(W1.t()))
            if sg is not None and (SPLIT_MASK_BWD == "1" or (SPLIT_MASK_BWD == "auto" and sg.world >= 4)) \
                    and F % 128 == 0 and EXCHANGE == "peer":
                # the same transposed weight cut into its dQ-independent part [W_i^T | W_cx^T] and its dQ part W_j^T
                bw = bw + (tg.split_weight(torch.cat([W1[:F], W1[2 * F:]], dim=0).t()), tg.split_weight(W1[F:2 * F].t()))
        else:
            WlWy = torch.matmul(Wl, Wy.view(Fo, S, A * F).permute(1, 0, 2))                         # [S, Co, A*F]
            if plan.big:
                Wc = torch.einsum("bsam,scaf->bcmf", plan.coef_big, WlWy.view(S, Co, A, F)).reshape(-1, Co, K).contiguous()   # [B, Co, K]
                hi, lo = tg.split_weight(Wc.view(-1, K))
            if sg is not None:
                wct = tg.split_weight(Wc.transpose(1, 2).contiguous().view(-1, Co)) if plan.big else None
                bw = (wct, tg.split_weight(W1.t()))
        if gathered is not None:
            Q, done = gathered
            torch.cuda.current_stream(dev).wait_event(done)
            Q.record_stream(torch.cuda.current_stream(dev))
        if ex is not None:
            Z, arg_min, arg_max, mean, var = _k1_fwd_windows(graph, P, ex, F, plan.mat, p_drop, seed, seed_dev)
            ex.join()
        else:
            Z, arg_min, arg_max, mean, var = _k1_fwd(graph, P, Q, R, keep, F, plan.mat, p_drop, seed, seed_dev)   # sorted rows
        out = torch.empty((n, Co), dtype=torch.float32, device=dev)
        if plan.big:
            tg.linear(Z, hi, lo, Co, tile_tab=plan.tile_tab, out=out, out_map=graph.row_map, add=XW,
                      name="gemm_post_grouped")                                                     # rows -> node order
        if plan.tail_idx is not None:
            ti = plan.tail_idx
            nodes = plan.tail_nodes
            Zt = Z.index_select(0, ti)
            Yt = torch.einsum("nsam,nmf->nsaf", plan.tail_coef, Zt.view(-1, Am, F)).reshape(Zt.shape[0], S * A * F)
            out.index_copy_(0, nodes, (Yt @ Wp[:, F:].t()) @ Wl.t() + XW.index_select(0, nodes))
        if KEEP_LAST is not None:
            KEEP_LAST.update(Z=Z, arg_min=arg_min, arg_max=arg_max, graph=graph)
        ctx.graph, ctx.sg, ctx.plan, ctx.cfg = graph, sg, plan, cfg
        ctx.has_R, ctx.has_b = R is not None, (bm is not None, bp is not None, bl is not None)
        need_sq = 4 in akinds or 5 in akinds
        Qsave = Q if (sg is not None and need_sq and ex is None) else None   # NCCL variant: gathered Q kept for the backward
        ctx.ex, ctx.ex_call = ex, (ex.fwd_calls if ex is not None else 0)
        ctx.bw, ctx.wprep = bw, wprep
        ctx.save_for_backward(x, PQX, Z, Wm, Wp, bp, Wl, Wc, R, keep, arg_min, arg_max, mean, var, Qsave,
                              seed_dev)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (x, PQX, Z, Wm, Wp, bp, Wl, Wc, R, keep, arg_min, arg_max, mean, var, Qsave,
         seed_dev) = ctx.saved_tensors
        graph, sg, plan = ctx.graph, ctx.sg, ctx.plan
        F, akinds, p_drop, seed = ctx.cfg[:4]
        ex = ctx.ex
        dev = d_out.device
        n = graph.n_dst
        A, S = len(akinds), plan.S
        Am, K = len(plan.mat), plan.K
        Fo, Co = Wp.shape[0], Wl.shape[0]
        P, Q = PQX[:, :F], PQX[:, F:2 * F]
        if Qsave is not None:
            Q = Qsave
        wprep = ctx.wprep
        Wx, Wy = Wp[:, :F], (Wp[:, F:] if wprep else Wp[:, F:].contiguous())
        d_out = d_out.contiguous()
        dO, dbc = tg.gather_rows_colsum(d_out, graph.row_map)                                       # sorted rows; [Co]
        # ---- dgrad of the grouped post transform (composed with lin), then K1's backward
        dZ = torch.empty((n, K), dtype=torch.float32, device=dev)
        bw = ctx.bw
        if plan.big:
            if bw is not None:
                hi, lo = bw[0]
            else:
                hi, lo = tg.split_weight(Wc.transpose(1, 2).contiguous().view(-1, Co))              # [B * K, Co]
            tg.linear(dO, hi, lo, K, tile_tab=plan.tile_tab_t, out=dZ, name="gemm_post_dgrad")
        if plan.tail_idx is not None:
            ti = plan.tail_idx
            dOt = dO.index_select(0, ti)
            dHt = dOt @ Wl
            dYt = (dHt @ Wy).view(-1, S, A, F)
            dZ.index_copy_(0, ti, torch.einsum("nsam,nsaf->nmf", plan.tail_coef, dYt).reshape(-1, K))
        dPQ = torch.empty((n, 2 * F), dtype=torch.float32, device=dev)
        need_R = ctx.has_R and ctx.needs_input_grad[7]
        pending = None
        if sg is None:
            dR = _k1_bwd(graph, P, Q, R, keep, F, plan.mat, p_drop, seed, dZ, arg_min, arg_max, mean, var,
                         dPQ[:, :F], dPQ[:, F:2 * F], need_R, seed_dev)
        elif ex is not None:
            # destination pass + transpose pass give this rank's PARTIAL dQ over all sources; its slices leave for their
            # owners on the copy engine, owner by owner, while the next owner's rows are summed; the owner adds the
            # slices in rank order after the weight-gradient GEMM below, which does not depend on dQ
            need_sq = 4 in akinds or 5 in akinds
            if need_sq and ctx.ex_call != ex.fwd_calls:
                raise RuntimeError("sharded MMAConv: the gathered Q of this forward was overwritten by a later forward of "
                                   "the same layer; run backward before the next forward of this layer")
            ex.begin_call()
            dR = None
            _k1_bwd_sharded(graph, P, ex, F, plan.mat, p_drop, seed, dZ, arg_min, arg_max, mean, var, dPQ[:, :F],
                            need_sq, seed_dev)
        else:
            # NCCL variant: partial dQ over ALL sources, then one reduce-scatter
            from .parallel import reduce_scatter_rows, _comm_stream
            dQ_part = torch.empty((graph.n_src, F), dtype=torch.float32, device=dev)
            dR = _k1_bwd(graph, P, Q, R, keep, F, plan.mat, p_drop, seed, dZ, arg_min, arg_max, mean, var,
                         dPQ[:, :F], dQ_part, need_R, seed_dev)
            if OVERLAP_EXCHANGE:
                comm, cur = _comm_stream(dev), torch.cuda.current_stream(dev)
                comm.wait_stream(cur)
                with torch.cuda.stream(comm):
                    dQ_loc = reduce_scatter_rows(dQ_part, sg.max_rows, n, sg.group)
                    done = torch.cuda.Event(); done.record(comm)
                dQ_part.record_stream(comm)
                pending = (dQ_loc, done)
            else:
                dPQ[:, F:2 * F] = reduce_scatter_rows(dQ_part, sg.max_rows, n, sg.group)
        del dZ
        # ---- wgrad of the grouped post transform
        dWl = torch.outer(dbc, bp) if bp is not None else torch.zeros_like(Wl)
        dWy = dWp = None
        if plan.big and wprep:
            part = tg.wgrad_partials(dO, Z, plan.slabs, plan.slabs.shape[0], name="gemm_post_wgrad")
            dWc = tg.reduce_slabs_segmented(part, plan.seg_ptr)                                     # [B, Co, K]
            # D = d(W_lin W_post) in W_post's column layout (its x-part, which needs the mask GEMM's weight gradient, is
            # added at the end); dW_lin += D W_post^T (long K: split-K slabs added in order), dW_post = W_lin^T D
            D = tg.compose_post_wgrad(plan.coef_big, plan.block_of_t, dWc, None, F, F)               # [Co, (S*A+1)*F]
            dWl = dWl + tg.small_gemm(D, Wp, trans_b=True, k_splits=max(1, D.shape[1] // 128))
            dWp = tg.small_gemm(Wl, D, trans_a=True)                                                # [Fo, (S*A+1)*F]
        elif plan.big:
            part = tg.wgrad_partials(dO, Z, plan.slabs, plan.slabs.shape[0], name="gemm_post_wgrad")
            dWc = tg.reduce_slabs_segmented(part, plan.seg_ptr)                                     # [B, Co, K]
            # W_c(d)[:, m] = sum_{s, a -> m} coef(d, s, a) W_lin W_{s,a}  ->  D_{s,a} = sum_d coef(d, s, a) dW_c(d)[:, m(a)];
            # dW_lin += sum_s D_s W_s^T,  dW_s = W_lin^T D_s
            D = torch.einsum("bsam,bcmf->scaf", plan.coef_big, dWc.view(-1, Co, Am, F)).reshape(S, Co, A * F)
            Ws = Wy.view(Fo, S, A * F).permute(1, 0, 2)                                             # [S, Fo, A*F]
            dWl = dWl + torch.matmul(D, Ws.transpose(1, 2)).sum(0)
            dWy = torch.matmul(Wl.t(), D).permute(1, 0, 2).reshape(Fo, S * A * F)
        if plan.tail_idx is not None:
            ti = plan.tail_idx
            Zt = Z.index_select(0, ti)
            Yt = torch.einsum("nsam,nmf->nsaf", plan.tail_coef, Zt.view(-1, Am, F)).reshape(Zt.shape[0], S * A * F)
            dWl = dWl + dOt.t() @ (Yt @ Wy.t())
            g = dHt.t() @ Yt
            if dWp is not None:
                dWp[:, F:] += g
            else:
                dWy = g if dWy is None else dWy + g
        del dO
        has_bm, has_bp, has_bl = ctx.has_b
        dbm = None
        split = ex is not None and bw is not None and len(bw) == 4
        dx = torch.empty((n, F), dtype=torch.float32, device=dev)
        if split:
            # everything of the mask GEMMs that does not need dQ, while the slices are still arriving
            tg.linear(dPQ[:, :F], bw[2][0], bw[2][1], F, A1=d_out, out=dx, name="gemm_mask_dgrad")
            dW_pe = tg.wgrad(dPQ[:, :F], x, G1=d_out, name="gemm_mask_wgrad")                       # rows [dP ; dOut]
            if has_bm:
                dbm = tg.colsum(dPQ[:, :F])
        if ex is not None:
            ex.sum_slices(dPQ[:, F:2 * F], n)
            ex.join()
        if pending is not None:
            dQ_loc, done = pending
            torch.cuda.current_stream(dev).wait_event(done)
            dPQ[:, F:2 * F] = dQ_loc
            dQ_loc.record_stream(torch.cuda.current_stream(dev))
        # ---- mask projection + composed x-part, all in node order: dx, dW1
        if bw is not None:
            w1hi, w1lo = bw[1]
        else:
            Wcx = Wl @ Wx
            W1 = torch.cat([Wm[:, :F], Wm[:, F:2 * F], Wcx], dim=0)
            w1hi, w1lo = tg.split_weight(W1.t())
        if split:
            tg.linear(dPQ[:, F:2 * F], bw[3][0], bw[3][1], F, out=dx, add=dx, name="gemm_mask_dgrad")   # dx += dQ W_j
            dW_q = tg.wgrad(dPQ[:, F:2 * F], x, name="gemm_mask_wgrad")
            dW1 = torch.cat([dW_pe[:F], dW_q, dW_pe[F:]], dim=0)
        elif (2 * F) % 128 == 0:
            tg.linear(dPQ, w1hi, w1lo, F, A1=d_out, out=dx, name="gemm_mask_dgrad")
            dW1 = tg.wgrad(dPQ, x, G1=d_out, name="gemm_mask_wgrad")
        else:
            dPQX = torch.cat([dPQ, d_out], dim=1)
            tg.linear(dPQX, w1hi, w1lo, F, out=dx, name="gemm_mask_dgrad")
            dW1 = tg.wgrad(dPQX, x, name="gemm_mask_wgrad")
        dWm = torch.zeros_like(Wm)
        dWm[:, :F] = dW1[:F]
        dWm[:, F:2 * F] = dW1[F:2 * F]
        dWcx = dW1[2 * F:]                                                                          # [Co, F]
        if wprep:
            dWl = dWl + tg.small_gemm(dWcx, Wx, trans_b=True)
            dWpx = tg.small_gemm(Wl, dWcx, trans_a=True)                                            # [Fo, F]
            if dWp is not None:
                dWp[:, :F] = dWpx
            else:
                dWp = torch.cat([dWpx, dWy if dWy is not None else torch.zeros_like(Wy)], dim=1)
        else:
            dWl = dWl + dWcx @ Wx.t()
            dWp = torch.cat([Wl.t() @ dWcx, dWy if dWy is not None else torch.zeros_like(Wy)], dim=1)
        if has_bm and dbm is None:
            dbm = tg.colsum(dPQ[:, :F]) if F % 4 == 0 else dPQ[:, :F].sum(0)
        dbp = dbc @ Wl if has_bp else None
        return dx, dWm, dbm, dWp, dbp, dWl, (dbc if has_bl else None), dR, None, None, None, None


def supported(F_in: int, F_out: int, out_channels: int) -> bool:
    """Shapes the TMA-fed GEMMs accept (16-byte aligned rows); anything else uses the torch.mm path."""
    return F_in % 4 == 0 and F_out % 4 == 0 and out_channels % 4 == 0


def fused_mmaconv(x: Tensor, graph, *, W_mask: Tensor, b_mask: Optional[Tensor], W_post: Tensor,
                  b_post: Optional[Tensor], W_lin: Tensor, b_lin: Optional[Tensor], R: Optional[Tensor],
                  keep: Optional[Tensor], aggregators: Sequence[str], scalers: Sequence[str],
                  avg_deg: Dict[str, float], p_drop: float, seed: int, min_rows: Optional[int],
                  seed_dev: Optional[Tensor] = None, n_windows: int = 2, owner: int = 0) -> Tensor:
    """x [N, F] (node order) -> MMAConv output [N, out] for towers == 1, pre_layers == post_layers == 1.
    W_mask [F, 2F or 3F] (the live mask Linear, Q2; only the first 2F columns are used here, the edge
    part arrives as R), W_post [Fo, (S*A+1)*F], W_lin [out, Fo]."""
    local = graph if isinstance(graph, Graph) else graph.local
    if local.row_map is None or local.buckets is None:
        raise RuntimeError("fused_mmaconv needs a Graph built with sort_rows=True (degree-sorted CSR rows)")
    if local is not graph and (R is not None or keep is not None):
        raise RuntimeError("the sharded fused path supports neither edge features nor explicit keep masks")
    F = x.shape[1]
    for a in aggregators:                       # aggregate(), mma_conv.py:164-177: exact names only
        if a not in _lib.AGGR_KINDS:
            raise ValueError(f'Unknown aggregator "{a}".')
    if len(aggregators) > _lib.MAX_AGGR or len(scalers) > _lib.MAX_SCALER:
        raise _lib.MMAError("more than 8 aggregators or scalers in one call")
    akinds = tuple(_lib.AGGR_KINDS[a] for a in aggregators)
    plan = post_plan(local, scalers, avg_deg, min_rows, W_lin.shape[0], akinds, F)
    return _FusedMMAConv.apply(x, W_mask, b_mask, W_post, b_post, W_lin, b_lin, R, keep, graph, plan,
                               (F, akinds, float(p_drop), int(seed), seed_dev, int(n_windows), int(owner)))
