"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (mma_b200/).

CPU restatement (torch CPU tensors, fp32) of the reference's multi-mask
aggregation hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module, and only as the
checker / the timed CPU baseline -- never as a fallback for the CUDA path.

PARITY STATUS: the reference has no tests, golden vectors or KATs for this
path => "parity unpinned" by the reference's own
fixtures.  It is pinned instead against OUTPUTS OF THE REFERENCE ITSELF: the
verbatim files run in the build container under oracle/ref_shims.py, and the
vectors they produce are committed under tests/golden/ by
oracle/make_golden.py.  tests/test_oracle_golden.py checks every function here
against those vectors; tests/test_oracle_vs_reference.py (skipped when
/root/reference is absent) re-runs the verbatim code side by side.

Third-party arithmetic on the path that is NOT in /root/reference:
torch_scatter.scatter (unpinned; torch-scatter 2.0.8/2.0.9 implied by
README.md:34-38).  `scatter` below restates its published CPU algorithm; the
sequential ground truth for tie-breaking is oracle/scatter_seq.c.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

FLT_MAX = torch.finfo(torch.float32).max


# ==========================================================================
# torch_scatter.scatter  (called at graph_regression/mma_conv.py:166,168,169)
# ==========================================================================
def _minmax_first(flat: Tensor, index: Tensor, n: int, is_max: bool) -> Tuple[Tensor, Tensor]:
    """min/max over segments with torch_scatter CPU semantics: init +-FLT_MAX,
    first strict improvement in edge order wins, empty -> (0, arg=E).
    Vectorised; bit-identical to oracle/scatter_seq.c (tests check that)."""
    E, K = flat.shape
    init = -FLT_MAX if is_max else FLT_MAX
    idx = index.view(E, 1).expand(E, K)
    flat = torch.where(torch.isnan(flat), torch.full_like(flat, init), flat)   # NaN never wins a strict compare
    ext = torch.full((n, K), init, dtype=flat.dtype).scatter_reduce_(
        0, idx, flat, "amax" if is_max else "amin", include_self=True)
    eid = torch.arange(E, dtype=torch.int64).view(E, 1).expand(E, K)
    cand = torch.where(flat == ext.index_select(0, index), eid, torch.full_like(eid, E))
    arg = torch.full((n, K), E, dtype=torch.int64).scatter_reduce_(
        0, idx, cand, "amin", include_self=True)
    arg = torch.where(ext == init, torch.full_like(arg, E), arg)  # never improved on init
    if E > 0:
        val = flat.gather(0, arg.clamp(max=E - 1))   # value of the FIRST occurrence (keeps -0.0/+0.0)
    else:
        val = torch.zeros((n, K), dtype=flat.dtype)
    out = torch.where(arg < E, val, torch.zeros_like(val))
    return out, arg


class _ScatterMinMax(torch.autograd.Function):
    """Backward routes the gradient to the arg edge only (torch_scatter's
    `scatter_min/max` backward: grad.gather by arg; no tie splitting)."""

    @staticmethod
    def forward(ctx, src, index, n, is_max):
        E = src.shape[0]
        out, arg = _minmax_first(src.reshape(E, int(math.prod(src.shape[1:]))), index, n, is_max)
        ctx.save_for_backward(arg)
        ctx.src_shape = src.shape
        shape = (n,) + tuple(src.shape[1:])
        arg = arg.reshape(shape)
        ctx.mark_non_differentiable(arg)
        return out.reshape(shape), arg

    @staticmethod
    def backward(ctx, g_out, _g_arg):
        (arg,) = ctx.saved_tensors
        E = ctx.src_shape[0]
        K = arg.shape[1]
        g = torch.zeros((E + 1, K), dtype=g_out.dtype)
        g.scatter_(0, arg, g_out.reshape(arg.shape[0], K))
        return g[:E].reshape(ctx.src_shape), None, None, None


def scatter_with_arg(src: Tensor, index: Tensor, dim_size: int, reduce: str):
    """(out, arg) for reduce in {'min','max'}; arg has E for empty segments."""
    return _ScatterMinMax.apply(src, index, dim_size, reduce == "max")


def scatter(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
            dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    """torch_scatter.scatter restricted to the reference's call shape
    (dim=0, out=None, 1-D index broadcast over trailing dims)."""
    assert dim == 0 and out is None and index.dim() == 1
    n = int(index.max()) + 1 if dim_size is None else int(dim_size)
    trailing = tuple(src.shape[1:])
    if reduce in ("sum", "add"):
        return torch.zeros((n,) + trailing, dtype=src.dtype).index_add_(0, index, src)
    if reduce == "mean":
        s = torch.zeros((n,) + trailing, dtype=src.dtype).index_add_(0, index, src)
        cnt = torch.zeros((n,), dtype=src.dtype).index_add_(
            0, index, torch.ones((src.shape[0],), dtype=src.dtype))
        cnt = cnt.clamp(min=1).view((n,) + (1,) * len(trailing))
        return s / cnt                                    # true division
    if reduce in ("min", "max"):
        return scatter_with_arg(src, index, n, reduce)[0]
    # torch_scatter raises ValueError for anything else ('min2', 'sum3', ...: Q6)
    raise ValueError(f"unknown reduce '{reduce}'")


# ==========================================================================
# graph_regression/mma_conv.py  (MultiMaskConv)
# ==========================================================================
def avg_deg_from_hist(deg: Tensor) -> Dict[str, float]:
    """mma_conv.py:73-78 -- statistics of the *histogram tensor itself* (Q5)."""
    deg = deg.to(torch.float)
    return {"lin": deg.mean().item(),
            "log": (deg + 1).log().mean().item(),
            "exp": deg.exp().mean().item()}


@dataclass
class MMAConvWeights:
    """The live parameters of one reference MMAConv (Q1/Q2: only the mask
    linears of aggregators[-1] are ever applied)."""
    in_channels: int
    out_channels: int
    aggregators: List[str]
    scalers: List[str]
    avg_deg: Dict[str, float]
    towers: int
    F_in: int
    F_out: int
    divide_input: bool
    enc: Optional[Tuple[Tensor, Tensor]]                 # edge_encoder (w [F_in,edge_dim], b)
    pre: List[List[Tuple[Tensor, Tensor]]]               # [tower][layer] -> (w, b) of aggregators[-1]
    post: List[List[Tuple[Tensor, Tensor]]]              # [tower][layer] -> (w, b)
    lin: Tuple[Tensor, Tensor]
    dropout: float = 0.5

    def tensors(self) -> List[Tensor]:
        out = []
        if self.enc is not None:
            out += list(self.enc)
        for t in self.pre + self.post:
            for w, b in t:
                out += [w, b]
        out += list(self.lin)
        return out


def weights_from_module(conv, clone: bool = True) -> MMAConvWeights:
    """Extracts the live tensors from a reference MMAConv *or* from the drop-in
    mma_b200 MMAConv (same attribute structure: mma_conv.py:84-105,
    mask_aggr.py:44-51)."""
    cp = (lambda t: t.detach().clone().cpu()) if clone else (lambda t: t)
    a_star = conv.aggregators[-1]
    pre = []
    for seq in conv.pre_nns[a_star]:
        layers = []
        for m in seq:
            if hasattr(m, "aggregation_layers"):
                lin = m.aggregation_layers[m.aggregation]
                layers.append((cp(lin.weight), cp(lin.bias)))
        pre.append(layers)
    post = []
    for seq in conv.post_nns:
        post.append([(cp(m.weight), cp(m.bias)) for m in seq if hasattr(m, "weight")])
    enc = None
    if conv.edge_dim is not None:
        enc = (cp(conv.edge_encoder.weight), cp(conv.edge_encoder.bias))
    return MMAConvWeights(conv.in_channels, conv.out_channels, list(conv.aggregators),
                          list(conv.scalers), dict(conv.avg_deg), conv.towers, conv.F_in,
                          conv.F_out, conv.divide_input, enc, pre, post,
                          (cp(conv.lin.weight), cp(conv.lin.bias)), conv.dropout)


def mmaconv_aggregate(inputs: Tensor, index: Tensor, dim_size: int, aggregators: Sequence[str],
                      scalers: Sequence[str], avg_deg: Dict[str, float],
                      return_args: bool = False):
    """mma_conv.py:159-196 `MMAConv.aggregate`, op for op."""
    outs, args = [], {}
    for aggregator in aggregators:
        if aggregator.startswith(("sum", "mean", "min", "max")):
            if aggregator in ("min", "max"):
                out, arg = scatter_with_arg(inputs, index, dim_size, aggregator)
                args[aggregator] = arg
            else:
                out = scatter(inputs, index, 0, None, dim_size, reduce=aggregator)   # :165-166
        elif aggregator in ("var", "std"):
            mean = scatter(inputs, index, 0, None, dim_size, reduce="mean")          # :168
            mean_squares = scatter(inputs * inputs, index, 0, None, dim_size, reduce="mean")
            out = mean_squares - mean * mean                                         # :170
            if aggregator == "std":
                out = torch.sqrt(torch.relu(out) + 1e-5)                             # :172
        else:
            raise ValueError(f'Unknown aggregator "{aggregator}".')
        outs.append(out)
    out = torch.cat(outs, dim=-1)                                                    # :176

    deg = torch.zeros((dim_size,), dtype=inputs.dtype).scatter_add_(
        0, index, torch.ones((index.numel(),), dtype=inputs.dtype))                  # :178
    deg = deg.clamp_(1).view((-1,) + (1,) * (inputs.dim() - 1))                      # :179

    outs = []
    for scaler in scalers:                      # cumulative re-assignment (Q4), :181-195
        if scaler == "identity":
            pass
        elif scaler == "amplification":
            out = out * (torch.log(deg + 1) / avg_deg["log"])
        elif scaler == "attenuation":
            out = out * (avg_deg["log"] / torch.log(deg + 1))
        elif scaler == "linear":
            out = out * (deg / avg_deg["lin"])
        elif scaler == "inverse_linear":
            out = out * (avg_deg["lin"] / deg)
        else:
            raise ValueError(f'Unknown scaler "{scaler}".')
        outs.append(out)
    res = torch.cat(outs, dim=-1)                                                    # :196
    return (res, args) if return_args else res


def mmaconv_message(w: MMAConvWeights, x_i: Tensor, x_j: Tensor, edge_attr: Optional[Tensor],
                    keep: Optional[Tensor], strict: bool = True) -> Tensor:
    """mma_conv.py:138-157.  `keep` [E,T,F_in] is the dropout keep-scale tensor
    (0 or 1/(1-p)); None draws torch's own Bernoulli like F.dropout(hs, 0.5)."""
    T, F_in = w.towers, w.F_in
    if edge_attr is not None:
        e = F.linear(edge_attr, w.enc[0], w.enc[1])                                   # :143
        e = e.view(-1, 1, F_in).repeat(1, T, 1)                                      # :144-145
        h = torch.cat([x_i, x_j, e], dim=-1)                                         # :146
    else:
        h = torch.cat([x_i, x_j], dim=-1)                                            # :148
    for aggregator in w.aggregators:                                                 # :150-154
        # strict=False: BASELINE config 4 names 'std', which upstream's message() rejects (Q6)
        if strict and not aggregator.startswith(("sum", "mean", "min", "max")):
            raise ValueError(f'Unknown aggregator "{aggregator}".')
    hs = []
    for t in range(T):                       # only aggregators[-1]'s mask linears survive (Q2)
        v = h[:, t]
        for li, (wt, bt) in enumerate(w.pre[t]):
            if li > 0:
                v = torch.relu(v)                                                    # :93-95
            v = F.linear(v, wt, bt)                                                  # mask_aggr.py:68
        hs.append(v)
    hs = torch.stack(hs, dim=1)                                                      # :156
    if keep is None:
        return F.dropout(hs, w.dropout)      # training flag never passed -> always on (Q3), :157
    return hs * keep


def mmaconv_forward(w: MMAConvWeights, x: Tensor, edge_index: Tensor,
                    edge_attr: Optional[Tensor] = None, keep: Optional[Tensor] = None,
                    strict: bool = True) -> Tensor:
    """mma_conv.py:121-136 `MMAConv.forward` through PyG's propagate
    (x_j = x[edge_index[0]], x_i = x[edge_index[1]], index = edge_index[1])."""
    T, F_in = w.towers, w.F_in
    if w.divide_input:
        xt = x.view(-1, T, F_in)                                                     # :126
    else:
        xt = x.view(-1, 1, F_in).repeat(1, T, 1)                                     # :128
    x_j = xt.index_select(0, edge_index[0])
    x_i = xt.index_select(0, edge_index[1])
    msg = mmaconv_message(w, x_i, x_j, edge_attr, keep, strict)
    out = mmaconv_aggregate(msg, edge_index[1], xt.size(0), w.aggregators, w.scalers, w.avg_deg)
    out = torch.cat([xt, out], dim=-1)                                               # :132
    outs = []
    for t in range(T):                                                               # :133
        v = out[:, t]
        for li, (wt, bt) in enumerate(w.post[t]):
            if li > 0:
                v = torch.relu(v)
            v = F.linear(v, wt, bt)
        outs.append(v)
    out = torch.cat(outs, dim=1)                                                     # :134
    return F.linear(out, w.lin[0], w.lin[1])                                         # :136


def mmconv_fused_op(P: Optional[Tensor], Q: Optional[Tensor], R: Optional[Tensor],
                    keep: Optional[Tensor], src: Tensor, dst: Tensor, n: int,
                    aggregators: Sequence[str], scalers: Sequence[str],
                    avg_deg: Dict[str, float], return_args: bool = False):
    """The fused op the CUDA kernel implements, in the kernel's arithmetic
    order: m_e = ((P[dst] + Q[src]) + R[e]) * keep[e]  (SURVEY.md A.1 step 2:
    the mask linear over cat([x_i,x_j,e]) is separable), then
    mma_conv.py:159-196.  Used to check min/max values and args bit-for-bit."""
    m = None
    if P is not None:
        m = P.index_select(0, dst)
    if Q is not None:
        q = Q.index_select(0, src)
        m = q if m is None else m + q
    if R is not None:
        m = R if m is None else m + R
    if keep is not None:
        m = m * keep
    return mmaconv_aggregate(m, dst, n, aggregators, scalers, avg_deg, return_args=return_args)


# ==========================================================================
# node_classification/layers.py + scalers.py  (masked multi-aggregator layer)
# ==========================================================================
NC_WORKING = ("sum", "sum2", "sum3", "sum4", "mean", "mean2", "mean3", "mean4",
              "max", "max2", "max3", "max4", "min", "min2", "min3", "min4",
              "softmax", "softmin")
NC_BROKEN = ("std", "normalized_mean", "moment_3")          # layers.py:731-851 (a16)
# aggregators whose mask stays the raw logit under activation == "new_sigmoid"
# (layers.py:381-385, 445-449, 555-559, 668-672, 708-712: result discarded, Q8)
NC_RAW_UNDER_NEW_SIGMOID = ("mean3", "max", "min", "softmax", "softmin")


def nc_family(name: str) -> str:
    for fam in ("softmax", "softmin", "sum", "mean", "max", "min"):
        if name.startswith(fam):
            return fam
    raise KeyError(name)


def add_all_to_csr(add_all) -> Tuple[Tensor, Tensor]:
    """utils.py:98-100: add_all[i] = neighbour ids of node i (row i of adj)."""
    deg = torch.tensor([len(r) for r in add_all], dtype=torch.int64)
    rowptr = torch.zeros(len(add_all) + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(deg, 0)
    if int(rowptr[-1]) == 0:
        return rowptr, torch.zeros(0, dtype=torch.int64)
    import numpy as np
    col = torch.from_numpy(np.concatenate([np.asarray(r, dtype=np.int64) for r in add_all]))
    return rowptr, col


def nc_aggregate(x: Tensor, rowptr: Tensor, col: Tensor, M: Tensor, name: str,
                 activation: str, p: float, keep: Optional[Tensor] = None) -> Tensor:
    """One `learnable_<name>` (layers.py:201-728), vectorised over edges instead
    of the per-node Python loop: for edge (i <- j)
      logits = [x_i || x_j] @ M           (:209-216)
      mask   = sigmoid(logits)  or raw    (:217 / Q8)
      mask   = dropout(mask, p)  ALWAYS   (:219, Q3)   -> `keep` injects it
      S_i    = sum_j mask * x_j           (:221)
    then sum: x_i+S_i; mean: (x_i+S_i)/D_i (:328-329); max/min: elementwise
    max/min(x_i, S_i) (:452,:562); softmax/softmin over a size-1 dim == S_i
    (:676-682; exp overflow -> NaN reproduced)."""
    n, Fd = x.shape
    deg = rowptr[1:] - rowptr[:-1]
    dst = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg)
    x_i = x.index_select(0, dst)
    x_j = x.index_select(0, col)
    logits = torch.mm(torch.cat([x_i, x_j], 1), M)
    raw = activation == "new_sigmoid" and name in NC_RAW_UNDER_NEW_SIGMOID
    mask = logits if raw else torch.sigmoid(logits)
    mask = F.dropout(mask, p) if keep is None else mask * keep
    S = torch.zeros((n, Fd), dtype=x.dtype).index_add_(0, dst, mask * x_j)
    fam = nc_family(name)
    if fam == "sum":
        return x + S
    if fam == "mean":
        return torch.div(x + S, deg.to(x.dtype).view(-1, 1))    # D_i = 0 -> inf/nan like the reference (Q9)
    if fam == "max":
        return torch.max(x, S)
    if fam == "min":
        return torch.min(x, S)
    X = S.unsqueeze(0)                                          # [1,N,F]: softmax over a size-1 dim
    X_exp = torch.exp(X if fam == "softmax" else -X)
    X_sum = torch.sum(X_exp, dim=0, keepdim=True)
    return torch.sum(torch.mul(torch.div(X_exp, X_sum), X), dim=0)


def nc_scale_factors(n: int, dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """scalers.py:26-62 as called from layers.py:856 with the sparse `adj` in
    the add_all slot (Q7): every 'degree' is len(adj[i]) == N."""
    all_degrees = torch.tensor([n] * n)                          # int64, scalers.py:28-29
    lg = torch.log(all_degrees + 1)                              # float32
    avg = torch.mean(lg)                                         # scalers.py:10-14
    return (lg / avg).unsqueeze(-1).to(dtype), (avg / lg).unsqueeze(-1).to(dtype)


def nc_forward(x: Tensor, adj: Tensor, rowptr: Tensor, col: Tensor, masks: Dict[str, Tensor],
               weight: Tensor, bias: Optional[Tensor], names: Sequence[str], activation: str,
               p: float, keeps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """layers.py:853-867 `MMA.forward`."""
    names = list(dict.fromkeys(names))       # AGGREGATORS is a dict: duplicates collapse (:108-112)
    for nm in names:
        if nm in NC_BROKEN:
            raise RuntimeError(f"reference aggregator '{nm}' is broken upstream (layers.py:731-851)")
        if nm not in NC_WORKING:
            raise KeyError(nm)                                   # layers.py:110
    A = len(names)
    n = x.shape[0]
    m = torch.cat([nc_aggregate(x, rowptr, col, masks[nm], nm, activation, p,
                                None if keeps is None else keeps[nm]) for nm in names], dim=0)
    if A > 4:    # scalers.py:33-40 tile the scale only for A in 1..4 -> torch.mul shape error
        raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (%d)"
                           % (n, A * n))
    amp, att = nc_scale_factors(n, x.dtype)
    amp, att = amp.repeat(A, 1), att.repeat(A, 1)
    m = torch.cat([m, torch.mul(amp, m), torch.mul(att, m)], dim=1)                 # :856
    w3 = torch.cat([weight, weight, weight], dim=0)                                  # :858
    support = torch.mm(m, w3)                                                        # :860
    adj_t = torch.cat((adj,) * A, 1)                                                 # :861
    out = torch.spmm(adj_t, support)                                                 # :862
    return out + bias if bias is not None else out                                   # :864-867


def gcn_forward(x: Tensor, adj: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """layers.py:38-45 `GraphConvolution.forward`."""
    out = torch.spmm(adj, torch.mm(x, weight))
    return out + bias if bias is not None else out


def csr_to_sparse_adj(rowptr: Tensor, col: Tensor, n: int) -> Tensor:
    """utils.py:139-146: binary COO adjacency, unnormalised, no self loops added (Q9)."""
    deg = rowptr[1:] - rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg)
    return torch.sparse_coo_tensor(torch.stack([row, col]), torch.ones(col.numel()), (n, n)).coalesce()


# ==========================================================================
# synthetic inputs of the BASELINE.json configs
# ==========================================================================
def zinc_like_batch(num_graphs: int = 128, seed: int = 42):
    """c2: molecules as random spanning tree + ring-closure edges, max degree 4,
    both directions.  Returns (edge_index [2,E] int64, batch [N] int64)."""
    g = torch.Generator().manual_seed(seed)
    srcs, dsts, batch = [], [], []
    base = 0
    for gi in range(num_graphs):
        n = int(torch.clamp(torch.round(torch.randn((), generator=g) * 4.5 + 23.2), 9, 37))
        degc = [0] * n
        und = set()
        for v in range(1, n):
            for _ in range(64):
                u = int(torch.randint(0, v, (), generator=g))
                if degc[u] < 4:
                    break
            else:
                u = min(range(v), key=lambda k: degc[k])
            und.add((u, v)); degc[u] += 1; degc[v] += 1
        extra = max(0, int(round(1.07 * n)) - (n - 1))
        tries = 0
        while extra > 0 and tries < 200:
            tries += 1
            u = int(torch.randint(0, n, (), generator=g)); v = int(torch.randint(0, n, (), generator=g))
            if u == v:
                continue
            a, b = min(u, v), max(u, v)
            if (a, b) in und or degc[a] >= 4 or degc[b] >= 4:
                continue
            und.add((a, b)); degc[a] += 1; degc[b] += 1; extra -= 1
        for (u, v) in sorted(und):
            srcs += [base + u, base + v]; dsts += [base + v, base + u]
        batch += [gi] * n
        base += n
    return torch.tensor([srcs, dsts], dtype=torch.int64), torch.tensor(batch, dtype=torch.int64)


def degree_histogram(edge_index: Tensor, n: int) -> Tensor:
    """graph_regression/mma.py:57-60: bincount of in-degrees."""
    d = torch.zeros(n, dtype=torch.int64).scatter_add_(0, edge_index[1], torch.ones_like(edge_index[1]))
    return torch.bincount(d)
