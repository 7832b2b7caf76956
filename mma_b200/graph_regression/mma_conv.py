"""Drop-in for /root/reference/graph_regression/mma_conv.py (`MMAConv`, "MultiMaskConv").

Constructor, forward signature, attributes and quirks are the reference's
(mma_conv.py:47-51,121-122); the compute is the fused sm_100a kernel K1 instead of
PyG propagate + T edge-level Linears + A torch_scatter passes:

  reference (per layer call)                         here
  -------------------------------------------------  ------------------------------------------
  x_i, x_j = x[dst], x[src]         (:130)           never materialised (gathered in-kernel)
  h = cat([x_i, x_j, enc(e)])       (:146)           never materialised
  hs = Linear_{a*}(h) per tower     (:150-156, Q2)   P = X W_i^T + b, Q = X W_j^T   (node-level GEMMs)
                                                     R = enc(e) W_e^T               (edge-level GEMM)
  dropout(hs, 0.5) always           (:157, Q3)       in-kernel Philox, p = self.dropout
  A scatters, degree, S scalers     (:159-196)       one pass over the destination CSR
  autograd with atomic index_add_   (mma.py:157)     deterministic dst pass + transpose-CSR pass

Quirks kept: only aggregators[-1]'s mask linears are applied (Q2); pre_nns /
aggregation_layers are plain dicts so their weights are unregistered (Q1); dropout ignores
self.training (Q3); scalers compound (Q4); avg_deg is computed from the histogram tensor
itself (Q5); aggregator names not starting with sum/mean/min/max raise ValueError in
forward (Q6) and names torch_scatter does not know ('min2', ...) raise ValueError in
aggregate.

Extensions (keyword-only, defaults keep reference behaviour):
  strict_reference=True   False lets 'var'/'std' through forward (unreachable upstream, Q6;
                          needed for BASELINE config 4 "all aggregators").
`edge_index` may also be a prebuilt `mma_b200.graph.Graph`.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.nn import ModuleList, ReLU, Sequential

from .. import _lib
from .. import functional as MF
from .. import fused_layer
from ..graph import Graph, cached_graph
from ..parallel import ShardedGraph, sharded_mmconv_aggregate
from ..linear import Linear, dense_linear, reset
from .mask_aggr import MaskAggregateLinear

_UID = itertools.count(1)


class MMAConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, aggregators: List[str], scalers: List[str],
                 deg: Tensor, edge_dim: Optional[int] = None, towers: int = 1, pre_layers: int = 1,
                 post_layers: int = 1, mask: bool = True, divide_input: bool = False, **kwargs):
        self.strict_reference = bool(kwargs.pop("strict_reference", True))
        kwargs.setdefault("aggr", None)
        super().__init__()
        self.aggr = kwargs.get("aggr")
        self.node_dim = 0                       # MessagePassing(node_dim=0), mma_conv.py:54

        if divide_input:
            assert in_channels % towers == 0
        assert out_channels % towers == 0

        self.in_channels = in_channels
        self.out_channels = out_channels
        self.aggregators = aggregators
        self.scalers = scalers
        self.edge_dim = edge_dim
        self.towers = towers
        self.divide_input = divide_input
        self.dropout = 0.5                      # mma_conv.py:67
        self.mask = mask
        self.pre_layers, self.post_layers = pre_layers, post_layers

        self.F_in = in_channels // towers if divide_input else in_channels
        self.F_out = self.out_channels // towers

        deg = deg.to(torch.float)               # statistics of the histogram itself (Q5), :72-78
        self.avg_deg: Dict[str, float] = {
            "lin": deg.mean().item(),
            "log": (deg + 1).log().mean().item(),
            "exp": deg.exp().mean().item(),
        }

        if self.edge_dim is not None:
            self.edge_encoder = Linear(edge_dim, self.F_in)

        self.pre_nns = {}                       # plain dict: unregistered (Q1), :84-86
        for aggr in aggregators:
            self.pre_nns[aggr] = ModuleList()

        self.post_nns = ModuleList()
        for _ in range(towers):
            for aggr in aggregators:
                modules = [MaskAggregateLinear((3 if edge_dim else 2) * self.F_in, self.F_in, aggregators,
                                               aggr, mask=self.mask)]
                for _ in range(pre_layers - 1):
                    modules += [ReLU()]
                    modules += [MaskAggregateLinear(self.F_in, self.F_in, aggregators, aggr, mask=self.mask)]
                self.pre_nns[aggr].append(Sequential(*modules))
            in_ch = (len(aggregators) * len(scalers) + 1) * self.F_in
            modules = [Linear(in_ch, self.F_out)]
            for _ in range(post_layers - 1):
                modules += [ReLU()]
                modules += [Linear(self.F_out, self.F_out)]
            self.post_nns.append(Sequential(*modules))

        self.lin = Linear(out_channels, out_channels)

        self.use_tensor_cores = True         # towers == 1: whole layer as one autograd node over tcgen05 GEMMs + K1
        self.fold_scalers = True             # towers == 1: fold the scalers into the post weight (see _forward_folded)
        self.fold_min_rows = None            # degree ranges smaller than this use the literal formula (None: auto --
                                             # no tail at all when the graph has at most 256 distinct in-degrees)
        self.comm_slices = 1                 # feature windows of the sharded exchange (1: full-width K1; narrow windows cost more in K1 than they hide)
        self.global_max_deg = None           # sharded runs: global max in-degree (else all-reduced per call)
        self.device_seed = False             # True: the dropout seed lives in a device tensor that is advanced ON the
                                             # device at every call, so a step captured in a CUDA graph draws a fresh
                                             # mask on every replay (fused path only; `last_seed` is then not updated)
        self._seed_dev: Optional[Tensor] = None
        self._uid = next(_UID)
        self._calls = 0
        self._inject_keep: Optional[Tensor] = None      # test hook: explicit keep-scale [E,T,F_in]
        self.last_seed: Optional[int] = None
        self.reset_parameters()

    def reset_parameters(self):
        if self.edge_dim is not None:
            self.edge_encoder.reset_parameters()
        for aggr in self.pre_nns:               # iterates the KEYS (strings): a no-op upstream, :113-115
            for nn in aggr:
                reset(nn)
        for nn in self.post_nns:
            reset(nn)
        self.lin.reset_parameters()

    # ------------------------------------------------------------------ helpers
    def mask_parameters(self) -> List[torch.nn.Parameter]:
        """The live (unregistered, Q1) mask-projection weights, e.g. to hand to an optimizer."""
        out = []
        if self.mask == "no_linear":
            return out
        for seq in self.pre_nns[self.aggregators[-1]]:
            for m in seq:
                if isinstance(m, MaskAggregateLinear):
                    lin = m.live()
                    out += [lin.weight] + ([lin.bias] if lin.bias is not None else [])
        return out

    def _next_seed(self) -> int:
        self._calls += 1
        s = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (self._uid * 1000003 + self._calls)) & 0xFFFFFFFFFFFFFFFF
        self.last_seed = s
        return s

    def _next_seed_dev(self, dev) -> Tensor:
        """Device-resident seed, advanced by a device-side add (capturable; create it before the capture)."""
        if self._seed_dev is None or self._seed_dev.device != dev:
            s = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (self._uid * 1000003)) & 0x7FFFFFFFFFFFFFFF
            self._seed_dev = torch.tensor([s], dtype=torch.int64, device=dev)
        self._seed_dev.add_(0x9E3779B97F4A7C15 - (1 << 64))     # golden-ratio increment, wraps modulo 2^64
        return self._seed_dev

    def _check_names(self):
        for aggregator in self.aggregators:     # message(), :150-154
            ok = ("sum", "mean", "min", "max") if self.strict_reference else ("sum", "mean", "min", "max", "var", "std")
            if not aggregator.startswith(ok):
                raise ValueError(f'Unknown aggregator "{aggregator}".')

    def _graph(self, edge_index, n: int, sort_rows: bool = False):
        if isinstance(edge_index, (Graph, ShardedGraph)):
            return edge_index
        if not edge_index.is_cuda:
            raise RuntimeError("mma_b200.MMAConv needs CUDA tensors (no CPU fallback)")
        return cached_graph(edge_index, n, sort_rows=sort_rows)

    # ------------------------------------------------------------------ forward
    def forward(self, x: Tensor, edge_index, edge_attr: Optional[Tensor] = None) -> Tensor:
        return self._forward_impl(x, edge_index, edge_attr, None)

    def forward_affine_relu(self, x: Tensor, edge_index, edge_attr: Optional[Tensor], scale: Tensor,
                            shift: Tensor) -> Tensor:
        """relu(forward(x, ...) * scale + shift) for per-channel scale / shift [out_channels]: BatchNorm in eval mode
        followed by ReLU (graph_regression/mma.py:120-121) folded into the layer's `lin` -- scaled weight rows, shifted
        bias, ReLU in the GEMM epilogue -- on the general (towers > 1) path; the other paths apply it after the layer."""
        return self._forward_impl(x, edge_index, edge_attr, (scale, shift))

    def _forward_impl(self, x: Tensor, edge_index, edge_attr: Optional[Tensor], epilogue) -> Tensor:
        T, F_in = self.towers, self.F_in
        post = (lambda y: y) if epilogue is None else (lambda y: F.relu(y * epilogue[0] + epilogue[1]))
        self._check_names()                     # ValueError first, whatever the device (message(), :150-154)
        if self.divide_input:
            xt = x.view(-1, T, F_in)
        else:
            xt = x.view(-1, 1, F_in)            # towers share x; the repeat (:128) is never materialised
        n = xt.size(0)
        if T == 1 and self.fold_scalers and self.pre_layers == 1 and self.mask != "no_linear":
            fused_ok = (self.use_tensor_cores and self.post_layers == 1 and
                        fused_layer.supported(F_in, self.F_out, self.out_channels))
            graph = self._graph(edge_index, n, sort_rows=True)
            if fused_ok and isinstance(graph, Graph) and graph.row_map is not None:
                return post(self._forward_fused(x.view(n, F_in), graph, edge_attr))
            if (fused_ok and isinstance(graph, ShardedGraph) and edge_attr is None and self._inject_keep is None
                    and graph.local is not None):
                return post(self._forward_fused(x.view(n, F_in), graph, None))
            local = graph.local if isinstance(graph, ShardedGraph) else graph
            if local.buckets is not None and not (isinstance(graph, ShardedGraph) and edge_attr is not None):
                return post(self._forward_folded(xt[:, 0], graph, edge_attr))
            edge_index = graph
        out = self.propagate(edge_index, x=xt, edge_attr=edge_attr, size=None)      # [N,T,S*A*F_in]

        # post_nns over cat([x, out]) (:132-133) without the cat: split the first Linear's weight
        K = out.size(-1)
        first = [seq[0] for seq in self.post_nns]
        if out.size(-1) + F_in != first[0].weight.size(1):
            # e.g. mask="no_linear": same shape error class as the reference's post Linear
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({n}x{K + F_in} and "
                               f"{first[0].weight.size(1)}x{first[0].weight.size(0)})")
        if T == 1:
            w = first[0].weight
            h = dense_linear(out[:, 0], w[:, F_in:], first[0].bias) + dense_linear(xt[:, 0], w[:, :F_in])
            hs = [h]
        else:
            # T per-tower Linears as ONE GEMM with a block-diagonal weight [T*F_out, T*K] (tcgen05 3xTF32): the
            # towers are tiny (ZINC: 450 -> 15), one launch beats T launches and the zero blocks cost nothing that
            # matters at this size; gradients of the zero blocks are simply never read back
            W = torch.stack([m.weight for m in first])                  # [T, F_out, F_in + K]
            b = torch.stack([m.bias for m in first])                    # [T, F_out]
            Fo = W.shape[1]
            h = dense_linear(out.reshape(n, T * K), torch.block_diag(*W[:, :, F_in:].unbind(0)), b.reshape(-1))
            if self.divide_input:
                h = h + dense_linear(xt.reshape(n, T * F_in), torch.block_diag(*W[:, :, :F_in].unbind(0)))
            else:
                h = h + dense_linear(xt[:, 0], W[:, :, :F_in].reshape(T * Fo, F_in))
            h = h.view(n, T, Fo)
            hs = [h[:, t] for t in range(T)]
        outs = []
        for t, seq in enumerate(self.post_nns):
            v = hs[t]
            for m in list(seq)[1:]:
                v = m(v)
            outs.append(v)
        out = outs[0] if T == 1 else torch.cat(outs, dim=1)
        if epilogue is None:
            return self.lin(out)
        scale, shift = epilogue
        b = self.lin.bias if self.lin.bias is not None else torch.zeros_like(shift)
        return dense_linear(out, self.lin.weight * scale.unsqueeze(1), b * scale + shift, relu=True)

    def _mask_projections(self, xt: Tensor, edge_attr: Optional[Tensor]):
        """P = X W_i^T + b, Q = X W_j^T (node-level) and R = enc(e) W_e^T (edge-level): the mask
        linear over cat([x_i, x_j, e]) of mma_conv.py:146-152 / mask_aggr.py:68, split by block."""
        T, F_in = self.towers, self.F_in
        n = xt.size(0)
        live = [seq[0].live() for seq in self.pre_nns[self.aggregators[-1]]]        # Q2
        W = torch.stack([m.weight for m in live])                                   # [T, F_in, (2|3)F_in]
        b = torch.stack([m.bias for m in live]).reshape(1, T * F_in)
        if xt.size(1) == 1:                     # towers share x: one GEMM gives P and Q of all towers
            Wpq = torch.cat([W[:, :, :F_in].reshape(T * F_in, F_in), W[:, :, F_in:2 * F_in].reshape(T * F_in, F_in)])
            PQ = dense_linear(xt[:, 0], Wpq)                                        # [N, 2*T*F_in]
            P = PQ[:, :T * F_in] + b
            Q = PQ[:, T * F_in:]
        else:                                   # divide_input: tower t sees its own slice -> block-diagonal weights
            xf = xt.reshape(n, T * F_in)
            P = dense_linear(xf, torch.block_diag(*W[:, :, :F_in].unbind(0))) + b
            Q = dense_linear(xf, torch.block_diag(*W[:, :, F_in:2 * F_in].unbind(0)))
        R = None
        if edge_attr is not None:
            e = self.edge_encoder(edge_attr)                                        # [E, F_in], :143
            R = dense_linear(e, W[:, :, 2 * F_in:].reshape(T * F_in, F_in))         # [E, T*F_in]
        return P, Q, R

    def _forward_fused(self, x: Tensor, graph: Graph, edge_attr: Optional[Tensor]) -> Tensor:
        """towers == 1 on a graph with degree-sorted CSR rows: the whole layer is one autograd node --
        two tcgen05 3xTF32 GEMMs around K1 in the forward, four in the backward (fused_layer.py)."""
        F_in = self.F_in
        self._check_names()
        for s_ in self.scalers:
            if s_ not in ("identity", "amplification", "attenuation", "linear", "inverse_linear"):
                raise ValueError(f'Unknown scaler "{s_}".')
        live = self.pre_nns[self.aggregators[-1]][0][0].live()                      # Q2
        R = None
        if edge_attr is not None:
            e = self.edge_encoder(edge_attr)                                        # [E, F_in], :143
            R = dense_linear(e, live.weight[:, 2 * F_in:])
        keep = self._inject_keep
        if keep is not None:
            keep = keep.reshape(graph.E, F_in)
        first = self.post_nns[0][0]
        seed_dev = self._next_seed_dev(x.device) if (self.device_seed and keep is None) else None
        return fused_layer.fused_mmaconv(
            x.contiguous(), graph, W_mask=live.weight, b_mask=live.bias, W_post=first.weight, b_post=first.bias,
            W_lin=self.lin.weight, b_lin=self.lin.bias, R=R, keep=keep, aggregators=self.aggregators,
            scalers=self.scalers, avg_deg=self.avg_deg, p_drop=self.dropout,
            seed=0 if seed_dev is not None else self._next_seed(), min_rows=self.fold_min_rows, seed_dev=seed_dev,
            n_windows=self.comm_slices, owner=self._uid)

    def _forward_folded(self, x: Tensor, graph: Graph, edge_attr: Optional[Tensor]) -> Tensor:
        """towers == 1 fast path: K1 emits the RAW aggregates Z [N, A*F_in] in degree-sorted row
        order; the S cumulative scalers are folded into the post weight per degree range
        (functional.scaled_post) -- the [N, S*A*F_in] tensor of mma_conv.py:196 and the cat with x
        (:132) are never materialised."""
        F_in, n = self.F_in, x.size(0)
        self._check_names()
        for s_ in self.scalers:
            if s_ not in ("identity", "amplification", "attenuation", "linear", "inverse_linear"):
                raise ValueError(f'Unknown scaler "{s_}".')
        P, Q, R = self._mask_projections(x.view(n, 1, F_in), edge_attr)
        if isinstance(graph, ShardedGraph):     # destination-range shard: x holds this rank's rows only
            if self._inject_keep is not None:
                raise RuntimeError("explicit keep masks are not supported on the sharded path")
            Z = sharded_mmconv_aggregate(P, Q, graph, F_in=F_in, aggregators=self.aggregators,
                                         scalers=["identity"], p_drop=self.dropout, seed=self._next_seed(),
                                         n_slices=self.comm_slices, sorted_rows=True)
            graph = graph.local
        else:
            keep = self._inject_keep
            if keep is not None:
                keep = keep.reshape(graph.E, F_in)
            Z = MF.mmconv_aggregate(P, Q, R, graph, towers=1, F_in=F_in, aggregators=self.aggregators,
                                    scalers=["identity"], keep=keep, p_drop=self.dropout,
                                    seed=self._next_seed())
        first = self.post_nns[0][0]
        W = first.weight                                                             # [F_out, (S*A+1)*F_in]
        Hs = MF.scaled_post(Z.view(n, -1), W[:, F_in:], graph, self.scalers, self.avg_deg,
                            min_rows=self.fold_min_rows or 512)
        h = Hs.index_select(0, graph.row_rank) + dense_linear(x, W[:, :F_in], first.bias)
        for m in list(self.post_nns[0])[1:]:
            h = m(h)
        return self.lin(h)

    def propagate(self, edge_index, size=None, **kwargs):
        """PyG's propagate for this layer (x_j = x[edge_index[0]], x_i = x[edge_index[1]],
        segments = edge_index[1]) -- fused when the mask linear is separable."""
        xt: Tensor = kwargs["x"]
        edge_attr: Optional[Tensor] = kwargs.get("edge_attr")
        n = xt.size(0)
        graph = self._graph(edge_index, n)
        T, F_in = self.towers, self.F_in
        self._check_names()
        if self.pre_layers != 1 or self.mask == "no_linear":
            return self._propagate_materialised(graph, edge_index, xt, edge_attr)

        P, Q, R = self._mask_projections(xt, edge_attr)
        if isinstance(graph, ShardedGraph):
            # destination-range shard of one large graph: x holds this rank's rows only
            if T != 1 or R is not None or self._inject_keep is not None:
                raise RuntimeError("the sharded path supports towers=1 without edge features")
            return sharded_mmconv_aggregate(P, Q, graph, F_in=F_in, aggregators=self.aggregators,
                                            scalers=self.scalers, avg_deg=self.avg_deg, p_drop=self.dropout,
                                            seed=self._next_seed(), n_slices=self.comm_slices,
                                            max_deg=self.global_max_deg)
        keep = self._inject_keep
        if keep is not None:
            keep = keep.reshape(graph.E, T * F_in)
        return MF.mmconv_aggregate(P, Q, R, graph, towers=T, F_in=F_in, aggregators=self.aggregators,
                                   scalers=self.scalers, avg_deg=self.avg_deg, keep=keep,
                                   p_drop=self.dropout, seed=self._next_seed())

    def _propagate_materialised(self, graph: Graph, edge_index, xt, edge_attr):
        """General path (pre_layers > 1 or mask == 'no_linear'): the mask MLP is not separable,
        so messages are built per edge with torch ops; reductions still run in K1."""
        T = self.towers
        if isinstance(edge_index, Graph):
            raise RuntimeError("the materialised-message path needs the raw edge_index tensor")
        if xt.size(1) == 1 and T > 1:
            xt = xt.repeat(1, T, 1)
        x_j = xt.index_select(0, edge_index[0])
        x_i = xt.index_select(0, edge_index[1])
        h = self._message_pre_dropout(x_i, x_j, edge_attr)                          # [E,T,F']
        keep = self._inject_keep
        Fm = h.size(-1)
        out = MF.mmconv_aggregate(None, None, h.reshape(h.size(0), T * Fm), graph, towers=T, F_in=Fm,
                                  aggregators=self.aggregators, scalers=self.scalers, avg_deg=self.avg_deg,
                                  keep=None if keep is None else keep.reshape(h.size(0), T * Fm),
                                  p_drop=self.dropout, seed=self._next_seed())
        return out

    def _message_pre_dropout(self, x_i: Tensor, x_j: Tensor, edge_attr: Optional[Tensor]) -> Tensor:
        if edge_attr is not None:
            edge_attr = self.edge_encoder(edge_attr)
            edge_attr = edge_attr.view(-1, 1, self.F_in).repeat(1, self.towers, 1)
            h = torch.cat([x_i, x_j, edge_attr], dim=-1)
        else:
            h = torch.cat([x_i, x_j], dim=-1)
        self._check_names()
        hs = [nn(h[:, i]) for i, nn in enumerate(self.pre_nns[self.aggregators[-1]])]
        return torch.stack(hs, dim=1)

    def message(self, x_i: Tensor, x_j: Tensor, edge_attr: Optional[Tensor]) -> Tensor:
        """mma_conv.py:138-157, callable on its own (dense torch ops; forward() does not use it)."""
        hs = self._message_pre_dropout(x_i, x_j, edge_attr)
        if self._inject_keep is not None:
            return hs * self._inject_keep
        return F.dropout(hs, self.dropout)

    def aggregate(self, inputs: Tensor, index: Tensor, dim_size: Optional[int] = None) -> Tensor:
        """mma_conv.py:159-196, callable on its own: inputs [E,T,F], index [E] -> [N,T,S*A*F]."""
        for a in self.aggregators:              # :164-174 / :182-194: exact names, ValueError whatever the device
            if a not in _lib.AGGR_KINDS:
                raise ValueError(f'Unknown aggregator "{a}".')
        for s_ in self.scalers:
            if s_ not in _lib.SCALER_KINDS:
                raise ValueError(f'Unknown scaler "{s_}".')
        if not inputs.is_cuda:
            raise RuntimeError("mma_b200.MMAConv.aggregate needs CUDA tensors (no CPU fallback)")
        n = int(index.max()) + 1 if dim_size is None else int(dim_size)
        E, T, Fm = inputs.shape
        graph = Graph.from_index(index, n)
        return MF.mmconv_aggregate(None, None, inputs.reshape(E, T * Fm), graph, towers=T, F_in=Fm,
                                   aggregators=self.aggregators, scalers=self.scalers, avg_deg=self.avg_deg)

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, "
                f"towers={self.towers}, edge_dim={self.edge_dim})")
