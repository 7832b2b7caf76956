"""Caller of the MultiMaskConv layer in BASELINE config 2: the `Net` of
/root/reference/graph_regression/mma.py:62-127 restated around the drop-in layer --
Embedding(21, 75) / Embedding(4, 50) -> 4 x [MMAConv(75 -> 75, towers 5, edge_dim 50) -> BatchNorm -> ReLU]
-> global_add_pool -> MLP(75 -> 50 -> 25 -> 1).

Same constructor as the reference's class plus the degree histogram, which mma.py reads from a module
global (`deg`, mma.py:57-60); `args` only needs a `.mask` attribute (mma.py:94).  `global_add_pool`
(PyG, a scatter-add by `batch`) runs through the deterministic segmented row sum of the C-ABI library
(K3, `mma_segment_sum_rows`): PyG batches number their nodes graph by graph, so `batch` is sorted and the
segments are contiguous; any other order goes through a stable sort first.  No CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.nn import BatchNorm1d, Embedding, Linear, ModuleList, ReLU, Sequential

from .. import _lib
from .. import functional as MF
from ..graph import csr_build
from ..linear import dense_linear
from .mma_conv import MMAConv


def global_add_pool(x: Tensor, batch: Tensor, size: Optional[int] = None) -> Tensor:
    """out[g] = sum of the rows of x whose batch id is g (torch_geometric.nn.global_add_pool, mma.py:124)."""
    _lib.require_cuda(x, batch)
    n_graphs = int(size) if size is not None else (int(batch.max()) + 1 if batch.numel() else 0)
    if batch.numel() == 0 or bool((batch[1:] >= batch[:-1]).all()):
        counts = torch.bincount(batch, minlength=n_graphs)
        ptr = torch.zeros(n_graphs + 1, dtype=torch.int32, device=x.device)
        ptr[1:] = torch.cumsum(counts, 0)
        idx = None
    else:
        ptr, _, idx = csr_build(batch, None, n_graphs)          # stable: rows of a graph keep their order
    ptr_t = torch.arange(x.shape[0] + 1, dtype=torch.int32, device=x.device)          # transpose: one graph per node
    idx_t = batch.to(torch.int32).contiguous()
    return MF.segment_sum_rows(x.contiguous(), ptr, idx, None, n_graphs, ptr_t, idx_t, None)


class _BatchNormReLU(torch.autograd.Function):
    """relu(batch_norm(x)) over the rows of x [n, F] as one kernel per direction (mma_bn_relu_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, use_batch_stats: bool, momentum: float, eps: float):
        dev = _lib.require_cuda(x)
        x = x if (x.dim() == 2 and x.stride(1) == 1) else x.contiguous()
        n, Fd = x.shape
        y = torch.empty((n, Fd), dtype=torch.float32, device=dev)
        save_mean = torch.empty(Fd, dtype=torch.float32, device=dev)
        save_rstd = torch.empty(Fd, dtype=torch.float32, device=dev)
        upd = use_batch_stats and running_mean is not None
        with _lib.kernel_scope("mma_bn_relu_fwd", dev):
            _lib.check(_lib.lib().mma_bn_relu_fwd(
                _lib.ptr(x), x.stride(0), n, Fd, _lib.ptr(gamma), _lib.ptr(beta), float(eps),
                None if use_batch_stats else _lib.ptr(running_mean), None if use_batch_stats else _lib.ptr(running_var),
                _lib.ptr(y), Fd, _lib.ptr(save_mean), _lib.ptr(save_rstd), _lib.ptr(running_mean) if upd else None,
                _lib.ptr(running_var) if upd else None, float(momentum), _lib.stream_ptr(dev)), "mma_bn_relu_fwd")
        ctx.save_for_backward(x, y, gamma, save_mean, save_rstd)
        ctx.batch_stats = bool(use_batch_stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, save_mean, save_rstd = ctx.saved_tensors
        dev = dy.device
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        n, Fd = x.shape
        dx = torch.empty((n, Fd), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        dgamma = torch.zeros(Fd, dtype=torch.float32, device=dev)
        dbeta = torch.zeros(Fd, dtype=torch.float32, device=dev)
        with _lib.kernel_scope("mma_bn_relu_bwd", dev):
            _lib.check(_lib.lib().mma_bn_relu_bwd(
                _lib.ptr(x), x.stride(0), _lib.ptr(y), Fd, _lib.ptr(dy), dy.stride(0), n, Fd, _lib.ptr(gamma),
                _lib.ptr(save_mean), _lib.ptr(save_rstd), 1 if ctx.batch_stats else 0, _lib.ptr(dx), Fd,
                _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.stream_ptr(dev)), "mma_bn_relu_bwd")
        return (dx, dgamma if gamma is not None else None, dbeta if gamma is not None else None,
                None, None, None, None, None)


def batch_norm_relu(x: Tensor, bn: BatchNorm1d) -> Tensor:
    """F.relu(bn(x)) for a torch BatchNorm1d module `bn` (its parameters and running statistics are used and, in
    training mode, updated exactly as the module would), one fused kernel per direction."""
    if not x.is_cuda:
        raise RuntimeError("mma_b200 needs CUDA tensors (no CPU fallback)")
    batch_stats = bn.training or bn.running_mean is None
    momentum = 0.0
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        momentum = (1.0 / float(bn.num_batches_tracked)) if bn.momentum is None else bn.momentum
    return _BatchNormReLU.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, batch_stats, momentum, bn.eps)


class _Linear(Linear):
    """torch.nn.Linear (the class mma.py:97-98 builds its MLP from: same parameters, same initialisation, same
    state_dict) whose forward runs on the tensor cores like every other dense projection of the path."""

    def forward(self, input: Tensor) -> Tensor:
        return dense_linear(input, self.weight, self.bias)


class Net(torch.nn.Module):
    def __init__(self, args, aggregator_list, scaler_list, deg: Tensor):
        super(Net, self).__init__()
        self.node_emb = Embedding(21, 75)
        self.edge_emb = Embedding(4, 50)
        self.convs = ModuleList()
        self.batch_norms = ModuleList()
        for _ in range(4):
            conv = MMAConv(in_channels=75, out_channels=75, aggregators=aggregator_list, scalers=scaler_list, deg=deg,
                           edge_dim=50, towers=5, pre_layers=1, post_layers=1,
                           mask=getattr(args, "mask", True), divide_input=False)
            self.convs.append(conv)
            self.batch_norms.append(BatchNorm1d(75))            # torch_geometric.nn.BatchNorm wraps BatchNorm1d
        self.mlp = Sequential(_Linear(75, 50), ReLU(), _Linear(50, 25), ReLU(), _Linear(25, 1))

    def forward(self, x, edge_index, edge_attr, batch):
        x = self.node_emb(x.squeeze())
        edge_attr = self.edge_emb(edge_attr)
        for conv, batch_norm in zip(self.convs, self.batch_norms):
            # mma.py:120-121  x = F.relu(batch_norm(conv(x, edge_index, edge_attr)))
            if batch_norm.training or batch_norm.running_mean is None:
                x = batch_norm_relu(conv(x, edge_index, edge_attr), batch_norm)       # batch statistics: one fused kernel
            else:
                # eval: BatchNorm is a per-channel affine map -> folded into the layer's `lin`, ReLU in the GEMM epilogue
                scale = batch_norm.weight * torch.rsqrt(batch_norm.running_var + batch_norm.eps)
                shift = batch_norm.bias - batch_norm.running_mean * scale
                x = conv.forward_affine_relu(x, edge_index, edge_attr, scale, shift)
        x = global_add_pool(x, batch)
        return self.mlp(x)
