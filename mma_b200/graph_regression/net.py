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
        self.mlp = Sequential(Linear(75, 50), ReLU(), Linear(50, 25), ReLU(), Linear(25, 1))

    def forward(self, x, edge_index, edge_attr, batch):
        x = self.node_emb(x.squeeze())
        edge_attr = self.edge_emb(edge_attr)
        for conv, batch_norm in zip(self.convs, self.batch_norms):
            x = F.relu(batch_norm(conv(x, edge_index, edge_attr)))
        x = global_add_pool(x, batch)
        return self.mlp(x)
