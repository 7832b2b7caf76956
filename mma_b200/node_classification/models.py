"""Caller of the hot path for BASELINE config 1/3: /root/reference/node_classification/models.py
restated (2-layer net: GraphConvolution -> ReLU -> dropout -> MMA -> log_softmax, models.py:64-68).
Same constructor signature; parameters are allocated with torch.empty on `device` instead of the
removed `torch.cuda.FloatTensor` (models.py:17-43)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .layers import GraphConvolution, MMA, _ALL


class MMAConv(nn.Module):
    def __init__(self, add_all, activation, k, nfeat, nhid, nclass, dropout, aggregator_list, device):
        super(MMAConv, self).__init__()
        self.device = device
        new = lambda *shape: nn.Parameter(torch.empty(*shape, device=device))
        self.weight0, self.bias0 = new(nfeat, nhid), new(nhid)
        self.weight1, self.bias1 = new(nhid, nclass), new(nclass)
        for name in _ALL:
            setattr(self, "weight_" + name, new(2 * nhid, nhid))
        self.add_all = add_all
        self.gc1 = GraphConvolution(nfeat, nhid, self.weight0, self.bias0, device)
        self.gc2 = MMA(self.add_all, activation, k, nhid, nclass, self.weight1, self.bias1,
                       *[getattr(self, "weight_" + name) for name in _ALL], dropout, aggregator_list, device)
        self.dropout = dropout

    def forward(self, x, adj):
        x = F.relu(self.gc1(x, adj))
        x = F.dropout(x, self.dropout, training=self.training)
        x = self.gc2(x, adj)
        return F.log_softmax(x, dim=1)
