"""Drop-in for /root/reference/graph_regression/mask_aggr.py (`MaskAggregateLinear`).

Same constructor, attributes and quirks:
  * one extra Linear(in,out) PER aggregator name, held in a plain dict
    `aggregation_layers` (mask_aggr.py:44-51) -> those weights are not registered
    parameters (absent from parameters()/state_dict(), Q1) but receive .grad;
  * they are created on 'cuda' when available (mask_aggr.py:39,50);
  * forward applies the one for `self.aggregation` (mask_aggr.py:68);
    mask == "no_linear" returns the input (mask_aggr.py:65-66).
Inside MMAConv the mask linear is not run per edge: it is split into node-level
GEMMs (P, Q) and an edge term R that the fused CUDA kernel adds per edge.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ..linear import Linear


class MaskAggregateLinear(Linear):
    def __init__(self, in_channels: int, out_channels: int, aggregation_list: List[str], aggregation: str,
                 mask=True, bias: bool = True, weight_initializer: Optional[str] = None,
                 bias_initializer: Optional[str] = None):
        super().__init__(in_channels, out_channels, bias, weight_initializer, bias_initializer)
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.mask = mask
        self.aggregation = aggregation
        self.aggregation_layers = {}
        for aggr in aggregation_list:
            name = "{}".format(aggr)
            if self.mask == "no_linear":
                self.aggregation_layers[name] = None
            else:
                self.aggregation_layers[name] = Linear(in_channels, out_channels, bias, weight_initializer,
                                                       bias_initializer).to(self.device)

    def live(self) -> Optional[Linear]:
        """The one Linear that forward applies."""
        return self.aggregation_layers[self.aggregation]

    def forward(self, input):
        if self.aggregation not in self.aggregation_layers:
            raise ValueError("Invalid aggregation type: {}".format(self.aggregation))
        if self.mask == "no_linear":
            return input
        return self.aggregation_layers[self.aggregation](input).to(self.device)
