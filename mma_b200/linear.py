"""Dense linear layer with PyG's `torch_geometric.nn.dense.linear.Linear` semantics
(the class the reference imports at graph_regression/mma_conv.py:9 and subclasses at
graph_regression/mask_aggr.py:7): weight [out,in], kaiming_uniform(fan=in, a=sqrt(5)) /
bias U(+-1/sqrt(in)) when no initializer is named, forward = x @ W^T + b."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor


class Linear(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None, bias_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight_initializer = weight_initializer
        self.bias_initializer = bias_initializer
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            if self.weight_initializer == "glorot":
                a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
                self.weight.uniform_(-a, a)
            elif self.weight_initializer in (None, "kaiming_uniform"):
                bound = math.sqrt(6.0 / ((1.0 + 5.0) * self.in_channels))
                self.weight.uniform_(-bound, bound)
            else:
                raise RuntimeError(f"Linear layer weight initializer '{self.weight_initializer}' is not supported")
            if self.bias is not None:
                if self.bias_initializer == "zeros":
                    self.bias.zero_()
                elif self.bias_initializer is None:
                    b = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
                    self.bias.uniform_(-b, b)
                else:
                    raise RuntimeError(f"Linear layer bias initializer '{self.bias_initializer}' is not supported")

    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, self.weight, self.bias)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, bias={self.bias is not None})"


def reset(value) -> None:
    """torch_geometric.nn.inits.reset (graph_regression/mma_conv.py:12)."""
    if hasattr(value, "reset_parameters"):
        value.reset_parameters()
    else:
        for child in value.children() if hasattr(value, "children") else []:
            reset(child)
