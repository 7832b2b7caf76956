"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (mma_b200/).

Loads the reference's hot-path files *verbatim* from /root/reference so that
golden vectors can be generated in the build container (oracle/make_golden.py)
and the restatement in oracle/restate.py can be validated against the real
code.  /root/reference does not exist on the GPU box: nothing here is used at
test/bench run time there; the committed fixtures under tests/golden/ are.

Why shims are needed:
  * graph_regression/mma_conv.py:2-12 and mask_aggr.py:4 import torch_geometric
    and torch_scatter, which are neither vendored in /root/reference nor
    installable here.  We register minimal stand-ins for exactly the seven
    third-party symbols those files import.  The only one with arithmetic on
    the path is torch_scatter.scatter (mma_conv.py:166,168,169), which is
    restated in oracle/restate.py::scatter (torch-scatter 2.0.9-era CPU
    semantics; see oracle/scatter_seq.c for the sequential ground truth).
  * node_classification/layers.py:10 imports scalers.py, whose line 2
    (`from utils import *`) fails under scipy>=1.8 and whose device is
    hard-coded to 'cuda:2' (scalers.py:12-61).  We exec the text of scalers.py
    minus line 2 with 'cuda:2' replaced by the CPU device, unchanged otherwise.

No reference source is copied into this repository: the files are read from
/root/reference at call time.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types
from typing import Optional

import torch
import torch.nn.functional as _F
from torch import Tensor

from . import restate

REF_ROOT = os.environ.get("MMA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "graph_regression", "mma_conv.py"))


# --------------------------------------------------------------------------
# torch_geometric / torch_scatter stand-ins
# --------------------------------------------------------------------------
class _PygLinear(torch.nn.Module):
    """torch_geometric.nn.dense.linear.Linear (PyG 2.0.x): weight [out,in],
    kaiming_uniform(fan=in, a=sqrt(5)) when weight_initializer is None, bias
    U(+-1/sqrt(in)) when bias_initializer is None; forward = F.linear."""

    def __init__(self, in_channels, out_channels, bias=True,
                 weight_initializer=None, bias_initializer=None):
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
        # inits.kaiming_uniform(weight, fan=in, a=sqrt(5)): bound = sqrt(6/((1+a^2) fan))
        bound = math.sqrt(6.0 / ((1.0 + 5.0) * self.in_channels))
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                b = 1.0 / math.sqrt(self.in_channels)
                self.bias.uniform_(-b, b)

    def forward(self, x):
        return _F.linear(x, self.weight, self.bias)


class _MessagePassing(torch.nn.Module):
    """torch_geometric.nn.conv.MessagePassing restricted to what mma_conv.py
    uses: ctor (aggr=None, node_dim=0), flow source_to_target, propagate ->
    message(x_i, x_j, edge_attr) -> aggregate(inputs, index, dim_size) ->
    identity update (mma_conv.py:53-54,130)."""

    def __init__(self, aggr=None, node_dim=0, **kwargs):
        super().__init__()
        self.aggr = aggr
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs["x"]
        x_j = x.index_select(0, edge_index[0])
        x_i = x.index_select(0, edge_index[1])
        msg = self.message(x_i=x_i, x_j=x_j, edge_attr=kwargs.get("edge_attr"))
        return self.aggregate(msg, edge_index[1], dim_size=x.size(0))


def _degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros((n,), dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.scatter_add_(0, index, torch.ones_like(index, dtype=out.dtype))


def _reset(value):
    if hasattr(value, "reset_parameters"):
        value.reset_parameters()
    else:
        for child in value.children() if hasattr(value, "children") else []:
            _reset(child)


def install_pyg_shims() -> None:
    if "torch_geometric" in sys.modules and not getattr(
            sys.modules["torch_geometric"], "_mma_oracle_shim", False):
        return  # a real PyG is importable: use it

    def mod(name):
        m = types.ModuleType(name)
        m._mma_oracle_shim = True
        sys.modules[name] = m
        return m

    tg = mod("torch_geometric")
    typing_m = mod("torch_geometric.typing")
    typing_m.Adj = Tensor
    typing_m.OptTensor = Optional[Tensor]
    nn_m = mod("torch_geometric.nn")
    conv_m = mod("torch_geometric.nn.conv")
    conv_m.MessagePassing = _MessagePassing
    dense_m = mod("torch_geometric.nn.dense")
    lin_m = mod("torch_geometric.nn.dense.linear")
    lin_m.Linear = _PygLinear
    utils_m = mod("torch_geometric.utils")
    utils_m.degree = _degree
    inits_m = mod("torch_geometric.nn.inits")
    inits_m.reset = _reset
    tg.typing, tg.nn, tg.utils = typing_m, nn_m, utils_m
    nn_m.conv, nn_m.dense, nn_m.inits = conv_m, dense_m, inits_m
    dense_m.linear = lin_m
    ts = mod("torch_scatter")
    ts.scatter = restate.scatter


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


class _FProxy:
    """Stands in for `torch.nn.functional` inside a loaded reference module so
    that F.dropout can be replaced by a deterministic keep-mask provider
    without touching torch globally (mma_conv.py:157, layers.py:219 ...)."""

    def __init__(self, dropout_fn):
        self._dropout_fn = dropout_fn

    def __getattr__(self, k):
        return getattr(_F, k)

    def dropout(self, x, p=0.5, training=True, inplace=False):
        return self._dropout_fn(x, p)


def load_graph_regression():
    """Returns (mma_conv module, mask_aggr module) executed verbatim."""
    install_pyg_shims()
    gr = os.path.join(REF_ROOT, "graph_regression")
    mask_aggr = _load("mask_aggr", os.path.join(gr, "mask_aggr.py"))
    mma_conv = _load("mma_conv", os.path.join(gr, "mma_conv.py"))
    return mma_conv, mask_aggr


def load_node_classification(device: str = "cpu"):
    """Returns (layers module, scalers module) executed verbatim (scalers.py
    minus its line 2, 'cuda:2' -> device)."""
    nc = os.path.join(REF_ROOT, "node_classification")
    with open(os.path.join(nc, "scalers.py")) as fh:
        lines = fh.read().split("\n")
    assert lines[1].strip() == "from utils import *", lines[1]
    lines[1] = ""
    text = "\n".join(lines).replace("'cuda:2'", repr(device))
    scalers = types.ModuleType("scalers")
    exec(compile(text, os.path.join(nc, "scalers.py"), "exec"), scalers.__dict__)
    sys.modules["scalers"] = scalers
    layers = _load("layers", os.path.join(nc, "layers.py"))
    return layers, scalers


def set_dropout(module, dropout_fn) -> None:
    """dropout_fn(x, p) -> tensor.  Pass None to restore torch's own."""
    module.F = _F if dropout_fn is None else _FProxy(dropout_fn)


class KeepMaskFeeder:
    """Feeds slices of a pre-drawn keep-scale tensor to successive F.dropout
    calls.  `chunks` is a list of tensors returned in call order (multiplied
    into x).  Used to inject identical dropout on the reference and on the
    restatement / CUDA side."""

    def __init__(self, chunks):
        self.chunks = list(chunks)
        self.pos = 0

    def __call__(self, x, p):
        k = self.chunks[self.pos]
        self.pos += 1
        assert k.shape == x.shape, (k.shape, x.shape)
        return x * k
