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
        return dense_linear(x, self.weight, self.bias)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, bias={self.bias is not None})"


# ------------------------------------------------------------------------------------------
# the dense projections on the tensor cores
# ------------------------------------------------------------------------------------------
USE_TENSOR_CORES = True      # False: cuBLAS fp32 (A/B runs); CPU tensors always take F.linear (host-side tests only)


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _aligned_cols(t: Tensor) -> Tensor:
    """[M, K] fp32 -> a view / copy the TMA-fed GEMMs accept: unit column stride, 16-byte aligned rows (K padded with
    zero columns to a multiple of 4 when necessary; one pad kernel)."""
    M, K = t.shape
    if (t.stride(1) == 1 and K % 4 == 0 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0) or M == 0:
        return t
    return F.pad(t, (0, _pad4(K) - K)) if K % 4 else t.contiguous()


def _pad_rows4(t: Tensor) -> Tensor:
    """[N, ...] -> [N4, ...] with zero rows appended (the GEMM entry points take output widths in multiples of 4)."""
    N = t.shape[0]
    if N % 4 == 0:
        return t
    return F.pad(t, (0, 0) * (t.dim() - 1) + (0, _pad4(N) - N))


class _TcLinear(torch.autograd.Function):
    """y = x W^T (+ b) through the tcgen05 3xTF32 GEMMs of the C ABI (mma_linear_tf32x3 / mma_wgrad_tf32x3): the
    reference's fp32 Linears (graph_regression/mask_aggr.py:68, mma_conv.py:133-136,143; node_classification/
    layers.py:40,860) at fp32 accuracy (~1e-6), forward, dgrad and wgrad.  Widths that are not multiples of 4 floats
    (ZINC's 75, 50, 15; Pubmed's 3 classes) are zero-padded to the next multiple: TMA wants 16-byte aligned rows."""

    @staticmethod
    def forward(ctx, x: Tensor, W: Tensor, b: Optional[Tensor], relu: bool = False):
        from . import tc_gemm as tg
        M, K = x.shape
        N = W.shape[0]
        N4 = _pad4(N)
        xp = _aligned_cols(x)
        Wp = _pad_rows4(_aligned_cols(W.detach())).contiguous()    # [N4, K4]
        out = torch.empty((M, N4), dtype=torch.float32, device=x.device)
        if M > 0:
            hi, lo = tg.split_weight(Wp)
            bp = None if b is None else _pad_rows4(b.detach().contiguous())
            tg.linear(xp, hi, lo, N4, out=out, bias=bp, relu=relu, name="gemm_linear")
        ctx.save_for_backward(xp, Wp, out if relu else None)
        ctx.dims = (M, K, N, b is not None)
        return out[:, :N]

    @staticmethod
    def backward(ctx, dy: Tensor):
        from . import tc_gemm as tg
        xp, Wp, y = ctx.saved_tensors
        M, K, N, has_b = ctx.dims
        dx = dW = db = None
        if M == 0:
            return (xp.new_zeros((0, K)), Wp.new_zeros((N, K)), Wp.new_zeros(N) if has_b else None, None)
        if y is not None:                                           # ReLU in the epilogue: gradient only where y > 0
            dy = dy * (y[:, :N] > 0)
        dyp = _aligned_cols(dy)                                     # [M, N4]
        K4 = Wp.shape[1]
        if ctx.needs_input_grad[0]:
            Wt = Wp.t().contiguous()                                # [K4, N4]
            hi, lo = tg.split_weight(Wt)
            buf = torch.empty((M, K4), dtype=torch.float32, device=dy.device)
            tg.linear(dyp, hi, lo, K4, out=buf, name="gemm_linear_dgrad")
            dx = buf[:, :K]
        if ctx.needs_input_grad[1]:
            dW = tg.wgrad(dyp, xp, name="gemm_linear_wgrad")[:N, :K]
        if has_b and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dW, db, None


def dense_linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, relu: bool = False) -> Tensor:
    """F.linear(x, weight, bias) (then ReLU, in the GEMM epilogue) for a [..., K] input: CUDA fp32 tensors go through
    the tensor-core GEMMs."""
    if not (USE_TENSOR_CORES and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32
            and x.shape[-1] == weight.shape[1]):
        y = F.linear(x, weight, bias)
        return F.relu(y) if relu else y
    lead = x.shape[:-1]
    y = _TcLinear.apply(x.reshape(-1, x.shape[-1]), weight, bias, relu)
    return y.reshape(*lead, weight.shape[0])


def dense_mm(a: Tensor, b: Tensor) -> Tensor:
    """torch.mm(a, b) with b [K, N] (node_classification/layers.py:40,219,860) on the tensor cores."""
    if not (USE_TENSOR_CORES and a.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32):
        return torch.mm(a, b)
    return _TcLinear.apply(a, b.t(), None, False)


def reset(value) -> None:
    """torch_geometric.nn.inits.reset (graph_regression/mma_conv.py:12)."""
    if hasattr(value, "reset_parameters"):
        value.reset_parameters()
    else:
        for child in value.children() if hasattr(value, "children") else []:
            reset(child)
