"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper over oracle/scatter_seq.c."""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libscatter_seq.so")
_lib = None


def build() -> str:
    src = os.path.join(_HERE, "scatter_seq.c")
    if not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def scatter_minmax(src: torch.Tensor, index: torch.Tensor, n: int, is_max: bool):
    src = src.contiguous().float(); index = index.contiguous().long()
    E, F = src.shape
    out = torch.empty((n, F), dtype=torch.float32); arg = torch.empty((n, F), dtype=torch.int64)
    fn = lib().seq_scatter_max if is_max else lib().seq_scatter_min
    fn(_p(src), _p(index), ctypes.c_int64(E), ctypes.c_int64(F), ctypes.c_int64(n), _p(out), _p(arg))
    return out, arg


def scatter_sum(src, index, n, mean=False):
    src = src.contiguous().float(); index = index.contiguous().long()
    E, F = src.shape
    out = torch.empty((n, F), dtype=torch.float32)
    fn = lib().seq_scatter_mean if mean else lib().seq_scatter_sum
    fn(_p(src), _p(index), ctypes.c_int64(E), ctypes.c_int64(F), ctypes.c_int64(n), _p(out))
    return out


def mmconv_aggregate(P, Q, R, keep, src_idx, dst_idx, n, F):
    """Returns dict of raw aggregates [n,F]: sum, mean, min, max, std, var, arg_min, arg_max."""
    c = lambda t: None if t is None else t.contiguous().float()
    P, Q, R, keep = c(P), c(Q), c(R), c(keep)
    src_idx = src_idx.contiguous().long(); dst_idx = dst_idx.contiguous().long()
    E = dst_idx.numel()
    o = {k: torch.empty((n, F), dtype=torch.float32) for k in ("sum", "mean", "min", "max", "std", "var")}
    o["arg_min"] = torch.empty((n, F), dtype=torch.int64); o["arg_max"] = torch.empty((n, F), dtype=torch.int64)
    lib().seq_mmconv_aggregate(_p(P), _p(Q), _p(R), _p(keep), _p(src_idx), _p(dst_idx),
                               ctypes.c_int64(E), ctypes.c_int64(F), ctypes.c_int64(n),
                               _p(o["sum"]), _p(o["mean"]), _p(o["min"]), _p(o["max"]),
                               _p(o["std"]), _p(o["var"]), _p(o["arg_min"]), _p(o["arg_max"]))
    return o
