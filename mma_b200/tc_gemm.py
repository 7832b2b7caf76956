"""Thin wrappers over the tcgen05 3xTF32 GEMM entry points of the C ABI (include/mma_b200.h:
mma_tf32_split, mma_linear_tf32x3, mma_wgrad_tf32x3, mma_reduce_slabs).

They stand in for the reference's fp32 Linears (mask_aggr.py:68, mma_conv.py:132-136) and their
autograd backward.  No CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

BM = BN = 128
BK = 32
ADD_BY_INPUT_ROW = 4    # MMA_GEMM_ADD_BY_INPUT_ROW
RELU = 8                # MMA_GEMM_RELU
MODE = 1            # 0: 3xTF32, hi rounded to nearest and written back; 1: raw operand as (truncated) hi -- same
                    # measured accuracy, a 5x cheaper split; 2: plain TF32 (not fp32-accurate)


def usable(*mats: Optional[Tensor]) -> bool:
    """TMA needs 16-byte aligned bases and row strides; everything else falls back to torch.mm."""
    for t in mats:
        if t is None:
            continue
        if (not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % 4 != 0
                or t.data_ptr() % 16 != 0 or t.shape[0] < 1 or t.shape[1] < 1):
            return False
    return True


def split_weight(W: Tensor) -> Tuple[Tensor, Tensor]:
    """(hi, lo) of a weight matrix, hi exactly representable in TF32."""
    dev = _lib.require_cuda(W)
    W = W.detach().contiguous()
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    with _lib.kernel_scope("mma_tf32_split", dev):
        _lib.check(_lib.lib().mma_tf32_split(_lib.ptr(W), _lib.ptr(hi), _lib.ptr(lo), W.numel(),
                                             _lib.stream_ptr(dev)), "mma_tf32_split")
    return hi, lo


def linear(A0: Tensor, Whi: Tensor, Wlo: Tensor, n_out: int, *, A1: Optional[Tensor] = None,
           tile_tab: Optional[Tensor] = None, out: Optional[Tensor] = None, out_map: Optional[Tensor] = None,
           bias: Optional[Tensor] = None, add: Optional[Tensor] = None, add_by_input_row: bool = False,
           mode: Optional[int] = None, max_ctas: int = 0, relu: bool = False,
           name: str = "mma_linear_tf32x3") -> Tensor:
    """out[out_map[r]] = [A0|A1][r] @ W[b_off : b_off + n_out].T (+ bias) (+ add[out_map[r]], or add[r]
    with add_by_input_row).

    Whi/Wlo: [b_rows, K0+K1] (stacked weights for a grouped GEMM; `tile_tab` int32 [tiles, 4] =
    (row0, row_end, b_off, 0) selects the weight per 128-row tile)."""
    dev = _lib.require_cuda(A0, Whi, Wlo)
    M, K0 = A0.shape
    K1 = 0 if A1 is None else A1.shape[1]
    if not usable(A0, A1, Whi, Wlo, out, add):
        raise RuntimeError("mma_linear_tf32x3: operands must be fp32, row-major, 16-byte aligned with ld % 4 == 0")
    if Whi.shape != Wlo.shape or Whi.shape[1] != K0 + K1 or Whi.stride(0) != Wlo.stride(0):
        raise RuntimeError(f"weight halves {tuple(Whi.shape)}/{tuple(Wlo.shape)} do not match K = {K0 + K1}")
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float32, device=dev)
    if tile_tab is not None and (tile_tab.dtype != torch.int32 or tile_tab.dim() != 2 or tile_tab.shape[1] != 4
                                 or not tile_tab.is_contiguous()):
        raise RuntimeError("tile_tab must be a contiguous int32 [tiles, 4] tensor")
    with _lib.kernel_scope(name, dev):
        _lib.check(_lib.lib().mma_linear_tf32x3(
            _lib.ptr(A0), A0.stride(0), K0, _lib.ptr(A1), 0 if A1 is None else A1.stride(0), K1,
            _lib.ptr(Whi), _lib.ptr(Wlo), Whi.stride(0), Whi.shape[0], M, n_out,
            _lib.ptr(tile_tab), 0 if tile_tab is None else tile_tab.shape[0],
            _lib.ptr(out), out.stride(0), _lib.ptr(out_map), _lib.ptr(bias), _lib.ptr(add),
            0 if add is None else add.stride(0),
            (MODE if mode is None else mode) | (ADD_BY_INPUT_ROW if add_by_input_row else 0) | (RELU if relu else 0),
            max_ctas,
            _lib.stream_ptr(dev)), name)
    return out


MAX_SLAB_ROWS = 1024    # the tensor core truncates on accumulate: keep accumulation chains short (128 MMAs)


def make_slabs(M: int, tiles: int, device, sms: int = 148, waves: int = 2, max_rows: int = MAX_SLAB_ROWS) -> Tensor:
    """Cuts the reduction rows [0, M) into slabs (multiples of 32 rows, at most `max_rows`) so that
    slabs x tiles fills at least `waves` waves of persistent CTAs; int32 [n_slabs, 4] = (row0, row_end, slot, 0)."""
    n_slabs = max(1, min((sms * waves + tiles - 1) // tiles, (M + BK - 1) // BK))
    rows = min(((M + n_slabs - 1) // n_slabs + BK - 1) // BK * BK, max_rows)
    tab = []
    r = 0
    while r < M:
        tab.append((r, min(M, r + rows), len(tab), 0))
        r += rows
    return torch.tensor(tab, dtype=torch.int32, device=device)


def wgrad_partials(G0: Tensor, A: Tensor, slabs: Tensor, n_slots: int, *, G1: Optional[Tensor] = None,
                   mode: Optional[int] = None, max_ctas: int = 0, name: str = "mma_wgrad_tf32x3") -> Tensor:
    """part[slot] = [G0|G1][rows of the slab].T @ A[rows of the slab]  -> [n_slots, N0+N1, K]."""
    dev = _lib.require_cuda(G0, A, slabs)
    M, N0 = G0.shape
    N1 = 0 if G1 is None else G1.shape[1]
    K = A.shape[1]
    if not usable(G0, G1, A) or A.shape[0] != M or K % 4 != 0:
        raise RuntimeError("mma_wgrad_tf32x3: operands must be fp32, row-major, 16-byte aligned with ld % 4 == 0")
    part = torch.empty((n_slots, N0 + N1, K), dtype=torch.float32, device=dev)
    with _lib.kernel_scope(name, dev):
        _lib.check(_lib.lib().mma_wgrad_tf32x3(
            _lib.ptr(G0), G0.stride(0), N0, _lib.ptr(G1), 0 if G1 is None else G1.stride(0), N1,
            _lib.ptr(A), A.stride(0), K, M, _lib.ptr(slabs), slabs.shape[0], _lib.ptr(part),
            MODE if mode is None else mode, max_ctas, _lib.stream_ptr(dev)), name)
    return part


def reduce_slabs(part: Tensor, coef: Optional[Tensor] = None) -> Tensor:
    dev = _lib.require_cuda(part)
    n_slots = part.shape[0]
    out = torch.empty(part.shape[1:], dtype=torch.float32, device=dev)
    with _lib.kernel_scope("mma_reduce_slabs", dev):
        _lib.check(_lib.lib().mma_reduce_slabs(_lib.ptr(part), _lib.ptr(coef), n_slots, out.numel(), _lib.ptr(out),
                                               _lib.stream_ptr(dev)), "mma_reduce_slabs")
    return out


def reduce_slabs_segmented(part: Tensor, seg_ptr: Tensor) -> Tensor:
    """out[g] = sum of part[seg_ptr[g] : seg_ptr[g+1]] (fixed ascending order) -> [n_segs, *part.shape[1:]]."""
    dev = _lib.require_cuda(part, seg_ptr)
    n_segs = seg_ptr.numel() - 1
    out = torch.empty((n_segs,) + tuple(part.shape[1:]), dtype=torch.float32, device=dev)
    n = part[0].numel() if part.shape[0] else 0
    with _lib.kernel_scope("mma_reduce_slabs", dev):
        _lib.check(_lib.lib().mma_reduce_slabs_segmented(_lib.ptr(part), _lib.ptr(seg_ptr), n_segs, n, _lib.ptr(out),
                                                         _lib.stream_ptr(dev)), "mma_reduce_slabs_segmented")
    return out


_SLAB_CACHE = {}


def cached_slabs(M: int, tiles: int, device) -> Tensor:
    key = (int(M), int(tiles), str(device))
    t = _SLAB_CACHE.get(key)
    if t is None:
        if len(_SLAB_CACHE) > 64:
            _SLAB_CACHE.clear()
        t = _SLAB_CACHE[key] = make_slabs(M, tiles, device)
    return t


def wgrad(G0: Tensor, A: Tensor, *, G1: Optional[Tensor] = None, mode: Optional[int] = None,
          name: str = "mma_wgrad_tf32x3") -> Tensor:
    """dW [N0+N1, K] = [G0|G1].T @ A  (split-K over slabs + fixed-order reduction)."""
    N = G0.shape[1] + (0 if G1 is None else G1.shape[1])
    tiles = ((N + BM - 1) // BM) * ((A.shape[1] + BN - 1) // BN)
    slabs = cached_slabs(G0.shape[0], tiles, G0.device)
    part = wgrad_partials(G0, A, slabs, slabs.shape[0], G1=G1, mode=mode, name=name)
    return reduce_slabs(part)


GATHER_PARTS = 148 * 32     # warps of the gather kernel = partial column sums


def gather_rows_colsum(src: Tensor, idx32: Tensor) -> Tuple[Tensor, Tensor]:
    """(src[idx], src[idx].sum(0)) in one pass: the row permutation into degree-sorted order fused
    with the bias-gradient column sum (mma_gather_rows + fixed-order mma_reduce_slabs)."""
    dev = _lib.require_cuda(src, idx32)
    n, F = idx32.numel(), src.shape[1]
    if not usable(src) or F % 4 != 0 or idx32.dtype != torch.int32:
        raise RuntimeError("gather_rows_colsum: src must be fp32 row-major, 16-byte aligned, F % 4 == 0; idx int32")
    out = torch.empty((n, F), dtype=torch.float32, device=dev)
    parts = max(1, min(GATHER_PARTS, (n + 63) // 64))
    part = torch.empty((parts, F), dtype=torch.float32, device=dev)
    with _lib.kernel_scope("mma_gather_rows", dev):
        _lib.check(_lib.lib().mma_gather_rows(_lib.ptr(src), src.stride(0), _lib.ptr(idx32), n, F, _lib.ptr(out), F,
                                              _lib.ptr(part), parts, _lib.stream_ptr(dev)), "mma_gather_rows")
    return out, reduce_slabs(part)


def colsum(src: Tensor) -> Tensor:
    """src.sum(0) of a (possibly strided) fp32 [n, F] view in one streaming pass: per-part partial sums
    (mma_gather_rows with out = NULL) + fixed-order mma_reduce_slabs -- deterministic, no atomics."""
    dev = _lib.require_cuda(src)
    n, F = src.shape
    if src.dtype != torch.float32 or src.stride(1) != 1 or F % 4 != 0 or src.stride(0) % 4 != 0 or src.data_ptr() % 16:
        raise RuntimeError("colsum: src must be fp32 with unit column stride, 16-byte aligned rows, F % 4 == 0")
    parts = max(1, min(GATHER_PARTS, (n + 63) // 64))
    part = torch.empty((parts, F), dtype=torch.float32, device=dev)
    with _lib.kernel_scope("mma_gather_rows", dev):
        _lib.check(_lib.lib().mma_gather_rows(_lib.ptr(src), src.stride(0), None, n, F, None, 0,
                                              _lib.ptr(part), parts, _lib.stream_ptr(dev)), "mma_gather_rows")
    return reduce_slabs(part)


# ---------------------------------------------------------------------------------------------
# weight-space algebra of the fused layer (csrc/weight_prep.cu)
# ---------------------------------------------------------------------------------------------
def small_gemm(A: Tensor, B: Tensor, *, trans_a: bool = False, trans_b: bool = False, k_splits: int = 1,
               name: str = "mma_small_gemm") -> Tensor:
    """op(A) @ op(B) for the small weight-space products (plain fp32 FFMA, fixed summation order).  A / B: fp32 with
    unit column stride.  k_splits > 1: split-K into slabs added in order by mma_reduce_slabs (long-K, few outputs)."""
    dev = _lib.require_cuda(A, B)
    if A.dtype != torch.float32 or B.dtype != torch.float32 or A.dim() != 2 or B.dim() != 2 or A.stride(1) != 1 or B.stride(1) != 1:
        raise RuntimeError("small_gemm: fp32 matrices with unit column stride")
    M, K = (A.shape[1], A.shape[0]) if trans_a else A.shape
    Kb, N = (B.shape[1], B.shape[0]) if trans_b else B.shape
    if K != Kb:
        raise RuntimeError(f"small_gemm: inner dimensions {K} and {Kb} differ")
    splits = _lib.lib().mma_small_gemm_splits(K, k_splits)
    C = torch.empty((splits, M, N) if splits > 1 else (M, N), dtype=torch.float32, device=dev)
    with _lib.kernel_scope(name, dev):
        _lib.check(_lib.lib().mma_small_gemm(_lib.ptr(A), A.stride(0), int(trans_a), _lib.ptr(B), B.stride(0), int(trans_b),
                                             _lib.ptr(C), N, M, N, K, splits, _lib.stream_ptr(dev)), name)
    if splits == 1:
        return C
    return reduce_slabs(C) if (M * N) % 4 == 0 else C.sum(0)        # slabs added in ascending order either way


def compose_post_weight(coef: Tensor, WlW: Tensor, col0: int, F: int, transposed: bool):
    """coef [B, S, A, Am], WlW [Co, col0 + S*A*F] -> (hi, lo) [B, Co, Am*F] of the composed per-range weight, split for
    the 3xTF32 GEMMs, and (hiT, loT) [B, Am*F, Co] when `transposed` (else None)."""
    dev = _lib.require_cuda(coef, WlW)
    B, S, A, Am = coef.shape
    Co = WlW.shape[0]
    if not coef.is_contiguous() or WlW.stride(1) != 1 or WlW.shape[1] < col0 + S * A * F:
        raise RuntimeError("compose_post_weight: coef contiguous [B,S,A,Am]; WlW [Co, >= col0 + S*A*F]")
    hi = torch.empty((B, Co, Am * F), dtype=torch.float32, device=dev)
    lo = torch.empty_like(hi)
    hiT = torch.empty((B, Am * F, Co), dtype=torch.float32, device=dev) if transposed else None
    loT = torch.empty_like(hiT) if transposed else None
    with _lib.kernel_scope("mma_compose_post_weight", dev):
        _lib.check(_lib.lib().mma_compose_post_weight(_lib.ptr(coef), B, S, A, Am, _lib.ptr(WlW), WlW.stride(0), col0, Co, F,
                                                      _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(hiT), _lib.ptr(loT),
                                                      _lib.stream_ptr(dev)), "mma_compose_post_weight")
    return hi, lo, hiT, loT


def compose_post_wgrad(coef: Tensor, block_of: Tensor, dWc: Tensor, dX: Optional[Tensor], col0: int, F: int) -> Tensor:
    """coef [B, S, A, Am], dWc [B, Co, Am*F], dX [Co, col0] -> D [Co, col0 + S*A*F]: the gradient of W_lin W_post in
    W_post's column layout ([x-part | (s, a) blocks])."""
    dev = _lib.require_cuda(coef, dWc, block_of)
    B, S, A, Am = coef.shape
    Co = dWc.shape[1]
    if not coef.is_contiguous() or not dWc.is_contiguous() or block_of.dtype != torch.int32 or block_of.numel() != A:
        raise RuntimeError("compose_post_wgrad: coef / dWc contiguous, block_of int32 [A]")
    if dX is not None and (dX.shape != (Co, col0) or dX.stride(1) != 1):
        raise RuntimeError("compose_post_wgrad: dX must be [Co, col0] with unit column stride")
    D = torch.empty((Co, col0 + S * A * F), dtype=torch.float32, device=dev)
    with _lib.kernel_scope("mma_compose_post_wgrad", dev):
        _lib.check(_lib.lib().mma_compose_post_wgrad(_lib.ptr(coef), B, S, A, Am, _lib.ptr(block_of), _lib.ptr(dWc), Co, F,
                                                     _lib.ptr(dX), 0 if dX is None else dX.stride(0), col0, _lib.ptr(D),
                                                     D.stride(0), _lib.stream_ptr(dev)), "mma_compose_post_wgrad")
    return D
