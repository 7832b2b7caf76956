"""Device-side graph structures for the aggregation kernels.

The reference hands an unsorted `edge_index` [2,E] to PyG's propagate
(graph_regression/mma_conv.py:130) and Python neighbour lists `add_all` to the
node-classification layer (node_classification/utils.py:98-100).  The kernels
work on a CSR by destination (stable sort => in-row order == original edge
order) and, for the backward, on its transpose.  Building them is a one-off per
graph; results are cached.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional

import torch
from torch import Tensor

from . import _lib


def csr_build(key: Tensor, other: Optional[Tensor], n_keys: int):
    """Stable sort of edges by `key` (int64 [E]) -> (rowptr [n_keys+1], col [E] = other[perm],
    perm [E]) as int32 CUDA tensors.  Calls mma_csr_build (CUB radix sort)."""
    dev = _lib.require_cuda(key, other)
    key = key.contiguous()
    if key.dtype != torch.int64:
        key = key.to(torch.int64)
    if other is not None:
        other = other.contiguous().to(torch.int64)
    E = key.numel()
    l = _lib.lib()
    rowptr = torch.empty(n_keys + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev) if other is not None else None
    perm = torch.empty(E, dtype=torch.int32, device=dev)
    nbytes = C.c_size_t(0)
    _lib.check(l.mma_csr_build_workspace_bytes(E, n_keys, C.byref(nbytes)), "mma_csr_build_workspace_bytes")
    ws = torch.empty(max(int(nbytes.value), 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(l.mma_csr_build(_lib.ptr(key), _lib.ptr(other), E, n_keys, _lib.ptr(rowptr),
                                   _lib.ptr(col), _lib.ptr(perm), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr(dev)), "mma_csr_build")
    return rowptr, col, perm


def invert_perm(perm: Tensor) -> Tensor:
    dev = _lib.require_cuda(perm)
    inv = torch.empty_like(perm)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mma_invert_perm(_lib.ptr(perm), perm.numel(), _lib.ptr(inv),
                                              _lib.stream_ptr(dev)), "mma_invert_perm")
    return inv


class Graph:
    """CSR by destination + its transpose for one edge list.

    rowptr [n_dst+1], col [E] (source of each slot), perm [E] (original edge id of each slot);
    colptr [n_src+1], row_t [E] (destination of each transposed slot), perm_t [E];
    csr2csc [E]: CSR slot -> CSC slot (row of the per-edge gradient buffer in the backward).
    `identity_perm` is True when the input was already sorted by destination.
    """

    def __init__(self, src: Tensor, dst: Tensor, n_dst: int, n_src: Optional[int] = None,
                 need_transpose: bool = True, sort_rows: bool = False):
        dev = _lib.require_cuda(src, dst)
        self.device = dev
        self.E = int(dst.numel())
        self.n_dst = int(n_dst)
        self.n_src = int(n_src if n_src is not None else n_dst)
        self.rng_row, self.rng_row0 = None, 0      # id of CSR row r in K1's dropout stream: rng_row0 + rng_row[r]
        self.row_map = None                        # node id of each CSR row (None = identity)
        self.row_rank = None                       # CSR row of each node (inverse of row_map)
        self.buckets = None                        # [(degree, row_lo, row_hi)] when rows are degree-sorted
        if sort_rows and self.n_dst > 0:
            # CSR rows in order of decreasing in-degree: nodes of equal degree become contiguous row
            # ranges (the post-transform then needs ONE effective weight per range, see
            # functional.scaled_post) and long rows are scheduled first.  In-row edge order is
            # untouched (stable sort by the relabelled destination).
            deg = torch.bincount(dst, minlength=self.n_dst)
            order = torch.argsort(deg, descending=True, stable=True)
            rank = torch.empty_like(order)
            rank[order] = torch.arange(self.n_dst, device=dev)
            vals, counts = torch.unique_consecutive(deg.index_select(0, order), return_counts=True)
            hi = torch.cumsum(counts, 0)
            vals, counts, hi = vals.tolist(), counts.tolist(), hi.tolist()
            self.buckets = [(int(d), int(h - c), int(h)) for d, c, h in zip(vals, counts, hi)]
            self._max_deg_hint = int(vals[0]) if vals else 0
            self.row_map = order.to(torch.int32).contiguous()
            self.rng_row = self.row_map                # stream keyed by the ORIGINAL node id: layout-invariant
            self.row_rank = rank
            dst = rank.index_select(0, dst)
        self.rowptr, self.col, self.perm = csr_build(dst, src, self.n_dst)
        self.gid, self.E_total = None, self.E      # set by the partitioner for a shard
        self._src, self._dst = src, dst
        self._t_built = False
        self.colptr = self.row_t = self.perm_t = self.csr2csc = None
        self._max_deg = None
        if need_transpose:
            self.build_transpose()

    def build_transpose(self):
        if self._t_built:
            return
        self.colptr, self.row_t, self.perm_t = csr_build(self._src, self._dst, self.n_src)
        if self.E > 0:
            inv_t = invert_perm(self.perm_t)                       # original edge id -> CSC slot
            self.csr2csc = inv_t.index_select(0, self.perm.to(torch.int64)).contiguous()
        else:
            self.csr2csc = torch.empty(0, dtype=torch.int32, device=self.device)
        self._t_built = True
        self._src = self._dst = None

    @property
    def max_deg(self) -> int:
        """Largest in-degree (one host sync, cached): sizes the scaler lookup table."""
        if self._max_deg is None:
            if self.n_dst == 0 or self.E == 0:
                self._max_deg = 0
            else:
                self._max_deg = int((self.rowptr[1:] - self.rowptr[:-1]).max().item())
        return self._max_deg

    K1_ROW_COST = 6          # a row costs about this many edges in the K1 kernels (epilogue, per-row loads)
    K1_CHUNKS_PER_WARP = 16

    K1_MAX_SEG = 4096        # rows longer than this are cut into segments of this many edges (multiple of 32)

    def k1_segments(self):
        """Virtual rows for the persistent K1 kernels (include/mma_b200.h, `vrowptr` / `seg_tab` / `split_tab`):
        rows longer than K1_MAX_SEG edges are cut into segments walked by different warps.  None when no row
        is that long (one host sync for the maximum degree, once per graph)."""
        if "_k1_seg" in self.__dict__:
            return self.__dict__["_k1_seg"]
        seg = None
        L = int(self.K1_MAX_SEG)
        if L % 32 != 0 or L < 32:
            raise ValueError("K1_MAX_SEG must be a positive multiple of 32")
        if self.max_deg > L:
            dev, n = self.device, self.n_dst
            rp = self.rowptr.to(torch.int64)
            deg = rp[1:] - rp[:-1]
            segs = ((deg + L - 1) // L).clamp_(min=1)
            first_v = torch.cumsum(segs, 0) - segs
            vrow_row = torch.repeat_interleave(torch.arange(n, device=dev), segs)
            n_v = int(vrow_row.numel())
            pos0 = (torch.arange(n_v, device=dev) - first_v[vrow_row]) * L
            split = segs[vrow_row] > 1
            slot = torch.where(split, torch.cumsum(split.to(torch.int64), 0) - 1, torch.full_like(pos0, -1))
            zeros = torch.zeros_like(pos0)
            rows_split = torch.nonzero(segs > 1).flatten()
            seg = type("K1Segments", (), {})()
            seg.vrowptr = torch.cat([rp[vrow_row] + pos0, rp[-1:]]).to(torch.int32).contiguous()
            seg.seg_tab = torch.stack([vrow_row, pos0, slot, zeros], 1).to(torch.int32).contiguous()
            seg.split_tab = torch.stack([rows_split, slot[first_v[rows_split]], segs[rows_split],
                                         torch.zeros_like(rows_split)], 1).to(torch.int32).contiguous()
            seg.n_vrows, seg.n_split, seg.n_slots = n_v, int(rows_split.numel()), int(split.sum().item())
        self.__dict__["_k1_seg"] = seg
        return seg

    def k1_chunks(self) -> Optional[Tensor]:
        """Work partition of the persistent K1 kernels (include/mma_b200.h, `row_chunks`): int32
        [n_chunks + 1] boundaries of chunks of (virtual, see k1_segments) rows of about equal cost (edges +
        K1_ROW_COST per row), K1_CHUNKS_PER_WARP chunks per resident warp, dealt round-robin in the kernel.
        Built once per graph."""
        ch = self.__dict__.get("_k1_chunks")
        if ch is None:
            seg = self.k1_segments()
            n = self.n_dst if seg is None else seg.n_vrows
            rowptr = self.rowptr if seg is None else seg.vrowptr
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            n_chunks = max(1, min(sms * 16 * self.K1_CHUNKS_PER_WARP, n // 8))
            cost = rowptr.to(torch.int64) + self.K1_ROW_COST * torch.arange(n + 1, device=self.device)
            targets = (torch.arange(n_chunks + 1, device=self.device, dtype=torch.int64) * int(cost[-1].item())) // n_chunks
            ch = torch.searchsorted(cost, targets, right=False).clamp_(max=n).to(torch.int32)
            ch[0], ch[-1] = 0, n
            ch = self.__dict__["_k1_chunks"] = ch.contiguous()
        return ch

    @staticmethod
    def from_edge_index(edge_index: Tensor, num_nodes: int, need_transpose: bool = True,
                        sort_rows: bool = False) -> "Graph":
        return Graph(edge_index[0], edge_index[1], num_nodes, num_nodes, need_transpose, sort_rows)

    @staticmethod
    def from_index(index: Tensor, dim_size: int) -> "Graph":
        """Segments only (MMAConv.aggregate called on its own): no sources, no transpose."""
        g = Graph.__new__(Graph)
        g.device = _lib.require_cuda(index)
        g.E = int(index.numel())
        g.n_dst = int(dim_size)
        g.n_src = 0
        g.rowptr, g.col, g.perm = csr_build(index, None, g.n_dst)
        g.gid, g.E_total = None, g.E
        g.row_map = g.row_rank = g.buckets = None
        g.rng_row, g.rng_row0 = None, 0
        g._t_built = False
        g.colptr = g.row_t = g.perm_t = g.csr2csc = None
        g._max_deg = None
        g._src = g._dst = None
        return g


_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CACHE_SIZE = 8


def cached_graph(edge_index: Tensor, num_nodes: int, sort_rows: bool = False) -> Graph:
    """Graph for an `edge_index` tensor, cached on (storage pointer, offset-free geometry, version, N).

    The entry OWNS a reference to the key tensor: while it sits in the cache the tensor's storage cannot be
    freed, so the CUDA caching allocator cannot hand the same address to another edge list (a per-step
    `torch.randint(...)` or a new mini-batch of the same shape would otherwise hit a stale entry -- the Graph
    itself drops its src/dst after build_transpose()).  A hit additionally requires the same storage
    (`untyped_storage().data_ptr()`), geometry and `_version` (bumped by any in-place write, shared by
    views).  Pass a `Graph` explicitly to the layers to bypass the cache altogether."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride()), str(edge_index.dtype),
           int(num_nodes), str(edge_index.device), bool(sort_rows))
    hit = _CACHE.get(key)
    if hit is not None:
        owner, version, g = hit
        if (owner.untyped_storage().data_ptr() == edge_index.untyped_storage().data_ptr()
                and version == edge_index._version):
            _CACHE.move_to_end(key)
            return g
        del _CACHE[key]                          # mutated in place since: rebuild
    g = Graph.from_edge_index(edge_index, num_nodes, sort_rows=sort_rows)
    _CACHE[key] = (edge_index, edge_index._version, g)
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return g


def clear_cache() -> None:
    _CACHE.clear()


class NeighbourLists:
    """CSR of the node-classification neighbour lists `add_all` (node_classification/
    utils.py:98-100: add_all[i] = column ids of row i of adj) plus its transpose.
    Edge id == CSR position, i.e. the order in which the reference consumes neighbours."""

    def __init__(self, rowptr: Tensor, col: Tensor):
        self.device = _lib.require_cuda(rowptr, col)
        self.rowptr = rowptr.to(torch.int32).contiguous()
        self.col = col.to(torch.int32).contiguous()
        self.N = int(rowptr.numel() - 1)
        self.E = int(col.numel())
        self.colptr = self.row_t = self.perm_t = None

    @staticmethod
    def from_add_all(add_all, device) -> "NeighbourLists":
        import numpy as np
        deg = np.fromiter((len(r) for r in add_all), dtype=np.int64, count=len(add_all))
        rowptr = np.zeros(len(add_all) + 1, dtype=np.int64)
        np.cumsum(deg, out=rowptr[1:])
        col = (np.concatenate([np.asarray(r, dtype=np.int64).reshape(-1) for r in add_all])
               if rowptr[-1] > 0 else np.zeros(0, dtype=np.int64))
        return NeighbourLists(torch.from_numpy(rowptr).to(device), torch.from_numpy(col).to(device))

    def build_transpose(self):
        if self.colptr is None:
            deg = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
            row = torch.repeat_interleave(torch.arange(self.N, device=self.device, dtype=torch.int64), deg)
            self.colptr, self.row_t, self.perm_t = csr_build(self.col.to(torch.int64), row, self.N)


class SparseAdj:
    """CSR (+ transposed CSR) of a torch sparse COO adjacency for K3 (torch.spmm replacement,
    node_classification/layers.py:41,862)."""

    def __init__(self, adj: Tensor):
        if adj.layout != torch.sparse_coo:
            raise RuntimeError("adj must be a torch sparse COO tensor (node_classification/utils.py:139-146)")
        adj = adj.coalesce()
        idx, val = adj.indices(), adj.values().to(torch.float32)
        self.device = _lib.require_cuda(idx, val)
        self.n_rows, self.n_cols = int(adj.shape[0]), int(adj.shape[1])
        self.rowptr, self.col, perm = csr_build(idx[0], idx[1], self.n_rows)
        self.val = val.index_select(0, perm.to(torch.int64)).contiguous()
        self.colptr, self.row_t, perm_t = csr_build(idx[1], idx[0], self.n_cols)
        self.val_t = val.index_select(0, perm_t.to(torch.int64)).contiguous()


_ADJ_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()


def cached_adj(adj: Tensor) -> SparseAdj:
    """SparseAdj of a sparse COO adjacency, cached.  As in `cached_graph` the entry keeps the caller's tensor
    alive (so its index/value storages cannot be recycled for another adjacency) and a hit requires the same
    storages and versions; a non-coalesced `adj` is keyed by its OWN raw indices/values, never by the
    temporaries of `coalesce()`."""
    idx, val = adj._indices(), adj._values()
    key = (idx.data_ptr(), val.data_ptr(), int(adj._nnz()), tuple(adj.shape), str(adj.device), bool(adj.is_coalesced()))
    hit = _ADJ_CACHE.get(key)
    if hit is not None:
        owner, versions, s = hit
        if (owner._indices().data_ptr() == idx.data_ptr() and owner._values().data_ptr() == val.data_ptr()
                and versions == (idx._version, val._version)):
            _ADJ_CACHE.move_to_end(key)
            return s
        del _ADJ_CACHE[key]
    s = SparseAdj(adj)
    _ADJ_CACHE[key] = (adj, (idx._version, val._version), s)
    while len(_ADJ_CACHE) > 8:
        _ADJ_CACHE.popitem(last=False)
    return s
