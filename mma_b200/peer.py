"""Peer-memory exchange of the destination-range sharded layer: one process per GPU on ONE node, every rank's
exchange buffers mapped into every other rank's address space (CUDA IPC over NVLink 5 / NVSwitch), payload moved by
the COPY ENGINES, arrival announced by an 8-byte copy into the receiver's flag array (csrc/peer_exchange.cu).

Why not NCCL for these two exchanges (north_star: "halo all-gather ... overlapped with interior aggregation"): an NCCL
collective is a kernel that needs SMs; the aggregation kernels and the tcgen05 GEMMs are persistent one-CTA-per-SM
kernels that own the whole register file, so a concurrent NCCL kernel either waits or displaces a CTA into a second
wave (measured in round 1: the overlap cancelled itself).  Copy-engine transfers need no SM at all, so the transfer of
feature window k+1 really runs under the aggregation of window k.  torch.distributed (NCCL) stays the plumbing: group
set-up, handle exchange, weight-gradient all-reduce.

The reference has no distributed code; this is new and follows BASELINE.json's north_star.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib

N_PHASES = 16
PHASE_ENTER = 15                # "this rank has entered the call": every earlier reader of its receive buffers is done
WAIT_TIMEOUT_NS = 20_000_000_000


class _RawCuda:
    """A raw device pointer presented through __cuda_array_interface__ so torch can alias it as a tensor."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2, "strides": None}


class SharedRegion:
    """One cudaMalloc allocation of this rank (zero-filled, made by the library: a CUDA IPC handle names a whole
    allocation) mapped into every rank of the group ON THAT RANK'S OWN DEVICE, i.e. as peer memory its kernels and
    copy engines reach over NVLink.  `base[r]` is the address of rank r's region in THIS process; `local` aliases this
    rank's region as a uint8 tensor.  Collective: all ranks must construct it together with the same size."""

    def __init__(self, nbytes: int, device, group=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = int(nbytes)
        self.device = torch.device(device)
        lib = _lib.lib()
        handle = (C.c_ubyte * 64)()
        p = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.mma_peer_alloc(self.nbytes, C.byref(p), handle), "mma_peer_alloc")
        self._own = int(p.value)
        handles: List = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.base: List[int] = []
        self._opened: List[int] = []
        for r in range(self.world):
            if r == self.rank:
                self.base.append(self._own)
                continue
            q = C.c_void_p()
            hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
            with torch.cuda.device(self.device):
                _lib.check(lib.mma_peer_open(hb, C.byref(q)), "mma_peer_open")
            self.base.append(int(q.value))
            self._opened.append(int(q.value))
        self.local = torch.as_tensor(_RawCuda(self._own, self.nbytes), device=self.device)
        dist.barrier(group)             # nobody proceeds before everyone has opened the handles

    def view(self, offset: int, shape, dtype) -> Tensor:
        """Tensor view of this rank's region at a byte offset."""
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.local[offset: offset + nbytes].view(dtype).view(*shape)

    def close(self) -> None:
        """Unmaps the peers' regions and frees this rank's (call on all ranks, after a barrier, when nothing is in flight)."""
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            for q in self._opened:
                lib.mma_peer_close(q)
            self._opened = []
            if self._own:
                self.local = None
                lib.mma_peer_free(self._own)
                self._own = 0


class PeerExchange:
    """Buffers, flags and the copy stream of the two exchanges for ONE (rows, width) geometry.

    recv_q   [world * max_rows, F]     gathered Q (rank-major, padded): every rank pushes its rows, feature window by
                                       feature window (strided 2-D copies), into every peer's copy; K1 reads it with
                                       its col0 / ncols ABI as soon as a window is complete
    slices   [world][max_rows, F]      partial-dQ slices as they arrive at their owner, summed in rank order
    flags    [N_PHASES][world] uint64  flags[phase][sender] = epoch * 16 + phase of the sender's last announcement

    ONE copy stream: on an 8-GPU NVSwitch box concurrent copies to different peers from several streams were SLOWER
    (434 GB/s with 4 streams, 414 with 7) than the same copies back to back on one stream in the staggered order
    rank+1, rank+2, ... (590 GB/s per GPU per direction; 766 GB/s between 2 GPUs) -- profiles/r2i_peer_probe_8gpu.log.
    """

    PHASE_BWD = 8

    def __init__(self, world: int, rank: int, max_rows: int, F: int, windows: Sequence[slice], device, group=None,
                 n_copy_streams: int = 1):
        self.world, self.rank, self.max_rows, self.F, self.group = world, rank, int(max_rows), int(F), group
        self.windows = list(windows)
        self.dev = device
        if len(self.windows) > self.PHASE_BWD:
            raise ValueError("too many feature windows for the phase space of the flags")
        self.widths = [s.stop - s.start for s in self.windows]
        tot = world * self.max_rows
        # ONE shared allocation per rank: [ recv_q | slices | flags ]
        q_elems = tot * self.F
        self._off_recv, self._off_slice, self._off_flags = 0, 4 * q_elems, 8 * q_elems
        self.region = SharedRegion(8 * q_elems + 8 * N_PHASES * world, device, group)
        self.recv_q = self.region.view(self._off_recv, (tot, self.F), torch.float32)
        self.slices = self.region.view(self._off_slice, (world, self.max_rows, self.F), torch.float32)
        self._flags = self.region.view(self._off_flags, (N_PHASES, world), torch.int64)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        self.vals = torch.zeros(N_PHASES, dtype=torch.int64, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.streams = [torch.cuda.Stream(device=device, priority=-1) for _ in range(max(1, min(n_copy_streams, world)))]
        self.fwd_calls = 0          # host-side count of forward calls (guards the gathered Q a backward re-reads)
        self._own_part: Optional[Tensor] = None     # partial-dQ tensor whose own-rank block sum_slices() reads in place
        torch.cuda.synchronize(device)
        dist.barrier(group)

    # ---------------------------------------------------------------- low level
    def _copy(self, dst_ptr: int, src_ptr: int, nbytes: int, stream: torch.cuda.Stream) -> None:
        _lib.check(_lib.lib().mma_peer_copy(dst_ptr, src_ptr, nbytes, stream.cuda_stream), "mma_peer_copy")

    def _stream_for(self, peer: int) -> torch.cuda.Stream:
        return self.streams[((peer - self.rank) % self.world) % len(self.streams)]

    def begin_call(self) -> None:
        """On the current stream: epoch += 1, announcement values for all phases, and the entry barrier's own
        announcement (every earlier reader of this rank's receive buffers ran before this point in stream order)."""
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().mma_peer_epoch_advance(self.epoch.data_ptr(), self.vals.data_ptr(), cur.cuda_stream),
                       "mma_peer_epoch_advance")
        ev = torch.cuda.Event()
        ev.record(cur)
        for st in self.streams:
            st.wait_event(ev)
        self._announce(PHASE_ENTER, range(self.world))

    def _announce(self, phase: int, peers) -> None:
        """8-byte copy of vals[phase] into flags[phase][my rank] of each peer, on that peer's copy stream."""
        src = self.vals.data_ptr() + 8 * phase
        for o in peers:
            dst = self.region.base[o] + self._off_flags + 8 * (phase * self.world + self.rank)
            with torch.cuda.device(self.dev):
                self._copy(dst, src, 8, self._stream_for(o))

    def wait(self, phase: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        """`stream` (default: current) continues once every rank's announcement of `phase` for this epoch is here.
        On the compute stream the wait kernel's duration IS the exposed (un-hidden) part of that exchange: it is
        counted and timed like every other launch (bench.py lists it as peer_wait_q / peer_wait_dq)."""
        def launch(st):
            _lib.check(_lib.lib().mma_peer_wait(self.epoch.data_ptr(), self._flags.data_ptr() + 8 * phase * self.world,
                                                self.world, phase, WAIT_TIMEOUT_NS, self.err.data_ptr(), st.cuda_stream),
                       "mma_peer_wait")
        if stream is not None:
            with torch.cuda.device(self.dev):
                launch(stream)
            return
        name = "peer_wait_dq" if phase == self.PHASE_BWD else ("peer_wait_enter" if phase == PHASE_ENTER else "peer_wait_q")
        with _lib.kernel_scope(name, self.dev):
            launch(torch.cuda.current_stream(self.dev))

    def _peer_order(self):
        """Peers in sending order: rank+1, rank+2, ... (at any moment the ranks address distinct destinations), this
        rank's own copy last."""
        return [(self.rank + 1 + k) % self.world for k in range(self.world)]

    owner_order = _peer_order

    # ---------------------------------------------------------------- forward: all-gather of Q, window by window
    def push_q(self, Q: Tensor, rows: int) -> None:
        """Q [rows, F] (unit column stride; may be a column view of a wider matrix) goes into rows [rank * max_rows,
        +rows) of every rank's recv_q; phase k announces window k.  One window (the default): a local copy makes the
        source contiguous, the peer copies are 1-D.  Several windows: the columns of window k go as strided 2-D copies
        straight out of Q.  The copy stream first waits for all ranks' entry announcement (their receive buffers are
        free)."""
        if Q.stride(1) != 1 or Q.dtype != torch.float32 or Q.shape[1] != self.F:
            raise RuntimeError("push_q: Q must be fp32 [rows, F] with unit column stride")
        cur = torch.cuda.current_stream(self.dev)
        lib = _lib.lib()
        own_ptr = self.region.base[self.rank] + self._off_recv + 4 * self.rank * self.max_rows * self.F
        staged = len(self.windows) == 1 and Q.stride(0) != self.F
        if staged:
            # Q is a column block of a wider matrix ([P | Q | XW] of the projection GEMM): pushed as it lies, every peer
            # copy would be a 2-D copy of 512-byte rows, which the copy engines move at ~440 GB/s.  One LOCAL strided copy
            # into this rank's own block of its gathered buffer (where K1 reads it anyway) makes the source contiguous:
            # the seven peer copies then are 1-D and run at the engines' full rate (590 GB/s per GPU at 8 GPUs).
            with torch.cuda.device(self.dev):
                _lib.check(lib.mma_peer_copy_2d(own_ptr, self.F * 4, Q.data_ptr(), Q.stride(0) * 4, self.F * 4, rows,
                                                cur.cuda_stream), "mma_peer_copy_2d")
        ev = torch.cuda.Event()
        ev.record(cur)
        for st in self.streams:
            st.wait_event(ev)
            self.wait(PHASE_ENTER, st)
        spitch, dpitch = Q.stride(0) * 4, self.F * 4
        full = len(self.windows) == 1 and (staged or Q.stride(0) == self.F)
        for k, (s_, w) in enumerate(zip(self.windows, self.widths)):
            src = (own_ptr if staged else Q.data_ptr()) + 4 * s_.start
            for o in self._peer_order():
                if staged and o == self.rank:
                    continue                                # already in place
                dst = self.region.base[o] + self._off_recv + 4 * (self.rank * self.max_rows * self.F + s_.start)
                with torch.cuda.device(self.dev):
                    if full:
                        self._copy(dst, src, rows * self.F * 4, self._stream_for(o))
                    else:
                        _lib.check(lib.mma_peer_copy_2d(dst, dpitch, src, spitch, w * 4, rows,
                                                        self._stream_for(o).cuda_stream), "mma_peer_copy_2d")
            self._announce(k, self._peer_order())
        for st in self.streams:
            Q.record_stream(st)

    # ---------------------------------------------------------------- backward: reduce-scatter of the partial dQ
    def push_partial_block(self, o: int, part: Tensor) -> None:
        """part [world * max_rows, F] (contiguous) holds this rank's partial dQ; its rows of OWNER o
        (rows [o * max_rows, +max_rows)) are final once the kernels enqueued so far on the current stream have run:
        they go into slices[my rank] of rank o, and PHASE_BWD announces them to o.  Calling it block by block, right
        after the transpose pass of each owner's source rows, lets the slices leave while the next block is still
        being summed."""
        cur = torch.cuda.current_stream(self.dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        st = self._stream_for(o)
        st.wait_event(ev)
        if o == self.rank:
            # this rank's own block never travels: sum_slices() reads it where it lies (`own`), and only the flag is
            # set -- one 1/world-sized local copy less at the tail of the exchange, which is the critical path
            self._own_part = part
            self._announce(self.PHASE_BWD, [o])
            return
        nbytes = self.max_rows * self.F * 4
        dst = self.region.base[o] + self._off_slice + 4 * self.rank * self.max_rows * self.F
        src = part.data_ptr() + 4 * o * self.max_rows * self.F
        with torch.cuda.device(self.dev):
            self._copy(dst, src, nbytes, st)
        self._announce(self.PHASE_BWD, [o])
        part.record_stream(st)

    def push_partial(self, part: Tensor) -> None:
        """All owners' blocks at once (the whole `part` is final)."""
        for o in self._peer_order():
            self.push_partial_block(o, part)

    def sum_slices(self, out: Tensor, rows: int) -> None:
        """out (a [rows, F] view with unit column stride) = sum over ranks, ascending, of the arrived slices -- after
        waiting for PHASE_BWD on the current stream."""
        self.wait(self.PHASE_BWD)
        where = [self.slices[r].data_ptr() for r in range(self.world)]
        own = self._own_part
        if own is not None:                                 # this rank's own partial block, in place (push_partial_block)
            where[self.rank] = own.data_ptr() + 4 * self.rank * self.max_rows * self.F
        ptrs = (C.c_void_p * self.world)(*where)
        with _lib.kernel_scope("mma_sum_slices", self.dev):
            _lib.check(_lib.lib().mma_sum_slices(ptrs, self.world, rows, self.F, out.data_ptr(), out.stride(0),
                                                 _lib.stream_ptr(self.dev)), "mma_sum_slices")
        self._own_part = None                               # the summing kernel is enqueued: stream order keeps the block alive

    def join(self) -> None:
        """The current stream waits for everything enqueued on the copy streams (end of a call / of a capture)."""
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            cur.wait_stream(st)

    def check(self) -> None:
        """Host-side: raises if a wait gave up (a peer never announced).  Synchronises."""
        torch.cuda.synchronize(self.dev)
        e = int(self.err.item())
        if e:
            raise _lib.MMAError(f"peer exchange: rank {self.rank} timed out waiting for rank {e - 1}")


_EXCHANGES: Dict[tuple, PeerExchange] = {}


def exchange_for(sg, F: int, windows: Sequence[slice], device, owner: int = 0) -> PeerExchange:
    """The PeerExchange of (ShardedGraph, layer `owner`, F, windows); built collectively on first use (every rank must
    reach this call), cached on the graph.  Every layer owns its buffers: the gathered Q of its forward has to survive
    until its own backward, whatever other layers run in between."""
    key = (int(owner), int(F), tuple((s.start, s.stop) for s in windows))
    cache = sg.__dict__.setdefault("_peer_exchanges", {})
    ex = cache.get(key)
    if ex is None:
        ex = cache[key] = PeerExchange(sg.world, sg.rank, sg.max_rows, F, windows, device, sg.group)
    return ex
