"""Copy-engine peer-to-peer bandwidth and the exchange protocol of mma_b200/peer.py, at the sizes of the sharded layer.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/peer_probe.py
Prints: per-GPU push bandwidth to one / all peers with 1..8 copy streams (contiguous blocks and strided 2-D windows),
the same bytes through NCCL all-gather / reduce-scatter, and the pipelined exchange (push_q -> wait -> read) checked
against NCCL's result, eager and replayed from a CUDA graph."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from mma_b200 import _lib
from mma_b200.peer import PeerExchange, SharedRegion
from mma_b200.parallel import _slices

N, F = 2_000_000, 128
rows = N // world
log = lambda *a: print(*a, flush=True) if rank == 0 else None


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


# ---- raw copy-engine bandwidth: this rank pushes `rows x F` floats to every peer
src = torch.randn(rows, F, device=dev)
region = SharedRegion(world * rows * F * 4, dev)
peers = region.base                      # address of every rank's receive buffer in THIS process
lib = _lib.lib()
main = torch.cuda.current_stream()
for n_streams in (1, 2, 4, 7):
    if n_streams > max(1, world - 1) and n_streams != 1:
        continue
    sts = [torch.cuda.Stream(priority=-1) for _ in range(n_streams)]

    def push_all(chunks=1):
        ev = torch.cuda.Event(); ev.record(main)
        for st in sts:
            st.wait_event(ev)
        nb = rows * F * 4 // chunks
        for c in range(chunks):
            for k in range(1, world):
                o = (rank + k) % world
                st = sts[(k - 1) % n_streams]
                _lib.check(lib.mma_peer_copy(peers[o] + 4 * rank * rows * F + c * nb, src.data_ptr() + c * nb, nb,
                                             st.cuda_stream), "copy")
        for st in sts:
            main.wait_stream(st)
    for chunks in (1, 4):
        ms = timeit(lambda: push_all(chunks))
        gb = rows * F * 4 * (world - 1) / 1e9
        log(f"copy engines: world {world}, {n_streams} stream(s), {chunks} chunk(s)/peer: {gb:.2f} GB out per GPU in {ms:.3f} ms "
            f"= {gb / ms * 1e3:.0f} GB/s per GPU per direction")
# strided 2-D window (no pack): Q window of 32 / 64 columns out of a 384-column row
pqx = torch.randn(rows, 384, device=dev)
sts = [torch.cuda.Stream(priority=-1) for _ in range(4)]
for w in (32, 64):
    def push2d():
        ev = torch.cuda.Event(); ev.record(main)
        for st in sts:
            st.wait_event(ev)
        for k in range(1, world):
            o = (rank + k) % world
            _lib.check(lib.mma_peer_copy_2d(peers[o] + 4 * rank * rows * w, w * 4, pqx.data_ptr() + 4 * 128, 384 * 4,
                                            w * 4, rows, sts[(k - 1) % 4].cuda_stream), "copy2d")
        for st in sts:
            main.wait_stream(st)
    ms = timeit(push2d)
    gb = rows * w * 4 * (world - 1) / 1e9
    log(f"copy engines, strided 2-D window of {w} columns (pitch 1536 B): {gb:.3f} GB in {ms:.3f} ms = {gb / ms * 1e3:.0f} GB/s")
# NCCL, same bytes
out = torch.empty(world * rows, F, device=dev)
big = torch.randn(world * rows, F, device=dev); rs = torch.empty(rows, F, device=dev)
t_ag = timeit(lambda: dist.all_gather_into_tensor(out, src))
t_rs = timeit(lambda: dist.reduce_scatter_tensor(rs, big))
gb = rows * F * 4 * (world - 1) / 1e9
log(f"NCCL: all_gather {t_ag:.3f} ms ({gb / t_ag * 1e3:.0f} GB/s rx per GPU), reduce_scatter {t_rs:.3f} ms ({gb / t_rs * 1e3:.0f} GB/s)")
del out, big
torch.cuda.synchronize(); dist.barrier(); region.close()
torch.cuda.empty_cache()

# ---- the protocol: push_q / wait / push_partial / sum_slices against NCCL, eager and graph replay
for n_win in (1, 2, 4):
    wins = _slices(F, n_win)
    ex = PeerExchange(world, rank, rows, F, wins, dev)
    Q = torch.randn(rows, 3 * F, device=dev)[:, F:2 * F]            # strided view like PQX[:, F:2F]
    ref = torch.empty(world * rows, F, device=dev)
    part = torch.randn(world * rows, F, device=dev)
    dq = torch.zeros(rows, 2 * F, device=dev)
    sink = torch.zeros(world * rows, F, device=dev)

    def call():
        ex.begin_call()
        ex.push_q(Q, rows)
        for k, s in enumerate(wins):
            ex.wait(k)
            sink[:, s] = ex.recv_q[:, s]                              # a consumer of window k on the compute stream
        ex.join()
        ex.begin_call()
        for o in ex.owner_order():
            ex.push_partial_block(o, part)
        ex.sum_slices(dq[:, F:], rows)
        ex.join()

    def check(tag):
        dist.all_gather_into_tensor(ref, Q.contiguous())
        ok_q = torch.equal(sink, ref)
        want = torch.empty(rows, F, device=dev)
        dist.reduce_scatter_tensor(want, part)
        err = float((dq[:, F:] - want).abs().max() / want.abs().max())
        ex.check()
        t = torch.tensor([float(ok_q), -err], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
        log(f"protocol, {n_win} window(s), {tag}: gathered Q identical to NCCL's on every rank: {bool(t[0] > 0)}; "
            f"summed slices vs NCCL reduce-scatter rel err {-float(t[1]):.2e}")
    call(); check("eager")
    ms = timeit(call)
    log(f"protocol, {n_win} window(s): eager all-gather + reduce-scatter of 2 x {gb:.2f} GB: {ms:.3f} ms per call")
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        call()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        call()
    for it in range(2):
        Q.normal_(); part.normal_(); sink.zero_(); dq.zero_()
        torch.cuda.synchronize(); dist.barrier()
        g.replay()
        check(f"CUDA graph replay {it}")
    ms = timeit(g.replay)
    log(f"protocol, {n_win} window(s): graph replay {ms:.3f} ms per call")
    g.reset(); torch.cuda.synchronize(); dist.barrier(); ex.region.close(); del g, ex
    torch.cuda.empty_cache()
torch.cuda.synchronize(); dist.barrier()
sys.stdout.flush()
os._exit(0)
