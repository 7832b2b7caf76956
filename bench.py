#!/usr/bin/env python
"""Benchmark of the MultiMaskConv hot path (BASELINE.json: "MultiMaskConv fwd+bwd edges/sec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c1|c2|c3|c4|c5] [--results F]

One step = ONE layer call, forward + backward (x -> out, all gradients), fp32.  The default workload is BASELINE
config 4, the one the metric is quoted on: synthetic uniform random graph, 2M nodes / 32M edges, hidden 128,
aggregators mean,sum,min,max,std x scalers identity,amplification,attenuation,linear, always-on dropout 0.5 (the
reference's behaviour, Q3).  N > 1: the graph is partitioned by destination range (strong scaling, one exchange per
direction, mma_b200/parallel.py); launched by `python -m torch.distributed.run --nproc-per-node N bench.py --gpus N`.
The other configurations (parity-test cases, not the bench line) are reachable with --config:
  c1  Cora topology, node-classification mask layer `MMA`, aggregators mean,mean2, hidden 64 -> 7 classes
  c2  ZINC-shaped batch of 128 graphs, MMAConv min,max x identity,amplification,linear, towers 5, edge features
      (N > 1: the batch is split by whole graphs, data parallel)
  c3  Pubmed topology, `MMA` with min,min2,min3,min4, hidden 16 -> 3 classes
  c5  power-law graph 10M nodes / 200M edges, hidden 64 (8 GPUs; destination ranges balanced by in-edge count)

Rank 0 prints ONE JSON line.  `value` = E / (device time per step), max over ranks.  `e2e` is the same call at the host
boundary of the reference's loop: x uploaded from pinned host memory and the loss read back every step (a second figure,
`e2e.with_y_readback`, also copies the layer output to the host every step).
`roofline` follows SURVEY.md 8(d): algorithmic bytes of the dominant part of the aggregate op (the forward kernel, or
the backward = destination pass + transpose pass) over its CUDA-event time, against the measured copy peak.
`cpu_baseline` is the oracle port of the reference layer (oracle/restate.py) timed on this box's host cores on a bounded
sample.  `--impl reference` prints the same line for that CPU path alone.  N > 1 also prints `parity_check`: every rank
re-runs the step on ONE GPU over the full graph with the same dropout key and compares its rows (outside the timed
region; --no-verify skips it).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mma_b200.synthetic import (powerlaw_edges, uniform_edges, zinc_like_batch, degree_histogram,  # noqa: E402,F401
                                csr_to_sparse_adj)

AGGR = ["mean", "sum", "min", "max", "std"]
SCAL = ["identity", "amplification", "attenuation", "linear"]
CONFIGS = {
    # name: (nodes, edges, hidden)
    "c4": (2_000_000, 32_000_000, 128),
    "c4s": (125_000, 2_000_000, 128),      # a 1/16 sub-graph, for quick checks
    # config 5 (8-GPU load-balance stress, not the bench line): power-law in-degrees (alpha = 2.1, capped at 10^6),
    # hidden 64, destination ranges balanced by in-edge count
    "c5": (10_000_000, 200_000_000, 64),
    "c5s": (1_250_000, 25_000_000, 64),    # one eighth of it, for dry runs
}
SMALL = {
    # name: (kind, topology / graphs, hidden, classes, aggregators, dropout)
    "c1": ("nc", "cora", 64, 7, ["mean", "mean2"], 0.75),
    "c3": ("nc", "pubmed", 16, 3, ["min", "min2", "min3", "min4"], 0.5),
    "c2": ("zinc", 128, 75, None, ["min", "max"], 0.5),
}
C2_SCAL = ["identity", "amplification", "linear"]
NC_ORDER = ["moment_3", "sum", "sum2", "sum3", "sum4", "mean", "mean2", "mean3", "mean4", "max", "max2", "max3",
            "max4", "min", "min2", "min3", "min4", "softmax", "softmin", "std", "normalized_mean"]
METRIC = "MultiMaskConv fwd+bwd edges/sec"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")     # ncu dram bytes per launch of the committed capture
CPU_SEC_PER_EDGE = 3.7e-6       # oracle port, config-4 layer, 16 host threads (7.3 s per 2M-edge fwd+bwd): sizes the sample


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# algorithmic bytes (int32 indices, fp32 data, one gathered row per edge,
# no L2 credit, only tensors materialised at the op boundary)
# ------------------------------------------------------------------------------------------
def algo_bytes(N, E, F, A, S, n_mm, std):
    """A, S: aggregator / scaler blocks MATERIALISED at the op boundary (S = 1 when the scalers are folded into the
    post GEMM; a `mean` folded into the `sum` block of the per-degree weight is not materialised either)."""
    fwd = 4 * (N + 1) + 4 * E + 4 * F * E + 4 * F * N + 4 * F * A * S * N + 4 * F * n_mm * N
    bwd = (4 * F * A * S * N + 4 * F * n_mm * N + 8 * F * N + 8 * (N + 1) + 8 * E + 4 * E + 4 * F * E
           + (4 * F * E if std else 0))
    # per kernel of THIS implementation (its own contract: inputs once, outputs once) -- DRAM-utilisation figures
    k_fwd = fwd + (8 * F * N if std else 0)                         # + saved mean/var
    k_dst = (4 * F * A * S * N + 4 * F * n_mm * N + 4 * F * N       # dY, args, dP
             + 4 * (N + 1) + 4 * E + 4 * E + 4 * E                  # rowptr, col, perm, csr2csc
             + (4 * F * E + 4 * F * N + 8 * F * N if std else 0)    # re-gather Q, P, mean/var
             + 4 * F * E)                                           # per-edge gradient rows (write)
    k_src = 4 * (N + 1) + 4 * F * E + 4 * F * N                     # colptr, G rows (read), dQ
    # dense projections of the fused layer (fused_layer.py): activations once in, once out; weights negligible
    Fo, K = F, A * F
    g = {"gemm_mask_proj": 4 * N * (F + 2 * F + Fo), "gemm_post_grouped": 4 * N * (K + Fo + Fo),
         "gemm_lin": 4 * N * (Fo + F), "gemm_lin_dgrad": 4 * N * (F + Fo), "gemm_lin_wgrad": 4 * N * (F + Fo),
         "gemm_post_dgrad": 4 * N * (Fo + K), "gemm_post_wgrad": 4 * N * (Fo + K),
         "gemm_mask_dgrad": 4 * N * (2 * F + Fo + F), "gemm_mask_wgrad": 4 * N * (2 * F + Fo + F)}
    out = {"fwd": fwd, "bwd": bwd, "mmconv_aggregate_fwd": k_fwd, "mmconv_aggregate_bwd_dst": k_dst,
           "mma_segment_sum_rows": k_src}
    out.update(g)
    return out


def nc_algo_bytes(N, E, F, A):
    """Node-classification flavour (K2), SURVEY 8(d): fwd = 4(N+1) + 4E + 4F(A+1)E + 4F(2A+1)N; bwd ~ 2x that plus
    4F(2A+1)N of outputs."""
    fwd = 4 * (N + 1) + 4 * E + 4 * F * (A + 1) * E + 4 * F * (2 * A + 1) * N
    return fwd, 2 * fwd + 4 * F * (2 * A + 1) * N


# ------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, windows=()):
        """windows: [(label, t0, t1)] in time.time() seconds, tried in order; the first one that holds a sample is
        summarised (nvidia-smi delivers ~10 samples/s, so a timed region of a few tens of ms -- 8 GPUs -- may hold
        none: the next window then adds the per-kernel timing pass and the e2e steps that follow it under the
        same load)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        picked, label = [ln for _, ln in self.lines], "whole run"
        for lab, t0, t1 in windows:
            inside = [ln for t, ln in self.lines if t0 <= t <= t1]
            if inside:
                picked, label = inside, lab
                break
        sm, mx, reasons = [], [], set()
        for ln in picked:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        hi = [v for v in sm if v >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": (hi[len(hi) // 2] if hi else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": sorted(reasons), "samples": len(sm), "window": label}


# ------------------------------------------------------------------------------------------
# CPU path: the oracle port of the reference layer on a bounded sample (cpu_baseline leg / --impl reference)
# ------------------------------------------------------------------------------------------
def cpu_c4_step_fn(N, E, F, seed=42, skewed=False):
    """restate.mmaconv_forward + backward (the reference's own op sequence: two [E,T,F] gathers, cat, edge-level
    Linear, dropout, A scatter passes, scalers, post Linears) on a graph of N nodes / E edges; returns (step, E)."""
    from oracle import restate
    g = torch.Generator().manual_seed(seed)
    if skewed:
        src, dst = powerlaw_edges(N, E, torch.device("cpu"), seed=seed)
        E = int(dst.numel())
    else:
        src = torch.randint(0, N, (E,), generator=g)
        dst = torch.randint(0, N, (E,), generator=g)
    ei = torch.stack([src, dst])
    hist = torch.bincount(torch.bincount(dst, minlength=N))
    avg = restate.avg_deg_from_hist(hist)
    A, S = len(AGGR), len(SCAL)
    lin = lambda o, i: ((torch.rand(o, i, generator=g) - 0.5) * (2 / i ** 0.5), (torch.rand(o, generator=g) - 0.5) * 0.1)
    w = restate.MMAConvWeights(F, F, AGGR, SCAL, avg, 1, F, F, False, None, [[lin(F, 2 * F)]],
                               [[lin(F, (A * S + 1) * F)]], lin(F, F))
    for t in w.tensors():
        t.requires_grad_()
    x = torch.randn(N, F, generator=g).requires_grad_()
    gy = torch.randn(N, F, generator=g)

    def step():
        y = restate.mmaconv_forward(w, x, ei, None, None, strict=False)
        torch.autograd.grad(y, [x] + w.tensors(), gy)
    return step, E


def small_inputs(name, seed=42):
    """CPU tensors of configs 1-3 (topologies: committed fixtures of the real Cora / Pubmed graphs; features and
    weights synthetic)."""
    kind, what, Fd, C, names, p = SMALL[name]
    g = torch.Generator().manual_seed(seed)
    if kind == "nc":
        topo = torch.load(os.path.join(ROOT, "tests", "golden", "planetoid_topology.pt"))[what]
        rowptr, col = topo["rowptr"], topo["col"]
        n, E = rowptr.numel() - 1, col.numel()
        d = {"kind": kind, "rowptr": rowptr, "col": col, "n": n, "E": E, "F": Fd, "C": C, "names": names, "p": p,
             "x": torch.randn(n, Fd, generator=g),
             "masks": {nm: torch.randn(2 * Fd, Fd, generator=g) * 0.1 for nm in NC_ORDER},
             "W": torch.randn(Fd, C, generator=g) * 0.1, "b": torch.zeros(C), "gy": torch.randn(n, C, generator=g)}
        d["workload"] = (f"config {name[1]}: {what.capitalize()} topology N={n} E={E}, node-classification mask layer MMA "
                         f"fwd+bwd, aggregators {','.join(names)}, hidden {Fd} -> {C} classes, dropout {p}")
        return d
    ei, batch = zinc_like_batch(what, seed=seed)
    n, E = int(batch.numel()), int(ei.shape[1])
    d = {"kind": kind, "ei": ei, "batch": batch, "n": n, "E": E, "F": Fd, "names": names, "p": p,
         "x": torch.randn(n, Fd, generator=g), "ea": torch.randn(E, 50, generator=g),
         "gy": torch.randn(n, Fd, generator=g), "hist": degree_histogram(ei, n)}
    d["workload"] = (f"config 2: ZINC-shaped batch of {what} graphs N={n} E={E}, MMAConv fwd+bwd, aggregators "
                     f"{','.join(names)} x scalers {','.join(C2_SCAL)}, towers=5, in=out=75, edge_dim=50, dropout {p}")
    return d


def cpu_small_step_fn(d, conv=None):
    """The oracle port on a small config at FULL size; `conv`: take the weights of this module (config 2)."""
    from oracle import restate
    if d["kind"] == "nc":
        names = d["names"]
        masks = {nm: d["masks"][nm].clone().requires_grad_() for nm in names}
        W, b, x = d["W"].clone().requires_grad_(), d["b"].clone().requires_grad_(), d["x"].clone().requires_grad_()
        adj = csr_to_sparse_adj(d["rowptr"], d["col"], d["n"])

        def step():
            y = restate.nc_forward(x, adj, d["rowptr"], d["col"], masks, W, b, names, "new_sigmoid", d["p"])
            torch.autograd.grad(y, [x, W, b] + [masks[nm] for nm in names], d["gy"])
        return step
    if conv is None:
        import mma_b200
        torch.manual_seed(42)
        conv = mma_b200.MMAConv(75, 75, d["names"], C2_SCAL, d["hist"], edge_dim=50, towers=5)
    w = restate.weights_from_module(conv)
    for t in w.tensors():
        t.requires_grad_()
    x = d["x"].clone().requires_grad_()

    def step():
        y = restate.mmaconv_forward(w, x, d["ei"], d["ea"], None)
        torch.autograd.grad(y, [x] + w.tensors(), d["gy"])
    return step


def time_cpu(step, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), torch.get_num_threads()


def reference_sample(E, steps, warmup, budget_s=100.0):
    """Sub-sampling factor of the CPU path for the big configs: the smallest power of two >= 16 such that
    (steps + warmup) layer calls fit the budget at ~3.7 us per edge per call."""
    sub = 16
    while sub < 1024 and (steps + warmup) * (E / sub) * CPU_SEC_PER_EDGE > budget_s:
        sub *= 2
    return sub


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path (the oracle PORT -- the reference is pure
    Python over PyG / torch_scatter, not installable here) on the host cores, rank 0 only.  Exactly `--steps` timed
    calls after `--warmup` untimed ones; the big configs run on a bounded sub-graph, named in `config.sample`."""
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.config in SMALL:
        d = small_inputs(args.config)
        step, E_run, n_run = cpu_small_step_fn(d), d["E"], d["n"]
        sample = f"full size (N={n_run}, E={E_run})"
        cfg["sample"] = {"fraction": "1/1", "nodes": n_run, "edges": E_run}
    else:
        N, E, F = CONFIGS[args.config]
        sub = reference_sample(E, steps, warmup)
        n_run = N // sub
        step, E_run = cpu_c4_step_fn(n_run, E // sub, F, skewed=args.config.startswith("c5"))
        sample = (f"1/{sub} sub-graph of the same generator (N={n_run}, E={E_run}, hidden={F}): the reference "
                  f"materialises >= 5 [E, F] tensors and cannot run the full size")
        cfg["sample"] = {"fraction": f"1/{sub}", "nodes": n_run, "edges": E_run,
                         "note": "the CPU path ran THIS sub-graph; `value` = its edges / its time per call"}
    sec, thr = time_cpu(step, steps, warmup)
    v = E_run / sec
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "edges/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": "edges/s", "cores": thr, "kind": "port",
                             "sample": f"{sample}; oracle port of the reference layer (oracle/restate.py), mean of "
                                       f"{steps} fwd+bwd after {warmup} warm-up"},
            "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return line


# ------------------------------------------------------------------------------------------
# N > 1 parity: the sharded step against the single-GPU step on the full graph, same dropout key
# ------------------------------------------------------------------------------------------
def verify_sharded(conv, sg, x, gy, params, src, dst, N, step, dist):
    """Every rank runs (a) the sharded step and (b) the single-GPU layer over the FULL graph with all ranks' x / gy
    and the same dropout key, then compares ITS rows: raw aggregates Z (all five blocks) and argmin / argmax edge ids
    bit for bit, the layer output, dx and the all-reduced weight gradients to 1e-5 (relative to the tensor's largest
    magnitude).  Eager, outside the timed region."""
    import mma_b200
    from mma_b200 import fused_layer
    dev, world, rank = x.device, sg.world, sg.rank
    F = x.shape[1]

    def gather_rows(t):
        pad = torch.zeros((sg.max_rows, t.shape[1]), dtype=t.dtype, device=dev)
        pad[: t.shape[0]] = t.detach()
        out = torch.empty((world * sg.max_rows, t.shape[1]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, pad)
        return torch.cat([out[r * sg.max_rows: r * sg.max_rows + sg.bounds[r + 1] - sg.bounds[r]] for r in range(world)])

    was_dev_seed, calls0 = conv.device_seed, conv._calls
    conv.device_seed = False
    keep = fused_layer.KEEP_LAST = {}
    y_s, dx_s = step(x)
    wg_s = [p.grad.clone() for p in params]
    Zs, amin_s, amax_s, gl = keep["Z"], keep["arg_min"], keep["arg_max"], keep["graph"]
    # (b) one GPU, full graph, same key
    x_full, gy_full = gather_rows(x).requires_grad_(), gather_rows(gy)
    gfull = mma_b200.Graph(src, dst, N, sort_rows=True)
    conv._calls = calls0
    keep2 = fused_layer.KEEP_LAST = {}
    y_f = conv(x_full, gfull)
    grads = torch.autograd.grad(y_f, [x_full] + params, gy_full)
    fused_layer.KEEP_LAST = None
    conv.device_seed = was_dev_seed
    Zf, amin_f, amax_f = keep2["Z"], keep2["arg_min"], keep2["arg_max"]
    lo, hi = sg.lo, sg.hi

    def to_nodes(t, g):          # CSR-row order -> node order
        out = torch.empty_like(t)
        out[g.row_map.long()] = t
        return out

    def gids_local(a):           # CSR slots of the shard -> global edge ids (-1: empty row)
        ok = a < gl.E
        e_loc = gl.perm.long()[a.clamp(max=max(gl.E - 1, 0)).long()]
        return torch.where(ok, gl.gid.long()[e_loc], torch.full_like(e_loc, -1))

    def gids_full(a):
        ok = a < gfull.E
        return torch.where(ok, gfull.perm.long()[a.clamp(max=gfull.E - 1).long()], torch.full_like(a.long(), -1))

    z_same = torch.equal(to_nodes(Zs, gl).view(torch.int32), to_nodes(Zf, gfull)[lo:hi].view(torch.int32))
    arg_same = True
    for a_s, a_f in ((amin_s, amin_f), (amax_s, amax_f)):
        if a_s is not None:
            arg_same &= torch.equal(to_nodes(gids_local(a_s), gl), to_nodes(gids_full(a_f), gfull)[lo:hi])
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    res = {"z_bit_identical": bool(z_same), "args_identical": bool(arg_same),
           "y_bit_identical": bool(torch.equal(y_s.detach(), y_f.detach()[lo:hi])),
           "y_max_rel_err": rel(y_s.detach(), y_f.detach()[lo:hi]), "dx_max_rel_err": rel(dx_s, grads[0][lo:hi]),
           "wgrad_max_rel_err": max(rel(a, b) for a, b in zip(wg_s, grads[1:]))}
    t = torch.tensor([float(res["z_bit_identical"]), float(res["args_identical"]), float(res["y_bit_identical"]),
                      -res["y_max_rel_err"], -res["dx_max_rel_err"], -res["wgrad_max_rel_err"]], device=dev,
                     dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)                      # worst rank
    out = {"z_bit_identical": bool(t[0] > 0), "args_identical": bool(t[1] > 0), "y_bit_identical": bool(t[2] > 0),
           "y_max_rel_err": -float(t[3]), "dx_max_rel_err": -float(t[4]), "wgrad_max_rel_err": -float(t[5]),
           "tolerance": 1e-5, "ranks": world,
           "what": "sharded step vs the single-GPU step on the full graph (same weights, same dropout key), every "
                   "rank compares its destination rows; worst rank reported"}
    out["passed"] = bool(out["z_bit_identical"] and out["args_identical"] and out["y_max_rel_err"] <= 1e-5
                         and out["dx_max_rel_err"] <= 1e-5 and out["wgrad_max_rel_err"] <= 1e-5)
    del gfull, x_full, gy_full, y_f, grads, keep, keep2
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------
# shared pieces of the GPU arms
# ------------------------------------------------------------------------------------------
class Timer:
    def __init__(self, dev, dist, world):
        self.dev, self.dist, self.world = dev, dist, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return ms


def traffic_rows():
    if os.path.isfile(TRAFFIC_FILE):
        return json.load(open(TRAFFIC_FILE))
    return None


def emit(line, args):
    print(json.dumps(line), flush=True)
    if args.results:
        with open(args.results, "a") as f:
            f.write(json.dumps(line) + "\n")


# ------------------------------------------------------------------------------------------
# configs 1-3: L2-resident, launch-latency bound layers, replayed as one CUDA graph
# ------------------------------------------------------------------------------------------
def run_small(args, cfg, rank, world, local_rank):
    assert torch.cuda.is_available(), "bench.py needs a GPU (mma_b200 has no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import mma_b200
    from mma_b200 import _lib
    from mma_b200.parallel import allreduce_grads, split_graph_batch
    tm = Timer(dev, dist, world)
    d = small_inputs(args.config)
    E_total, F = d["E"], d["F"]
    torch.manual_seed(42)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    conv_cpu = None
    if d["kind"] == "nc":
        from mma_b200.node_classification.layers import MMA
        names = d["names"]
        add_all = [d["col"][d["rowptr"][i]:d["rowptr"][i + 1]].numpy() for i in range(d["n"])]
        ps = {nm: torch.nn.Parameter(d["masks"][nm].to(dev)) for nm in NC_ORDER}
        W, b = torch.nn.Parameter(d["W"].to(dev)), torch.nn.Parameter(d["b"].to(dev))
        layer = MMA(add_all, "new_sigmoid", 2, F, d["C"], W, b, *[ps[nm] for nm in NC_ORDER], d["p"], names, dev)
        with torch.no_grad():            # the constructor re-initialises (layers.py:143-198): put the bench weights back
            W.copy_(d["W"]); b.copy_(d["b"])
            for nm in NC_ORDER:
                ps[nm].copy_(d["masks"][nm])
        layer.device_seed = True         # dropout key on the device, advanced inside the graph
        adj = csr_to_sparse_adj(d["rowptr"], d["col"], d["n"]).to(dev)
        x = d["x"].to(dev).requires_grad_()
        gy = d["gy"].to(dev)
        plist = [W, b] + [ps[nm] for nm in names]
        call = lambda: layer(x, adj)
        E_loc, par = E_total, ("single GPU" if world == 1 else f"{world} independent replicas (the layer does not shard)")
        units = E_total * world
        cfg["hidden"], cfg["classes"], cfg["aggregators"] = F, d["C"], names
    else:
        conv_cpu = mma_b200.MMAConv(75, 75, d["names"], C2_SCAL, d["hist"], edge_dim=50, towers=5)
        import copy
        conv = copy.deepcopy(conv_cpu).to(dev)
        for a_cpu, a_dev in zip(conv_cpu.mask_parameters(), conv.mask_parameters()):
            with torch.no_grad():
                a_dev.copy_(a_cpu)       # the unregistered mask linears (Q1) are created on the device: same values
        conv.dropout = d["p"]
        conv.device_seed = True
        ei, xs, eas, gys = d["ei"], d["x"], d["ea"], d["gy"]
        if world > 1:                    # data parallel: whole graphs per rank, weight gradients all-reduced
            ei, _, nodes, eids, (xs, gys), (eas,) = split_graph_batch(d["ei"], d["batch"], rank, world, d["x"], d["gy"],
                                                                      edge_tensors=(d["ea"],))
        eig, x, ea, gy = ei.to(dev), xs.to(dev).requires_grad_(), eas.to(dev), gys.to(dev)
        plist = list(conv.parameters()) + conv.mask_parameters()
        graph = mma_b200.Graph.from_edge_index(eig, x.shape[0])
        call = lambda: conv(x, graph, ea)
        E_loc, par = int(ei.shape[1]), ("single GPU" if world == 1 else f"graph-batch data parallel x{world}")
        units = E_total
        cfg["hidden"], cfg["aggregators"], cfg["scalers"], cfg["towers"] = F, d["names"], C2_SCAL, 5
    cfg.update({"workload": d["workload"], "nodes": d["n"], "edges": E_total, "parallelism": par,
                "l2": "working set (<= 10 MB) is L2-resident by nature; a 256 MB buffer is rewritten between timed "
                      "iterations (L2 flush), each iteration timed by its own CUDA events",
                "launch": "one CUDA graph per step (fwd+bwd), dropout key advanced on the device"})

    def step():
        y = call()
        grads = torch.autograd.grad(y, [x] + plist, gy)
        if world > 1 and d["kind"] == "zinc":
            for p, g in zip(plist, grads[1:]):
                p.grad = g
            allreduce_grads(plist)
        return y, grads[0]

    warm = max(args.warmup, 3)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(warm):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    tm.barrier()
    _lib.reset_counters()
    step()
    launches_per_step = sum(_lib.LAUNCH_COUNTS.values())
    own_kernels = dict(_lib.LAUNCH_COUNTS)
    tm.barrier()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        y_s, gx_s = step()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for _ in range(warm):
        cg.replay()
    tm.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.time()
    for a, bb in evs:
        flush.fill_(1.0)
        a.record(); cg.replay(); bb.record()
    tm.barrier()
    t1 = time.time()
    ms = tm.max_over_ranks(sum(a.elapsed_time(bb) for a, bb in evs) / args.steps)
    # hot-L2 steady state (what a training loop of these layers sees), for the record
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cg.replay()
    e1.record(); tm.barrier()
    ms_hot = tm.max_over_ranks(e0.elapsed_time(e1) / args.steps)
    # per-kernel times: eager pass with CUDA events around each of our launches
    was = (layer if d["kind"] == "nc" else conv).device_seed
    (layer if d["kind"] == "nc" else conv).device_seed = False
    _lib.enable_timing(True)
    for _ in range(args.steps):
        flush.fill_(1.0)
        step()
    tm.barrier()
    ktimes = _lib.timing_summary()
    _lib.enable_timing(False)
    (layer if d["kind"] == "nc" else conv).device_seed = was
    # e2e: host buffers in, host buffers out, every step (the call a user of the layer makes)
    e2e = None
    if not args.no_e2e:
        xh = x.detach().cpu().pin_memory()
        yh = torch.empty(y_s.shape, dtype=torch.float32).pin_memory()
        for _ in range(2):
            with torch.no_grad():
                x.copy_(xh, non_blocking=True)
            cg.replay(); yh.copy_(y_s, non_blocking=True)
        tm.barrier()
        e0.record()
        for _ in range(args.steps):
            flush.fill_(1.0)
            with torch.no_grad():
                x.copy_(xh, non_blocking=True)
            cg.replay()
            yh.copy_(y_s, non_blocking=True)
        e1.record(); tm.barrier()
        flush_ms = 0.0
        e0b, e1b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0b.record()
        for _ in range(args.steps):
            flush.fill_(1.0)
        e1b.record(); tm.barrier()
        flush_ms = e0b.elapsed_time(e1b) / args.steps
        ms_e = tm.max_over_ranks(e0.elapsed_time(e1) / args.steps - flush_ms)
        e2e = {"value": units / (ms_e * 1e-3), "unit": "edges/s", "ms_per_step": ms_e,
               "h2d_bytes_per_step": x.numel() * 4 * world, "d2h_bytes_per_step": y_s.numel() * 4 * world,
               "note": "x from pinned host memory and the layer output y back to pinned host memory every step (the "
                       "L2-flush fill between steps is timed separately and subtracted)"}
    clocks = sampler.stop([("timed region", t0, t1), ("timed region + following passes", t0, time.time())]) if rank == 0 else None
    torch.cuda.synchronize()
    cg.reset()
    if rank != 0:
        if world > 1:
            dist.barrier(); sys.stdout.flush(); os._exit(0)
        return 0
    peak, peak_kind = peaks()
    per_kernel = {k: {"launches_per_step": c / args.steps, "ms_per_launch": m, "ms_per_step": m * c / args.steps}
                  for k, (c, m) in ktimes.items()}
    if d["kind"] == "nc":
        fb, bb_ = nc_algo_bytes(d["n"], E_loc, F, len(d["names"]))
        agg = [k for k in per_kernel if k.startswith("mma_nc_aggregate")]
    else:
        ab = algo_bytes(x.shape[0], E_loc, 5 * F, 2, 3, 2, False)
        fb, bb_ = ab["fwd"] + 4 * 5 * F * E_loc, ab["bwd"] + 8 * 5 * F * E_loc           # + the edge term R, dR
        agg = [k for k in per_kernel if k in ("mmconv_aggregate_fwd", "mmconv_aggregate_bwd_dst", "mma_segment_sum_rows")]
    agg_ms = sum(per_kernel[k]["ms_per_step"] for k in agg)
    roofline = {"bound": "hbm", "kernel": "aggregate op (" + " + ".join(sorted(agg)) + ")",
                "achieved": (fb + bb_) / 1e9 / (agg_ms * 1e-3) if agg_ms else None, "peak": peak, "unit": "GB/s",
                "frac": (fb + bb_) / 1e9 / (agg_ms * 1e-3) / peak if agg_ms else None, "traffic": None,
                "peak_kind": peak_kind, "algorithmic_GB": (fb + bb_) / 1e9, "launch_ms": agg_ms,
                "share_of_step": agg_ms / ms if ms else None,
                "note": "L2-resident and launch-latency bound (SURVEY 8(d)): the fraction is reported, not a target; "
                        "per-kernel times from the eager pass (cold L2)"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sec, thr = time_cpu(cpu_small_step_fn(d, conv_cpu), 3, 1)
        cpu = {"value": E_total / sec, "unit": "edges/s", "cores": thr, "kind": "port",
               "sample": f"full size (N={d['n']}, E={E_total}); oracle port of the reference layer, mean of 3 fwd+bwd "
                         f"= {sec * 1e3:.1f} ms"}
    line = {"metric": METRIC, "value": units / (ms * 1e-3), "unit": "edges/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if d["kind"] == "zinc" else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "detail": {"us_per_layer_call": ms * 1e3, "us_per_layer_call_hot_l2": ms_hot * 1e3,
                       "own_kernel_launches_per_step": own_kernels, "kernels": per_kernel}}
    emit(line, args)
    if world > 1:
        dist.barrier(); sys.stdout.flush(); os._exit(0)
    return 0


# ------------------------------------------------------------------------------------------
# configs 4 / 5: one large graph
# ------------------------------------------------------------------------------------------
def run_large(args, cfg, rank, world, local_rank):
    assert torch.cuda.is_available(), "bench.py needs a GPU (mma_b200 has no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import mma_b200
    from mma_b200 import _lib
    from mma_b200.parallel import ShardedGraph, allreduce_grads
    tm = Timer(dev, dist, world)
    barrier = tm.barrier
    N, E, F = CONFIGS[args.config]
    A, S = len(AGGR), len(SCAL)
    warm = max(args.warmup, 3)
    skewed = args.config.startswith("c5")

    torch.manual_seed(42)
    gen = torch.Generator(device=dev).manual_seed(42)
    if skewed:
        src, dst = powerlaw_edges(N, E, dev)
        E = int(dst.numel())
        cfg["edges"] = E
    else:
        src = torch.randint(0, N, (E,), generator=gen, device=dev)
        dst = torch.randint(0, N, (E,), generator=gen, device=dev)
    deg = torch.bincount(dst, minlength=N)
    hist = torch.bincount(deg).cpu()
    max_deg = int(deg.max().item())
    del deg
    conv = mma_b200.MMAConv(F, F, AGGR, SCAL, hist, towers=1, strict_reference=False).to(dev)
    conv.dropout = args.dropout
    conv.use_tensor_cores = not args.no_tc
    conv.comm_slices = args.slices
    conv.global_max_deg = max_deg
    verify = world > 1 and (args.verify if args.verify is not None else not skewed)
    if world > 1:
        graph = ShardedGraph(src, dst, N, rank, world, balance="edges" if skewed else "nodes")
        cfg["partition"] = {"balance": "edges" if skewed else "nodes", "rows": graph.rows, "edges": graph.E,
                            "max_rows": graph.max_rows}
        rows = graph.rows
        graph.local.build_transpose()
        from mma_b200 import fused_layer as _fl
        cfg["exchange"] = {"kind": ("copy engines over NVLink peer memory (mma_b200/peer.py)" if _fl.EXCHANGE == "peer"
                                    else "NCCL all-gather / reduce-scatter"),
                           "feature_windows": args.slices,
                           "bytes_on_the_wire_per_rank_per_step": 2 * (world - 1) * graph.max_rows * F * 4}
    else:
        graph = mma_b200.Graph(src, dst, N, sort_rows=True)      # degree-sorted CSR rows: scalers folded into the post GEMM
        rows = N
        _ = graph.max_deg
    if not verify:
        del src, dst
    torch.cuda.empty_cache()
    x = torch.randn(rows, F, device=dev, generator=gen).requires_grad_()
    gy = torch.randn(rows, F, device=dev, generator=gen)
    params = list(conv.parameters()) + conv.mask_parameters()

    def step(xin):
        y = conv(xin, graph)
        grads = torch.autograd.grad(y, [xin] + params, gy)
        if world > 1:
            for p, g in zip(params, grads[1:]):
                p.grad = g
            allreduce_grads(params)
        return y, grads[0]

    # ---------------- warm-up (eager, side stream), then capture ONE step in a CUDA graph ----------------
    # The step is ~60 kernel launches driven from Python (~10 ms of host time): eager replay is host-bound as
    # soon as the per-GPU work shrinks (N > 1).  The graph holds the whole fwd+bwd (+ the exchanges and the
    # weight-gradient all-reduce for N > 1); the dropout seed lives on the device and is advanced inside the
    # graph, so every replay draws a fresh mask exactly like an eager call.
    use_graph = not args.no_graph
    conv.device_seed = use_graph
    sampler = ClockSampler(local_rank)          # started early: nvidia-smi needs ~1 s before its first sample
    if rank == 0:
        sampler.start()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(warm):
            step(x)
    torch.cuda.current_stream(dev).wait_stream(side)
    barrier()
    parity = None
    if verify:
        parity = verify_sharded(conv, graph, x, gy, params, src, dst, N, step, dist)
        del src, dst
        torch.cuda.empty_cache()
        barrier()
    _lib.reset_counters()
    step(x)                                     # one counted eager step: kernels launched per step
    launches_per_step = sum(_lib.LAUNCH_COUNTS.values())
    barrier()
    loss_static = None
    cg = None
    if use_graph:
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            y_s, gx_s = step(x)
            loss_static = (y_s * gy).sum() + gx_s[0, 0] * 0

        def run():
            cg.replay()
    else:
        def run():
            step(x)
    for _ in range(warm):
        run()
    barrier()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_timed0 = time.time()
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    t_timed1 = time.time()
    ms = tm.max_over_ranks(e0.elapsed_time(e1) / args.steps)

    # ---------------- per-kernel launch durations: the same steps once more, eager, with CUDA events around every
    # launch of our kernels on the launching stream (events cannot bracket nodes inside a captured graph) --------
    conv.device_seed = False
    _lib.enable_timing(True)
    for _ in range(args.steps):
        step(x)
    barrier()
    ktimes = _lib.timing_summary()
    _lib.enable_timing(False)
    conv.device_seed = use_graph
    launches = launches_per_step * args.steps

    # ---------------- e2e: the layer called with HOST buffers, the boundary of the reference's own loop
    # (graph_regression/mma.py:153-160: `data.to(device)` in, `loss.item()` out): every step uploads its x from pinned
    # memory (copy stream, staging buffer, under the previous step) and reads the loss back.  A second measurement
    # also copies the layer output y [rows, hidden] to pinned host memory every step (`with_y_readback`). ----------
    e2e = None
    if not args.no_e2e:
        xh = torch.randn(rows, F).pin_memory()
        xs = torch.empty(rows, F, device=dev)
        yh = torch.empty(rows, F).pin_memory()
        ys = torch.empty(rows, F, device=dev)
        lossh = torch.empty((), dtype=torch.float32).pin_memory()
        up_s, down_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        main_s = torch.cuda.current_stream(dev)
        ready, free, y_ready, y_free = (torch.cuda.Event() for _ in range(4))

        def upload():
            with torch.cuda.stream(up_s):
                up_s.wait_event(free)
                xs.copy_(xh, non_blocking=True)
                ready.record(up_s)

        def e2e_step(read_y):
            main_s.wait_event(ready)
            with torch.no_grad():
                x.copy_(xs)                     # into the step's input buffer
            free.record(main_s)
            upload()                            # next step's input, under this step's compute
            if use_graph:
                cg.replay()
                y_out, loss = y_s, loss_static
            else:
                y_out, gx = step(x)
                loss = (y_out * gy).sum() + gx[0, 0] * 0
            lossh.copy_(loss, non_blocking=True)
            if read_y:
                main_s.wait_event(y_free)       # the previous step's download has left the staging buffer
                with torch.no_grad():
                    ys.copy_(y_out)
                y_ready.record(main_s)
                with torch.cuda.stream(down_s):
                    down_s.wait_event(y_ready)
                    yh.copy_(ys, non_blocking=True)
                    y_free.record(down_s)

        def e2e_loop(read_y, steps):
            free.record(main_s); y_free.record(main_s)
            upload()
            for _ in range(2):
                e2e_step(read_y)
            barrier()
            e0.record()
            for _ in range(steps):
                e2e_step(read_y)
            main_s.wait_stream(down_s)          # the last output has reached the host inside the timed region
            e1.record()
            barrier()
            main_s.wait_event(ready)            # drain the upload issued by the last step
            return tm.max_over_ranks(e0.elapsed_time(e1) / steps)

        ms_e = e2e_loop(False, args.steps)
        ms_y = e2e_loop(True, max(3, args.steps // 2))
        e2e = {"value": E / (ms_e * 1e-3), "unit": "edges/s", "ms_per_step": ms_e,
               "h2d_bytes_per_step": rows * F * 4 * world, "d2h_bytes_per_step": 4 * world,
               "note": "host boundary of the reference's own loop (graph_regression/mma.py:153-160: batch to the device, "
                       "loss.item() back): x uploaded from pinned host memory every step (copy stream, double-buffered "
                       "under the previous step), the loss read back every step; y, dx and the weight gradients stay on "
                       "the device, where the next layer / the previous layer / the optimizer consume them",
               "with_y_readback": {"ms_per_step": ms_y, "value": E / (ms_y * 1e-3),
                                   "d2h_bytes_per_step": (rows * F * 4 + 4) * world,
                                   "note": "same, plus the layer output y [rows, hidden] copied to pinned host memory "
                                           "every step on a second copy stream (PCIe-bound for N > 1)"}}

    clocks = None
    if rank == 0:
        clocks = sampler.stop([("timed region", t_timed0, t_timed1),
                               ("timed region + per-kernel timing pass + e2e steps", t_timed0, time.time())])

    def finish():
        """Tears the run down without dist.destroy_process_group(): destroying a communicator that a live CUDA
        graph has captured collectives on was seen to hang; the graph is reset first and the process then
        leaves through os._exit once everything is flushed."""
        nonlocal cg
        torch.cuda.synchronize()
        if use_graph:
            cg.reset()
            cg = None
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return 0

    # ---------------- roofline: algorithmic bytes / CUDA-event time / measured copy peak ----------------
    peak, peak_kind = peaks()
    n_loc, e_loc = (graph.rows, graph.E) if world > 1 else (N, E)
    from mma_b200 import fused_layer
    A_mat = fused_layer.materialised_blocks(AGGR) if hasattr(fused_layer, "materialised_blocks") else A
    S_mat = 1                               # scaler blocks materialised by K1 (folded into the post GEMM)
    ab = algo_bytes(n_loc, e_loc, F, A_mat, S_mat, 2, True)
    per_kernel = {}
    for name, (cnt, mean_ms) in ktimes.items():
        per_step = cnt / args.steps
        nbytes = ab.get(name)
        if name == "mma_segment_sum_rows" and world > 1:
            nbytes = 4 * (N + 1) + 4 * F * e_loc + 4 * F * N          # partial dQ over ALL sources
        per_kernel[name] = {"launches_per_step": per_step, "ms_per_launch": mean_ms,
                            "ms_per_step": mean_ms * per_step,
                            "own_contract_GB_per_step": None if nbytes is None else nbytes / 1e9,
                            "dram_utilisation_GBps": None if nbytes is None else nbytes / 1e9 / (mean_ms * per_step * 1e-3)}
    kms = lambda k: per_kernel[k]["ms_per_step"] if k in per_kernel else 0.0
    t_fwd = kms("mmconv_aggregate_fwd")
    t_bwd = kms("mmconv_aggregate_bwd_dst") + kms("mma_segment_sum_rows") + kms("mmconv_aggregate_bwd_src")
    agg_ms = t_fwd + t_bwd
    traffic = traffic_rows() if (args.config == "c4" and world == 1) else None
    parts = {"forward (mmconv_aggregate_fwd)": (ab["fwd"], t_fwd, ["mmconv_aggregate_fwd"]),
             "backward (mmconv_aggregate_bwd_dst + transpose pass)": (ab["bwd"], t_bwd,
                                                                      ["mmconv_aggregate_bwd_dst", "mma_segment_sum_rows",
                                                                       "mmconv_aggregate_bwd_src"])}
    dom = max(parts, key=lambda k: parts[k][1])
    roofline = None
    if parts[dom][1] > 0:
        nbytes, t_ms, names = parts[dom]
        tr = None
        if traffic:
            tr = sum(traffic["bytes_per_launch"].get(k, 0.0) for k in names) or None
        own = sum((ab.get(k) or 0) for k in names if k in per_kernel)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": nbytes / 1e9 / (t_ms * 1e-3), "peak": peak,
                    "unit": "GB/s", "frac": nbytes / 1e9 / (t_ms * 1e-3) / peak, "traffic": tr,
                    "traffic_source": None if not traffic else traffic.get("source"),
                    "peak_kind": peak_kind, "algorithmic_GB_per_launch": nbytes / 1e9,
                    "bytes_formula": "SURVEY.md 8(d) with the materialised shapes "
                                     f"(A={A_mat} aggregate blocks, S={S_mat}: scalers folded into the post GEMM)",
                    "launch_ms": t_ms, "share_of_step": t_ms / ms,
                    "dram_utilisation_frac": own / 1e9 / (t_ms * 1e-3) / peak if own else None,
                    "dram_utilisation_note": "own-contract bytes of these kernels (incl. the per-edge gradient rows G "
                                             "written and re-read) / time / peak: how busy DRAM is, NOT algorithmic "
                                             "efficiency"}
    own_ms = sum(v["ms_per_step"] for v in per_kernel.values())
    step_algo = (ab["fwd"] + ab["bwd"]) / 1e9
    extra = {"kernels": per_kernel,
             "aggregate_op": {"ms_per_step": agg_ms, "edges_per_s": E / (agg_ms * 1e-3) if agg_ms else None,
                              "algorithmic_GB_per_step_per_gpu": step_algo,
                              "frac_of_hbm_peak": step_algo / (agg_ms * 1e-3) / peak if agg_ms else None,
                              "forward_frac": ab["fwd"] / 1e9 / (t_fwd * 1e-3) / peak if t_fwd else None,
                              "backward_frac": ab["bwd"] / 1e9 / (t_bwd * 1e-3) / peak if t_bwd else None,
                              "note": "K1 fwd + bwd-dst + transpose pass; bytes from the SURVEY 8(d) formulas with the "
                                      "materialised shapes; dense GEMMs excluded"},
             "whole_layer": {"frac_of_hbm_peak_8d_bytes": step_algo / (ms * 1e-3) / peak,
                             "target_ms": 25.3 if args.config == "c4" and world == 1 else None},
             "dense_and_other_ms_per_step": ms - agg_ms, "own_kernels_ms_per_step": own_ms,
             "torch_glue_ms_per_step": ms - own_ms}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sub = reference_sample(E, 2, 0, budget_s=25.0)
        fn, e_run = cpu_c4_step_fn(N // sub, E // sub, F, skewed=skewed)
        sec, thr = time_cpu(fn, 1, 0)
        cpu = {"value": e_run / sec, "unit": "edges/s", "cores": thr, "kind": "port",
               "sample": f"1/{sub} sub-graph (N={N // sub}, E={e_run}, hidden={F}), oracle port of the reference "
                         f"layer, one fwd+bwd = {sec:.1f} s"}

    line = {"metric": METRIC, "value": E / (ms * 1e-3), "unit": "edges/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "parity_check": parity, "detail": extra}
    emit(line, args)
    finish()
    return 0


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mma_b200", choices=["mma_b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS) + sorted(SMALL))
    ap.add_argument("--dropout", type=float, default=0.5, help="0.5 = the reference's always-on dropout")
    ap.add_argument("--slices", type=int, default=1, help="feature windows of the sharded exchange pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="cuBLAS fp32 GEMMs instead of the tcgen05 3xTF32 layer")
    ap.add_argument("--no-graph", action="store_true", help="eager launches in the timed region (host-bound for N > 1)")
    ap.add_argument("--verify", dest="verify", action="store_true", default=None,
                    help="N > 1: compare the sharded step with the single-GPU step on the full graph (default: on for "
                         "config 4, off for config 5)")
    ap.add_argument("--no-verify", dest="verify", action="store_false")
    ap.add_argument("--results", default=None, help="append the JSON line to this file as well")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config in SMALL:
        cfg = {"workload": small_inputs(args.config)["workload"] if args.impl == "reference" else None}
    else:
        N, E, F = CONFIGS[args.config]
        skewed = args.config.startswith("c5")
        cfg = {"workload": (f"config 5: power-law graph (alpha 2.1, max in-degree {E // 200}) N={N} E~{E} hidden={F}, "
                            if skewed else f"config 4: uniform random graph N={N} E={E} hidden={F}, ") +
                           f"MMAConv fwd+bwd, aggregators "
                           f"{','.join(AGGR)} x scalers {','.join(SCAL)}, towers=1, dropout {args.dropout}",
               "nodes": N, "edges": E, "hidden": F, "aggregators": AGGR, "scalers": SCAL,
               "parallelism": f"dst-range x{world}" if world > 1 else "single GPU",
               "l2": "inputs larger than L2 (x, P, Q, Z, G are 1-16 GB each vs 126 MB L2): no flush needed",
               "launch": "eager" if args.no_graph else "one CUDA graph per step (fwd+bwd+exchanges), seed advanced on device",
               "kernel_timing": "CUDA events around every launch in an eager pass of the same steps after the timed region"}
    if args.impl == "reference":
        if rank != 0:
            return 0
        run_reference(args, cfg)
        return 0
    if args.config in SMALL:
        return run_small(args, cfg, rank, world, local_rank)
    return run_large(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
