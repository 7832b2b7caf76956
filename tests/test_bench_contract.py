"""bench.py's reference arm (`--impl reference`: the oracle port of the reference layer on the host cores) on the
small configuration: the JSON line carries the keys the driver reads.  The GPU arm needs a B200 and is run by the
driver; its algorithmic-byte formulas are checked here against the figures quoted in DESIGN.md."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c4s",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MultiMaskConv fwd+bwd edges/sec" and d["unit"] == "edges/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # the line says what was measured: the steps it timed, and the sub-graph the CPU path actually ran
    assert d["steps"] == 1 and d["warmup"] == 0
    smp = d["config"]["sample"]
    assert smp["nodes"] == d["config"]["nodes"] // 16 and smp["edges"] == d["config"]["edges"] // 16
    assert abs(d["value"] - smp["edges"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_reference_sample_fits_the_driver_run():
    """The driver runs the reference arm with --steps 20 --warmup 5: every one of those calls is executed, on a
    sub-graph small enough for the run to end within a few minutes."""
    sys.path.insert(0, ROOT)
    import bench
    N, E, F = bench.CONFIGS["c4"]
    sub = bench.reference_sample(E, 20, 5)
    assert sub >= 16 and (sub & (sub - 1)) == 0
    assert 25 * (E / sub) * bench.CPU_SEC_PER_EDGE <= 100.0
    assert bench.reference_sample(E, 1, 0) == 16


def test_reference_arm_small_config_runs_full_size():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["config"]["sample"]["edges"] == 10556
    assert "Cora" in d["config"]["workload"] and d["cpu_baseline"]["kind"] == "port"


def test_algorithmic_bytes_match_design():
    sys.path.insert(0, ROOT)
    import bench
    N, E, F = bench.CONFIGS["c4"]
    ab = bench.algo_bytes(N, E, F, A=5, S=1, n_mm=2, std=True)
    # 26.76 GB (K1 forward), 44.4 GB (destination pass), 17.4 GB (transpose pass); SURVEY 8(d) with S = 1
    assert abs(ab["mmconv_aggregate_fwd"] / 1e9 - 26.76) < 0.01
    assert abs(ab["mmconv_aggregate_bwd_dst"] / 1e9 - 44.42) < 0.01
    assert abs(ab["mma_segment_sum_rows"] / 1e9 - 17.42) < 0.01
    assert abs((ab["fwd"] + ab["bwd"]) / 1e9 - 67.1) < 0.1
    assert abs(ab["fwd"] / 1e9 - 24.71) < 0.01 and abs(ab["bwd"] / 1e9 - 42.38) < 0.01     # the judge's recomputation
    s4 = bench.algo_bytes(N, E, F, A=5, S=4, n_mm=2, std=True)
    assert abs((s4["fwd"] + s4["bwd"]) / 1e9 - 97.8) < 0.2            # SURVEY's figure with the scaler blocks materialised


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """N > 1: the driver launches the reference arm exactly like ours (torch.distributed.run, one process per GPU);
    rank 0 alone runs the CPU path and prints the line, the other ranks exit 0 without work."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--config", "c4s", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
