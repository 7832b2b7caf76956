"""Selected columns of an `ncu --set full` report as a small CSV for profiles/ (the .ncu-rep files themselves are too
large to commit):  python scripts/ncu_select.py gpurun_out/X.ncu-rep profiles/X_selected.csv
(the first argument may also be the `--page raw --csv` export of a report, made on the GPU box)"""
import csv
import subprocess
import sys

COLS = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        # shared-memory pipe and L2 -> SM traffic (what bounds the tcgen05 GEMMs)
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "derived__lts__lts2xbar_bytes.sum.per_second"]


def main(rep, out):
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] if i < len(r) else "" for i in idx])
    print(f"{out}: {len(rows) - 2} kernels, {len(idx)} columns")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
