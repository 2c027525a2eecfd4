"""per-kernel summary of an `ncu --page raw --csv` export (one `ncu --set full` capture)
usage: python tools/ncu_summary.py raw.csv > profiles/rNN_ncu_full_summary.txt"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sectors_op_red.sum",
    "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_op_write.sum",
    "lts__t_sectors_op_read.sum",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__grid_size",
    "launch__block_size",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
    "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
    "smsp__average_warp_latency_issue_stalled_membar.ratio",
    "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    head, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(head):
        for w in WANT:
            if h == w or h.endswith("." + w):
                col.setdefault(w, i)
    ki = head.index("Kernel Name")
    seen = {}
    for r in rows[2:]:
        if len(r) != len(head):
            continue
        name = r[ki]
        seen.setdefault(name, []).append(r)
    for name, rs in seen.items():
        r = rs[-1]  # the last (warm) launch of the kernel
        print(f"--- {name[:110]}   [{len(rs)} launch(es) captured, last shown]")
        for w in WANT:
            if w in col:
                print(f"    {w:74s} {r[col[w]]} {units[col[w]]}")


if __name__ == "__main__":
    main()
