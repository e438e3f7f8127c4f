"""ncu_summary.py RAW.csv > SUMMARY.csv — boil an `ncu -i rep --page raw --csv` dump down to the columns the roofline
discussion uses (one row per profiled launch)."""
import csv
import sys

COLS = [
    ("Kernel Name", "Kernel Name"), ("Grid Size", "Grid Size"), ("Block Size", "Block Size"),
    ("gpu__time_duration.sum", "gpu__time_duration.sum"),
    ("dram__bytes_read.sum", "dram__bytes_read.sum"), ("dram__bytes_write.sum", "dram__bytes_write.sum"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "registers_per_thread"),
    ("launch__occupancy_limit_registers", "occ_limit_registers"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_shared_mem"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}
    w = csv.writer(sys.stdout)
    present = [(n, label) for n, label in COLS if n in idx]
    w.writerow([label + (f" [{units[idx[n]]}]" if units[idx[n]] else "") for n, label in present])
    for r in rows[hdr + 2:]:
        if len(r) == len(names):
            w.writerow([r[idx[n]] for n, _ in present])


if __name__ == "__main__":
    main()
