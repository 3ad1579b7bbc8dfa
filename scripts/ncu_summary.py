"""Summarise .ncu-rep captures (run here, no GPU needed) into a small text file for profiles/.
usage: python scripts/ncu_summary.py out.txt rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed",
    "smsp__mem_tensor_reads_op_utcmma_matrix_c.sum.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    out = open(sys.argv[1], "w")
    for rep in sys.argv[2:]:
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(r.stdout.splitlines()))
        if len(rows) < 3:
            out.write(f"== {rep}: no data\n")
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            out.write(f"== {rep}\n   kernel: {d.get('Kernel Name', '?')}\n")
            for k in KEYS:
                if k in d:
                    out.write(f"   {k} = {d[k]} {u.get(k, '')}\n")
    out.close()


if __name__ == "__main__":
    main()
