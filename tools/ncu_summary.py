"""Rows of profiles/rNN_ncu_full_summary.csv from an `ncu --set full` report.
usage: python tools/ncu_summary.py <report.ncu-rep> <capture-name>[:<instance>] [...more pairs] >> csv
Each pair appends `capture,metric,value,unit` rows for one kernel instance (default 0) of the
report; bench.py's ncu_traffic() reads the dram__bytes_* rows by capture name."""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
]


def rows_of(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    w = csv.writer(sys.stdout)
    args = sys.argv[1:]
    for report, name in zip(args[::2], args[1::2]):
        name, _, inst = name.partition(":")
        head, units, body = rows_of(report)
        ix = {n: i for i, n in enumerate(head)}
        r = body[int(inst or 0)]
        w.writerow([name, "Kernel Name", r[ix["Kernel Name"]], ""])
        for m in METRICS:
            if m in ix:
                w.writerow([name, m, r[ix[m]], units[ix[m]]])


if __name__ == "__main__":
    main()
