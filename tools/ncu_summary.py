"""Summarise an .ncu-rep (one or more kernel captures) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_trace_closest.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "local_load", "local_store",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# summary of %s (ncu --set full --clock-control none); one block per captured launch\n" % rep)
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write("\n== %s  grid=%s block=%s\n" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
            for i, h in enumerate(hdr):
                if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")) or "local_op" in h:
                    f.write("%-90s %-14s %s\n" % (h, units[i], vals[i]))


if __name__ == "__main__":
    main()
