"""Print selected metrics of an `ncu --page raw --csv` file: python tools/ncu_csv_show.py file.raw.csv [regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|launch__registers|launch__grid_size|launch__occupancy_limit|sm__warps_active.avg.pct|smsp__thread_inst_executed_per_inst|smsp__issue_active.avg.pct|sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu|cbu|adu)\.avg\.pct_of_peak_sustained_active|dram__bytes_(read|write)\.sum$|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|lts__throughput.avg.pct|dram__throughput.avg.pct|issue_stalled.*per_issue_active|smsp__inst_executed.sum$|local_op_(ld|st)\.sum$|l1tex__data_pipe_lsu_wavefronts.avg.pct")
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:110], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            print("  %-86s %-10s %s" % (h, u, v))
