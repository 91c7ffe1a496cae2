"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_*] --csv): python tools/launch_summary.py file.csv [skip]"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = defaultdict(lambda: defaultdict(float)); cnt = defaultdict(int); seen = set()
for r in rows[hi + 1:]:
    if len(r) <= vi or int(r[ii]) < skip: continue
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"^void (spb::)?", "", name)[:70]
    v = float(r[vi].replace(",", "") or 0)
    per[name][r[mi]] += v
    if (r[ii], name) not in seen: seen.add((r[ii], name)); cnt[name] += 1
tot = sum(p["gpu__time_duration.sum"] for p in per.values())
print("%-72s %6s %10s %6s %10s %10s" % ("kernel", "n", "time", "%", "dram rd MB", "dram wr MB"))
for name, p in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = p["gpu__time_duration.sum"]
    print("%-72s %6d %10.1f %6.1f %10.1f %10.1f" % (name, cnt[name], t, 100 * t / tot, p.get("dram__bytes_read.sum", 0), p.get("dram__bytes_write.sum", 0)))
print("total time %.1f (unit as in the csv)" % tot)
