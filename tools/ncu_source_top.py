"""Top stalled SASS instructions of an `ncu --page source --csv` dump (gzip ok): python tools/ncu_source_top.py file [N] [kernel-index]"""
import csv, gzip, sys
path = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25; want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
kernels = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kernels.append(cur); continue
    if cur is None or not r: continue
    if r[0] == "Address": cur["hdr"] = r; continue
    cur["rows"].append(r)
k = kernels[want]
h = k["hdr"]; ci = {n: i for i, n in enumerate(h)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in k["rows"])
inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in k["rows"])
print(k["name"][:120]); print("samples", tot, "warp-instructions", inst, "SASS lines", len(k["rows"]))
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
order = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][ci["# Samples"]] or 0))[:N]
for i in sorted(order):
    r = k["rows"][i]
    s = int(r[ci["# Samples"]] or 0)
    top = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print("%5d %5.1f%%  line %4d  exec %9s thr %4s  %-60s %s" % (s, 100.0 * s / max(tot, 1), i, r[ci["Instructions Executed"]], r[ci["Avg. Threads Executed"]][:4], r[ci["Source"]].strip()[:60], " ".join("%s:%d" % (c, v) for v, c in top if v)))
