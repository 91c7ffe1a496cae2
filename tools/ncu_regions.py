"""Per-region lane utilisation of one kernel capture (development tool):
python tools/ncu_regions.py gpurun_out/prof.ncu-rep [rays] [chunk]"""
import csv, subprocess, sys
rep = sys.argv[1]; rays = float(sys.argv[2]) if len(sys.argv) > 2 else 8388608.0; chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:120])
hdr = rows[1]; data = rows[2:]
ia = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed'); isamp = hdr.index('# Samples'); isrc = hdr.index('Source')
tot_i = sum(int(r[ia]) for r in data); tot_t = sum(int(r[it]) for r in data); tot_s = sum(int(r[isamp]) for r in data)
print('warp-instr/ray %.1f  thread-instr/ray %.0f  avg lanes %.2f' % (tot_i / rays, tot_t / rays, tot_t / tot_i))
for b in range(0, len(data), chunk):
    seg = data[b:b + chunk]
    wi = sum(int(r[ia]) for r in seg); ti = sum(int(r[it]) for r in seg); ss = sum(int(r[isamp]) for r in seg)
    if wi == 0: continue
    ops = {}
    for r in seg:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = ' '.join('%s:%d' % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print('%4d-%4d warp-instr/ray %6.1f (%4.1f%%) lanes %5.1f stall-samples %4.1f%%  %s' % (b, b + chunk, wi / rays, 100 * wi / tot_i, ti / max(wi, 1), 100 * ss / tot_s, top))
