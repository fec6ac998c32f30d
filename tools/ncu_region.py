"""Stall-reason totals for a range of source lines (ncu source CSV + nvdisasm line info).
usage: ncu_region.py <src.csv> <cubin> <kernel-substr> <file> <line_lo> <line_hi> [...more lo hi]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname, fname = sys.argv[1:5]
ranges = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(5, len(sys.argv), 2)]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []; infn = False; cur = None
for ln in dis:
    if ln.startswith(".text."): infn = kname in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append((cur, ln.strip()))
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; data = rows[2:]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in reasons}; si = hdr.index("# Samples")
tot_all = sum(float(r[si] or 0) for r in data if len(r) > si)
agg = collections.Counter(); n = 0; opc = collections.Counter()
for k in range(min(len(data), len(lines))):
    (f, l) = lines[k][0] if lines[k][0] else ("", 0)
    if f == fname and any(lo <= l <= hi for lo, hi in ranges):
        n += float(data[k][si] or 0)
        for h in reasons: agg[h] += float(data[k][idx[h]] or 0)
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", lines[k][1])
        if m: opc[m.group(1).split(".")[0]] += float(data[k][si] or 0)
print(f"region samples {n:.0f} = {100*n/tot_all:.1f}% of all")
for h, v in agg.most_common(8): print(f"  {h:22s} {v:8.0f}  {100*v/max(n,1):5.1f}%")
print("  by opcode:", [(k, int(v)) for k, v in opc.most_common(10)])
