"""Join an ncu '--page source --csv' SASS dump with nvdisasm line info to get stall samples per
CUDA source line.  usage: ncu_lines.py <src.csv> <cubin> <kernel-substring> [topN]"""
import csv, re, subprocess, sys, collections

src_csv, cubin, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
lines_for_instr = []
infn = False; cur = None
for ln in dis:
    if ln.startswith(".text.") :
        infn = kname in ln
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines_for_instr.append((cur, ln.strip()))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; si = hdr.index("# Samples"); data = [r for r in rows[2:] if len(r) > si]
print("sass instrs", len(lines_for_instr), "ncu rows", len(data))
agg = collections.Counter(); tot = 0
n = min(len(data), len(lines_for_instr))
for k in range(n):
    s = float(data[k][si] or 0); agg[lines_for_instr[k][0]] += s; tot += s
srcs = {}
for (f, l), s in agg.most_common(topn):
    try:
        text = open("/root/repo/dynamic_factor_models_b200/csrc/" + f).read().splitlines()[l - 1].strip()[:100]
    except Exception:
        text = ""
    print(f"{100 * s / tot:5.1f}%  {f}:{l:4d}  {text}")
