#!/usr/bin/env python
"""Coarse stall-sample map of a kernel from `ncu --page source --csv`: samples per chunk of SASS with hints of what the
chunk does (DMMA / TMA / barriers / shuffles), plus the stall-reason totals.  usage: ncu_regions.py report.ncu-rep [kernel-regex] [chunk]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 120
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
# several kernels/functions may follow each other: split on "Kernel Name" rows
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    hdr = b["rows"][0]; data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    S = sum(int(r[ix["# Samples"]]) for r in data)
    print("==", b["name"][:90], "instructions", len(data), "samples", S)
    tot = collections.Counter()
    for r in data:
        for h in stalls:
            tot[h] += int(r[ix[h]])
    print("  " + "  ".join(f"{h[6:]} {100 * v / max(S, 1):.1f}%" for h, v in tot.most_common(9)))
    for c0 in range(0, len(data), chunk):
        seg = data[c0:c0 + chunk]
        n = sum(int(r[ix["# Samples"]]) for r in seg)
        if n < 0.004 * S:
            continue
        src = " ".join(r[ix["Source"]] for r in seg)
        hints = [k for k in ("DMMA", "UTMALDG", "SYNCS.PHASECHK", "BAR.SYNC", "SHFL", "MUFU", "LDG", "STG", "LDL", "STL", "CALL", "RET") if k in src]
        st = collections.Counter()
        for r in seg:
            for h in stalls:
                st[h] += int(r[ix[h]])
        ex = max(int(r[ix["Instructions Executed"]]) for r in seg)
        print(f"  [{c0:6d}] {100 * n / S:5.1f}%  maxexec {ex:9d}  " + ",".join(hints) + "  | " + " ".join(f"{h[6:]}:{100 * v / max(n, 1):.0f}" for h, v in st.most_common(4)))
