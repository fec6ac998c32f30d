"""Generate the committed fixtures under tests/golden/.  Run ONLY in the build container
(needs /root/reference).  Usage:  python tests/golden/make_golden.py

Writes
  hom_fac_1_panels.npz   output of the ingestion oracle (oracle/readin.py) on
                         /root/reference/data/hom_fac_1.xlsx for datatype :All and :Real
                         (the notebook's `dataset_all` / `dataset`, Stock_Watson.ipynb:160,180)
  notebook_tables.json   the numeric tables stored as cell outputs of Stock_Watson.ipynb
                         (Tables 2A, 2B, 2C, 3 (visible part), 4, 5) -- the reference's only
                         golden values
"""
import json, os, re, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.readin import readin_data  # noqa: E402

REF = "/root/reference"
ANSI = re.compile(r"\x1b\[[0-9;]*m")


def parse_millboard(text):
    rows = []
    for line in ANSI.sub("", text).splitlines():
        cells = [c.strip() for c in line.strip().strip("|").split("|")]
        try:
            vals = [float(c) for c in cells[1:]]
        except ValueError:
            continue
        if vals:
            rows.append(vals)
    return rows


def main():
    xlsx = os.path.join(REF, "data", "hom_fac_1.xlsx")
    a = readin_data(xlsx, "All"); r = readin_data(xlsx, "Real")
    np.savez_compressed(os.path.join(HERE, "hom_fac_1_panels.npz"),
                        all_bpdata=a["bpdata"], all_inclcode=a["inclcode"], all_names=np.array(a["bpnamevec"]),
                        real_bpdata=r["bpdata"], real_inclcode=r["inclcode"], real_names=np.array(r["bpnamevec"]),
                        calds=np.array(a["calds"]))
    nb = json.load(open(os.path.join(REF, "Stock_Watson.ipynb")))
    outs = {}
    for i, c in enumerate(nb["cells"]):
        if c["cell_type"] != "code":
            continue
        txt = []
        for o in c.get("outputs", []):
            if "text" in o:
                txt.append("".join(o["text"]))
            elif "data" in o and "text/plain" in o["data"]:
                txt.append("".join(o["data"]["text/plain"]))
        outs[i] = txt
    tables = {}
    tables["table2A"] = parse_millboard(outs[35][0])      # cols: nfac, traceR2, margR2, BN-ICp2, AH-ER
    tables["table2B"] = parse_millboard(outs[37][0])
    tables["table2C"] = parse_millboard(outs[39][0])      # rows: n dynamic; cols: n dyn, then static 1..10
    # Table 3: 207x10 R2, visible: first 13 + last 12 rows, columns 1-3 and 8-10
    t3 = []
    for line in outs[55][0].splitlines()[1:]:
        toks = line.replace("…", " ").replace("⋱", " ").replace("⋮", " ").split()
        if len(toks) == 6:
            t3.append([float(x) for x in toks])
    tables["table3_visible"] = {"rows_head": 13, "rows_tail": 12, "cols": [1, 2, 3, 8, 9, 10], "values": t3}
    nums = lambda s: [[float(x) for x in ln.split()] for ln in s.splitlines()[1:] if ln.strip()]
    tables["table4"] = {"chow_qlr_r4": nums(outs[58][0]), "chow_qlr_r8": nums(outs[58][1]),
                        "cor_r4": nums(outs[58][2]), "cor_r8": nums(outs[58][3])}
    t5 = {}
    lines = outs[61][0].splitlines()
    for k in range(0, len(lines), 3):
        name = lines[k].split()[1]
        t5[name] = {"resid": [float(x) for x in lines[k + 1].strip("[]").split()],
                    "level": [float(x) for x in lines[k + 2].strip("[]").split()]}
    tables["table5"] = t5
    tables["source"] = "stored cell outputs of /root/reference/Stock_Watson.ipynb (Julia 1.0.2)"
    json.dump(tables, open(os.path.join(HERE, "notebook_tables.json"), "w"), indent=1)
    print({k: (len(v) if isinstance(v, list) else "...") for k, v in tables.items()})


if __name__ == "__main__":
    main()
