"""Minimal stdlib .xlsx reader (zipfile + ElementTree).  Oracle/test infrastructure.

Stands in for ``ExcelReaders.readxlsheet`` used at readin_functions.jl:204-205.
Returns a dense 2-D list-of-lists over the sheet's used range; cell values are
``float`` (numbers, incl. Excel serial dates), ``str`` or ``None`` (blank / error).
"""
import re
import zipfile
import xml.etree.ElementTree as ET

_NS = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main",
       "r": "http://schemas.openxmlformats.org/officeDocument/2006/relationships",
       "pr": "http://schemas.openxmlformats.org/package/2006/relationships"}
_REF = re.compile(r"([A-Z]+)([0-9]+)")


def _col_index(letters):
    n = 0
    for ch in letters:
        n = n * 26 + (ord(ch) - 64)
    return n - 1


def read_sheet(path, sheet_name):
    with zipfile.ZipFile(path) as z:
        wb = ET.fromstring(z.read("xl/workbook.xml"))
        rels = ET.fromstring(z.read("xl/_rels/workbook.xml.rels"))
        rid2target = {r.get("Id"): r.get("Target") for r in rels.findall("pr:Relationship", _NS)}
        target = None
        for s in wb.find("m:sheets", _NS).findall("m:sheet", _NS):
            if s.get("name") == sheet_name:
                target = rid2target[s.get("{%s}id" % _NS["r"])]
        if target is None:
            raise KeyError(sheet_name)
        shared = []
        if "xl/sharedStrings.xml" in z.namelist():
            sst = ET.fromstring(z.read("xl/sharedStrings.xml"))
            for si in sst.findall("m:si", _NS):
                shared.append("".join(t.text or "" for t in si.iter("{%s}t" % _NS["m"])))
        root = ET.fromstring(z.read("xl/" + target.lstrip("/").replace("xl/", "")))
    cells = {}
    max_r = max_c = 0
    for c in root.iter("{%s}c" % _NS["m"]):
        m = _REF.match(c.get("r"))
        ci, ri = _col_index(m.group(1)), int(m.group(2)) - 1
        t = c.get("t")
        v = c.find("m:v", _NS)
        val = None
        if t == "s" and v is not None:
            val = shared[int(v.text)]
        elif t == "inlineStr":
            val = "".join(x.text or "" for x in c.iter("{%s}t" % _NS["m"]))
        elif t == "str" and v is not None:
            val = v.text
        elif t == "e":
            val = None
        elif t == "b" and v is not None:
            val = float(v.text)
        elif v is not None and v.text is not None:
            val = float(v.text)
        if val is not None:
            cells[(ri, ci)] = val
            max_r, max_c = max(max_r, ri), max(max_c, ci)
    grid = [[None] * (max_c + 1) for _ in range(max_r + 1)]
    for (ri, ci), val in cells.items():
        grid[ri][ci] = val
    return grid


def excel_serial_to_ymd(serial):
    """Excel 1900-system serial date -> (year, month, day)."""
    import datetime
    d = datetime.date(1899, 12, 30) + datetime.timedelta(days=int(round(serial)))
    return d.year, d.month, d.day
