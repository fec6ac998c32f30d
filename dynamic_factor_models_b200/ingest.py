"""Panel ingestion: the step BEFORE the hot path.

Host-side mirror of the reference's `readin_data` (readin_functions.jl:355-382 and callees) for the Stock-Watson
workbook `data/hom_fac_1.xlsx`, so that the C1 / C4 configurations can be run end to end without Julia:
xlsx sheet -> deflate -> monthly-to-quarterly -> transform -> outlier adjustment -> merge -> biweight detrending.
One-off CPU work on a 224 x 207 panel (milliseconds), hence plain numpy; the estimation itself is CUDA only
(`api.py`).  Missing values are NaN.  Every function cites the reference lines it mirrors; the result is pinned
against the committed fixture `tests/golden/hom_fac_1_panels.npz` (`tests/test_ingest.py`).
"""
import datetime
import re
import zipfile
import xml.etree.ElementTree as ET
from dataclasses import dataclass

import numpy as np

_MAIN = "{http://schemas.openxmlformats.org/spreadsheetml/2006/main}"
_RELS = "{http://schemas.openxmlformats.org/package/2006/relationships}"
_RID = "{http://schemas.openxmlformats.org/officeDocument/2006/relationships}id"


@dataclass
class Panel:
    """What `readin_data` returns (readin_functions.jl:376-381), restricted to what the estimation uses."""
    bpdata: np.ndarray            # T x ns, transformed, outlier-adjusted, detrended; NaN = missing
    bpdata_unfiltered: np.ndarray  # same before the biweight detrending
    inclcode: np.ndarray          # ns, 1 = used for factor estimation
    bpcatcode: np.ndarray         # ns, category codes (columns are sorted by them)
    bpnamevec: list               # ns series names (upper case)
    calds: list                   # T (year, quarter)
    calvec: np.ndarray            # T, year + (quarter - 1) / 4

    def row(self, year, quarter):
        """`find_row_number` (dfm_functions.ipynb:555-556): 1-based row of (year, quarter)."""
        return self.calds.index((year, quarter)) + 1


# --------------------------------------------------------------------------------------- xlsx
def read_xlsx_sheet(path, sheet):
    """Dense object array of one worksheet (stands in for ExcelReaders.readxlsheet, readin_functions.jl:204-205):
    numbers (incl. Excel serial dates) as float, text as str, blanks / error cells as None."""
    with zipfile.ZipFile(path) as z:
        book = ET.fromstring(z.read("xl/workbook.xml"))
        rel = {r.get("Id"): r.get("Target") for r in ET.fromstring(z.read("xl/_rels/workbook.xml.rels")).iter(_RELS + "Relationship")}
        part = next((rel[s.get(_RID)] for s in book.iter(_MAIN + "sheet") if s.get("name") == sheet), None)
        if part is None:
            raise KeyError(f"no sheet {sheet!r} in {path}")
        strings = []
        if "xl/sharedStrings.xml" in z.namelist():
            for si in ET.fromstring(z.read("xl/sharedStrings.xml")).iter(_MAIN + "si"):
                strings.append("".join(t.text or "" for t in si.iter(_MAIN + "t")))
        root = ET.fromstring(z.read("xl/" + part.split("xl/")[-1].lstrip("/")))
    found = []
    for c in root.iter(_MAIN + "c"):
        letters, digits = re.match(r"([A-Z]+)(\d+)", c.get("r")).groups()
        col = 0
        for ch in letters:
            col = col * 26 + ord(ch) - 64
        kind, v = c.get("t"), c.find(_MAIN + "v")
        if kind == "inlineStr":
            val = "".join(t.text or "" for t in c.iter(_MAIN + "t"))
        elif v is None or v.text is None or kind == "e":
            continue
        elif kind == "s":
            val = strings[int(v.text)]
        elif kind == "str":
            val = v.text
        else:
            val = float(v.text)
        found.append((int(digits) - 1, col - 1, val))
    grid = np.full((max(f[0] for f in found) + 1, max(f[1] for f in found) + 1), None, dtype=object)
    for r, c, val in found:
        grid[r, c] = val
    return grid


def _serial_to_year_month(serial):
    d = datetime.date(1899, 12, 30) + datetime.timedelta(days=int(round(serial)))      # Excel 1900 date system
    return d.year, d.month


# --------------------------------------------------------------------------------------- series operations
def transform_series(x, tcode):
    """`transform` readin_functions.jl:105-115: 1 level, 2 first difference, 3 second difference, 4 log,
    5 log first difference, 6 log second difference."""
    x = np.asarray(x, float)
    if tcode in (4, 5, 6):
        with np.errstate(invalid="ignore", divide="ignore"):
            x = np.log(x)
    order = {1: 0, 4: 0, 2: 1, 5: 1, 3: 2, 6: 2}[tcode]
    out = np.full(x.shape, np.nan)
    if order == 0:
        out[:] = x
    elif order == 1:
        out[1:] = x[1:] - x[:-1]
    else:
        out[2:] = x[2:] - 2.0 * x[1:-1] + x[:-2]
    return out


def adjust_outliers(x, outliercode, io_method=4):
    """`adjust_outlier!` readin_functions.jl:127-198 (in place).  Outliers = observations further than
    4.5 (code 1) or 3 (code 2) interquartile ranges from the median; replacement rule `io_method`
    (0 missing, 1 threshold, 2 median, 3 local median +-3, 4 one-sided median of the 5 preceding)."""
    if outliercode == 0:
        return x
    k = {1: 4.5, 2: 3.0}[outliercode]
    obs = x[~np.isnan(x)]
    med = np.median(obs)
    q1, q3 = np.quantile(obs, [0.25, 0.75])
    cut = k * (q3 - q1)
    with np.errstate(invalid="ignore"):
        hits = np.flatnonzero(np.abs(x - med) > cut)
    for i in hits:
        if io_method == 0:
            x[i] = np.nan
        elif io_method == 1:
            x[i] = med + np.sign(x[i]) * cut
        elif io_method == 2:
            x[i] = med
        elif io_method in (3, 4):
            win = x[max(0, i - 3):i + 4] if io_method == 3 else x[max(0, i - 5):i + 1]
            x[i] = np.median(win[~np.isnan(win)])
        else:
            raise ValueError(f"io_method {io_method}")
    return x


def biweight_trend(X, bandwidth):
    """`bi_weight_filter` readin_functions.jl:335-348 for all columns at once: kernel-weighted local mean
    with weights 15/16 (1 - (dt/bw)^2)^2 for |dt| < bw, renormalised over the observed periods of each series."""
    T = X.shape[0]
    dt = (np.arange(T)[:, None] - np.arange(T)[None, :]) / float(bandwidth)
    K = np.where(np.abs(dt) < 1.0, 15.0 / 16.0 * (1.0 - dt ** 2) ** 2, 0.0)
    obs = ~np.isnan(X)
    num = K @ np.where(obs, X, 0.0)
    den = K @ obs.astype(float)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(obs, num / den, np.nan)


# --------------------------------------------------------------------------------------- one sheet
def _n_periods(first, last, per_year):
    """MonthlyData / QuarterlyData constructors, readin_functions.jl:29-36."""
    return per_year * (last[0] - first[0] - 1) + last[1] + (per_year - first[1] + 1)


def _read_block(path, sheet, monthly, nobs, ns, datatype, correct_outlier, io_method, cat_include):
    """`readin_monthly_data` / `readin_quarterly_data` readin_functions.jl:200-253 with the header layouts of
    :258-283 (row 0 names, then 2 description rows, then the code rows, then `nobs` data rows)."""
    g = read_xlsx_sheet(path, sheet)
    ncodes = 6 if monthly else 5
    head = 1 + 2 + ncodes
    body = g[head:head + nobs, :ns + 1]
    names = [str(s).upper() for s in g[0, 1:ns + 1]]
    code_row = {k: r for k, r in zip(("t", "defl", "outl", "incl", "cat"), range(4, 9) if monthly else range(3, 8))}
    ints = lambda r: np.array([int(v) for v in g[r, 1:ns + 1]])                      # noqa: E731
    tcode, defl, outl, incl = (ints(code_row[k]) for k in ("t", "defl", "outl", "incl"))
    cat = np.array([float(v) for v in g[code_row["cat"], 1:ns + 1]])
    dates = [_serial_to_year_month(v) for v in body[:, 0]]
    D = np.array([[v if isinstance(v, float) else np.nan for v in row[1:]] for row in body], float)
    col = names.index
    if monthly:
        # standardize_killian! :306-313 (sample standard deviation) and the monthly deflators :285-292
        k = col("GLOBAL_ACT"); ok = ~np.isnan(D[:, k])
        D[ok, k] = (D[ok, k] - D[ok, k].mean()) / D[ok, k].std(ddof=1)
        deflator = {1: D[:, col("PCEPI")].copy(), 2: D[:, col("PCEPILFE")].copy()}
    else:                                                                             # :294-301
        deflator = {1: D[:, col("PCECTPI")].copy(), 2: D[:, col("JCXFE")].copy(), 3: D[:, col("GDPCTPI")].copy()}
    use = incl != 0
    if datatype == "Real":                                                            # :254-256
        use &= np.isin(np.floor(cat), cat_include)
    use = np.flatnonzero(use)
    X = D[:, use].copy()
    for j, s in enumerate(use):                                                       # deflate_series! :40-76
        if defl[s] in deflator:
            X[:, j] = X[:, j] / deflator[defl[s]]
    if monthly:                                                                       # monthly_to_quarterly :83-100
        quarters = sorted({(y, (m + 2) // 3) for y, m in dates})
        which = np.array([quarters.index((y, (m + 2) // 3)) for y, m in dates])
        X = np.stack([X[which == q].mean(axis=0) for q in range(len(quarters))])      # NaN if a month is missing
    else:
        quarters = [(y, (m + 2) // 3) for y, m in dates]
    for j, s in enumerate(use):                                                       # transform! :117-125, outliers :247
        X[:, j] = transform_series(X[:, j], tcode[s])
        if correct_outlier:
            adjust_outliers(X[:, j], outl[s], io_method)
    return X, quarters, cat[use], incl[use], [names[s] for s in use]


# --------------------------------------------------------------------------------------- the panel
def readin_data(path, datatype="All", biweight=100.0, m_first=(1959, 1), m_last=(2014, 12), m_ns=148,
                q_first=(1959, 1), q_last=(2014, 4), q_ns=85, correct_outlier=True, io_method=4,
                cat_include=(1, 2, 3, 5)):
    """`readin_data` readin_functions.jl:355-382 with the notebook's arguments as defaults
    (Stock_Watson.ipynb:143-144, `BiWeight(100)`, datatype `:All` :180 or `:Real` :160)."""
    if datatype not in ("All", "Real"):
        raise ValueError("datatype must be 'All' or 'Real'")
    Xm, qm, cm, im, nm = _read_block(path, "Monthly", True, _n_periods(m_first, m_last, 12), m_ns, datatype,
                                     correct_outlier, io_method, cat_include)
    Xq, qq, cq, iq, nq = _read_block(path, "Quarterly", False, _n_periods(q_first, q_last, 4), q_ns, datatype,
                                     correct_outlier, io_method, cat_include)
    if qm != qq:
        raise ValueError("monthly and quarterly sheets do not cover the same quarters")
    cat = np.concatenate([cm, cq])
    order = np.argsort(cat, kind="stable")                                            # sortperm :368
    unfiltered = np.hstack([Xm, Xq])[:, order]
    data = unfiltered.copy()
    if biweight is not None:                                                          # detrend_var! :317-324
        data = data - biweight_trend(data, biweight)
    names = nm + nq
    return Panel(data, unfiltered, np.concatenate([im, iq])[order], cat[order], [names[i] for i in order], qm,
                 np.array([y + (q - 1) / 4 for y, q in qm]))
