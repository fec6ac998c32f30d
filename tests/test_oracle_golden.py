"""Pin the oracle (oracle/dfm_ref.py) against the only golden values the reference has:
the stored outputs of Stock_Watson.ipynb.  CPU only."""
import numpy as np
import pytest

from oracle import dfm_ref as R

CFG = dict(nt_min_f=20, nt_min_fl=40, tol=1e-8, n_uarlag=4, n_factorlag=4)   # Stock_Watson.ipynb:245-251


def model(data, incl, r, i0=3, i1=224):
    return R.DFMModel(data, incl, CFG["nt_min_f"], CFG["nt_min_fl"], i0, i1, 0, r, CFG["tol"],
                      CFG["n_uarlag"], CFG["n_factorlag"])


def _table2(data, incl, nmax, gold):
    tr, bn = [], []
    for r in range(1, nmax + 2):
        m = model(data, incl, r)
        R.estimate_factor(m, computeR2=False)
        tr.append(1 - m.fes.ssr / m.fes.tss); bn.append(R.bai_ng_criterion(m))
    tr = np.array(tr); marg = np.diff(np.concatenate([[0], tr])); ah = marg[:-1] / marg[1:]
    got = np.column_stack([np.arange(1, nmax + 1), tr[:nmax], marg[:nmax], bn[:nmax], ah[:nmax]])
    # values were printed with round(., digits=3): allow half a unit in the last place + slack
    np.testing.assert_allclose(got, np.array(gold), atol=6e-4)


def test_table2A_real_panel(panels, notebook_tables):
    """Stock_Watson.ipynb:569-577 (Real panel N=58, r=1..5)."""
    _table2(panels["real_bpdata"], panels["real_inclcode"], 5, notebook_tables["table2A"])


def test_table2B_all_panel(panels, notebook_tables):
    """Stock_Watson.ipynb:616-629 (All panel N=139, r=1..10; r=8 -> 0.501 / -0.223)."""
    _table2(panels["all_bpdata"], panels["all_inclcode"], 10, notebook_tables["table2B"])


@pytest.mark.slow
def test_table2C_amengual_watson(panels, notebook_tables):
    """Stock_Watson.ipynb:669-683: AW ICp for static 1..10 x dynamic 1..10."""
    gold = np.array(notebook_tables["table2C"])[:, 1:]
    got = np.full((10, 10), np.nan)
    for ns_ in range(1, 11):
        m = model(panels["all_bpdata"], panels["all_inclcode"], ns_)
        R.estimate_factor(m, computeR2=False)
        aw, _, _ = R.amengual_watson_test(m, 4)
        got[:ns_, ns_ - 1] = aw
    mask = ~np.isnan(gold)
    assert (np.isnan(got) == np.isnan(gold)).all()
    np.testing.assert_allclose(got[mask], gold[mask], atol=6e-4)


def test_table3_series_r2(panels, notebook_tables):
    """Stock_Watson.ipynb:991-1017: per-series R2 of estimate!() for r in {1,2,3,8,9,10},
    first 13 / last 12 series, 6 significant digits."""
    t3 = notebook_tables["table3_visible"]
    gold = np.array(t3["values"])
    rows = list(range(13)) + list(range(207 - 12, 207))
    for c, r in enumerate(t3["cols"]):
        m = model(panels["all_bpdata"], panels["all_inclcode"], r)
        R.estimate_factor(m, computeR2=False)
        R.estimate_factor_loading(m)
        np.testing.assert_allclose(m.r2[rows], gold[:, c], rtol=2e-5, atol=1e-7)
