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


def _canonical_correlations(X, Y):
    """MultivariateStats.fit(CCA, X', Y'; method=:svd) with means removed: singular values of Qx'Qy."""
    Xc, Yc = X - X.mean(0), Y - Y.mean(0)
    qx, _ = np.linalg.qr(Xc); qy, _ = np.linalg.qr(Yc)
    return np.linalg.svd(qx.T @ qy, compute_uv=False)


TABLE5_VARS = {   # Stock_Watson.ipynb cell "Table 5" (variable sets A, B, O; C needs the stepwise selection, not restated)
    "A": ["GDPC96", "PAYEMS", "PCECTPI", "FEDFUNDS"],
    "B": ["GDPC96", "PAYEMS", "PCECTPI", "FEDFUNDS", "NAPMPRI", "WPU0561", "CP90_TBILL", "GS10_TB3M"],
    "O": ["OILPROD_SA", "GLOBAL_ACT", "WPU0561", "GDPC96", "PAYEMS", "PCECTPI", "FEDFUNDS", "TWEXMMTH"],
}


@pytest.mark.parametrize("vset", ["A", "B", "O"])
def test_table5_canonical_correlations(panels, notebook_tables, vset):
    """Stock_Watson.ipynb:1250-1261: canonical correlations between a small VAR's variables (residuals) and
    the 8 factors (factor-VAR residuals).  Pins estimate!() end to end including estimate_var! (:444-468) and
    the residual placement `varm.resid` (:464) that the C4 bootstrap resamples.  6 significant digits."""
    names = [str(n) for n in panels["all_names"]]
    m = model(panels["all_bpdata"], panels["all_inclcode"], 8)
    R.estimate(m)
    cols = [names.index(v) for v in TABLE5_VARS[vset]]
    X = panels["all_bpdata"][:, cols]
    v = R.VARModel(X, m.factor_var_model.nlag, m.factor_var_model.withconst, m.factor_var_model.initperiod,
                   m.factor_var_model.lastperiod)
    R.estimate_var(v)
    gold = notebook_tables["table5"][vset]
    ok = ~np.isnan(np.column_stack([X, m.factor])).any(1)
    lev = _canonical_correlations(X[ok], m.factor[ok])
    np.testing.assert_allclose(lev, gold["level"], rtol=5e-5, atol=5e-7)
    ok = ~np.isnan(np.column_stack([v.resid, m.factor_var_model.resid])).any(1)
    res = _canonical_correlations(v.resid[ok], m.factor_var_model.resid[ok])
    np.testing.assert_allclose(res, gold["resid"], rtol=5e-5, atol=5e-7)


@pytest.mark.slow
def test_table4a_instability_rejection_rates(panels, notebook_tables):
    """Stock_Watson.ipynb Table 4(a): share of series whose Chow / QLR statistic (HAC, 6 lags; break 1984Q4) exceeds the
    1 / 5 / 10 % critical values, r = 4 and 8 -- pins compute_chow, compute_qlr, regress_hac, hac, form_hscrc."""
    from scipy.stats import chi2
    qlr_thresh = {4: 4 * np.array([5.12, 4.09, 3.59]), 8: 8 * np.array([3.57, 2.98, 2.69])}
    for r, key in ((4, "chow_qlr_r4"), (8, "chow_qlr_r8")):
        m = model(panels["all_bpdata"], panels["all_inclcode"], r)
        R.estimate_factor(m, computeR2=False)
        chow, qlr = R.instability_tests(m, 104)
        ok = ~np.isnan(chow)
        got = [[np.mean(chow[ok] > chi2.ppf(lv, r)), np.mean(qlr[ok] > th)] for lv, th in zip((0.99, 0.95, 0.9), qlr_thresh[r])]
        np.testing.assert_allclose(np.array(got), np.array(notebook_tables["table4"][key]), atol=1e-6)


@pytest.mark.slow
def test_table4a_fitted_value_correlations(panels, notebook_tables):
    """Stock_Watson.ipynb Table 4(a), lower half: quantiles over the series of cor(yhat_full, yhat_pre) and
    cor(yhat_full, yhat_post) (factors re-estimated on 1959Q3-1984Q4 / 1985Q1-2014Q4), r = 4 and 8."""
    pct = [0.05, 0.25, 0.50, 0.75, 0.95]
    for r, key in ((4, "cor_r4"), (8, "cor_r8")):
        ms = [model(panels["all_bpdata"], panels["all_inclcode"], r, i0, i1) for i0, i1 in ((3, 224), (3, 104), (105, 224))]
        for m in ms:
            R.estimate_factor(m, computeR2=False)
        got = [np.quantile(c[~np.isnan(c)], pct) for c in (R.fitted_value_correlations(ms[0], ms[1], 104),
                                                           R.fitted_value_correlations(ms[0], ms[2], 104))]
        np.testing.assert_allclose(np.array(got), np.array(notebook_tables["table4"][key]), atol=2e-6)
