"""GPU: replication-level drivers (configs C4 / C5 in miniature) vs the oracle."""
import numpy as np
import pytest

import parity_checks as P
from oracle import dfm_ref as R, kalman_em as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from dynamic_factor_models_b200 import Library
    L = Library()
    yield L
    L.close()


def test_monte_carlo_em_matches_oracle(lib):
    from dynamic_factor_models_b200 import replicate
    n_rep, N, r, T, iters = 6, 40, 3, 120, 8
    rec = replicate.monte_carlo_em(lib, n_rep, N, r, T, em_iters=iters)
    assert rec.shape == (n_rep, 4) and (rec[:, 2] == 0).all() and (rec[:, 1] == iters).all()
    for b in (0, n_rep - 1):
        X = replicate.simulate_panel(N, r, T, rep=b, lib=lib)
        m = R.DFMModel(X, np.ones(N, int), 20, 40, 1, T, 0, r, 1e-8, 4, 1)
        R.estimate_factor(m, max_iter=1, computeR2=False)
        Lam, Rv, A, Q = K.init_from_factors(m.xs, m.factor, 1)
        ref = K.em_kalman(m.xs, Lam, Rv, A, Q, p=1, max_iter=iters)
        np.testing.assert_allclose(rec[b, 0], ref["loglik"][-1], rtol=1e-9)
        np.testing.assert_allclose(rec[b, 3], 1 - m.fes.ssr / m.fes.tss, rtol=1e-9)


def test_bootstrap_irf_c1(lib, panels):
    """C4 in miniature: 4 bootstrap replications of the hom_fac_1 model; one replication is
    re-estimated with the oracle pipeline and compared."""
    import dynamic_factor_models_b200 as D
    from dynamic_factor_models_b200 import replicate
    g = P.gpu_model(panels["all_bpdata"], panels["all_inclcode"], 4)
    D.estimate(g, lib=lib)
    irfs, bands = replicate.bootstrap_irf(lib, g, 4, H=8)
    assert irfs.shape == (4, 4, 8, 4) and np.isfinite(irfs).all()
    assert (bands[5] <= bands[95] + 1e-12).all()
    for q in (5, 50, 95):                                        # device percentile kernel == numpy.percentile
        np.testing.assert_allclose(bands[q], np.percentile(irfs, q, axis=0), rtol=1e-12, atol=1e-14)
    # oracle re-estimation of replication 2 (the fused call resamples the ESTIMATION series only)
    check_bootstrap_replication(lib, g, panels, irfs, 2)


def check_bootstrap_replication(lib, g, panels, irfs, rep, H=8):
    i0, i1 = g.initperiod, g.lastperiod
    v = g.factor_var_model; p = v.nlag; incl = g.inclcode == 1
    Xs = lib.bootstrap_panels(g.factor[i0 - 1:i1], v.resid[i0 - 1:i1][p:], v.betahat, g.lambda_[incl], g.uar_coef[incl], g.uar_ser[incl],
                              g.data[i0 - 1:i1][:, incl], rep, 1, 20260922)[0]
    full = np.full_like(panels["all_bpdata"], np.nan); full[i0 - 1:i1, np.flatnonzero(incl)] = Xs
    m = P.ref_model(full, panels["all_inclcode"], g.nfac_t)
    R.estimate_factor(m, computeR2=False); R.estimate_var(m.factor_var_model)
    F0 = g.factor[i0 - 1:i1]
    s = np.sign((m.factor[i0 - 1:i1] * F0).sum(0)); s[s == 0] = 1
    ref = R.impulse_response(m.factor_var_model, list(range(g.nfac_t)), H) * s[:, None, None] * s[None, None, :]
    np.testing.assert_allclose(irfs[rep], ref, rtol=1e-5, atol=1e-8)


def test_generators_are_shard_invariant_on_device(lib):
    """Panel b is bit-identical for world = 1 and 8 (device-resident generation, shards of dfm_shard_range)."""
    from dynamic_factor_models_b200 import replicate
    whole = lib.simulate_panels(0, 24, 40, 4, 100, replicate.SEED)
    parts = []
    for rank in range(8):
        b, e = lib.shard_range(24, rank, 8)
        parts.append(lib.simulate_panels(b, e - b, 40, 4, 100, replicate.SEED))
    assert np.array_equal(np.concatenate(parts), whole)
