"""Validate the Kalman-EM spec oracle (parity unpinned: no reference code exists) by
brute force and invariants.  CPU only."""
import numpy as np
import pytest

from oracle import kalman_em as K
from oracle.dgp import simulate_panel
from oracle import dfm_ref as R


def _brute_force(X, Lam, Rv, A, Q, P0, p):
    """Joint Gaussian of (z_1..z_T, x_obs): smoothed means/covs and loglik by dense algebra."""
    T, N = X.shape; r = Lam.shape[1]; k = r * p
    M = K.companion(A, r, p); Qt = np.zeros((k, k)); Qt[:r, :r] = Q
    # prior covariance of stacked z
    covs = [P0]
    for t in range(1, T):
        covs.append(M @ covs[-1] @ M.T + Qt)
    Sz = np.zeros((T * k, T * k))
    for s in range(T):
        blk = covs[s]
        for t in range(s, T):
            Sz[t * k:(t + 1) * k, s * k:(s + 1) * k] = blk
            Sz[s * k:(s + 1) * k, t * k:(t + 1) * k] = blk.T
            blk = M @ blk
    H = np.zeros((T * N, T * k))
    for t in range(T):
        H[t * N:(t + 1) * N, t * k:t * k + r] = Lam
    o = ~np.isnan(X).ravel()
    Ho = H[o]; xo = X.ravel()[o]
    Sx = Ho @ Sz @ Ho.T + np.diag(np.tile(Rv, T)[o])
    Kg = Sz @ Ho.T @ np.linalg.inv(Sx)
    zs = (Kg @ xo).reshape(T, k)
    Pz = Sz - Kg @ Ho @ Sz
    sign, ld = np.linalg.slogdet(Sx)
    ll = -0.5 * (len(xo) * np.log(2 * np.pi) + ld + xo @ np.linalg.solve(Sx, xo))
    return zs, Pz, ll


@pytest.mark.parametrize("p,miss", [(1, 0.0), (2, 0.0), (1, 0.2), (2, 0.15)])
def test_estep_matches_brute_force(p, miss):
    rng = np.random.default_rng(5 + p)
    T, N, r = 9, 6, 2; k = r * p
    X, _ = simulate_panel(N, r, T, rep=3, missing_frac=miss)
    Lam = rng.standard_normal((N, r)); Rv = rng.uniform(0.5, 1.5, N)
    A = 0.3 * rng.standard_normal((r, k)); Q = np.eye(r) + 0.1 * np.ones((r, r))
    Qt = np.zeros((k, k)); Qt[:r, :r] = Q
    P0 = K.lyapunov_doubling(K.companion(A, r, p), Qt)
    es = K.e_step(X, Lam, Rv, A, Q, P0, p)
    zs, Pz, ll = _brute_force(X, Lam, Rv, A, Q, P0, p)
    np.testing.assert_allclose(es["zs"], zs, atol=1e-10)
    np.testing.assert_allclose(es["loglik"], ll, rtol=1e-11)
    for t in range(T):
        np.testing.assert_allclose(es["Ps"][t], Pz[t * k:(t + 1) * k, t * k:(t + 1) * k], atol=1e-10)
    # cross moments: S00, S11 from brute force
    S00 = sum(np.outer(zs[t], zs[t]) + Pz[t * k:(t + 1) * k, t * k:(t + 1) * k] for t in range(T - 1))
    S11 = sum(np.outer(zs[t + 1][:r], zs[t]) + Pz[(t + 1) * k:(t + 1) * k + r, t * k:(t + 1) * k]
              for t in range(T - 1))
    np.testing.assert_allclose(es["S00"], S00, atol=1e-9)
    np.testing.assert_allclose(es["S11"], S11, atol=1e-9)


@pytest.mark.parametrize("p,miss", [(1, 0.0), (2, 0.1)])
def test_em_monotone_loglik_and_subspace(p, miss):
    N, r, T = 40, 3, 120
    X, tr = simulate_panel(N, r, T, rep=11, missing_frac=miss)
    # init: PCA factors of zero-filled panel
    F0 = R.pca_score(np.nan_to_num(X), r)
    Lam, Rv, A, Q = K.init_from_factors(X, F0, p)
    out = K.em_kalman(X, Lam, Rv, A, Q, p=p, max_iter=25, tol=0.0)
    d = np.diff(out["loglik"])
    assert (d >= -1e-8 * np.abs(out["loglik"][:-1])).all(), d
    # smoothed factors span the true factor space
    Fh, Ft = out["F"], tr["F"]
    proj = Fh @ np.linalg.lstsq(Fh, Ft, rcond=None)[0]
    r2 = 1 - ((Ft - proj) ** 2).sum() / ((Ft - Ft.mean(0)) ** 2).sum()
    assert r2 > 0.9, r2


def test_lyapunov_doubling_fixed_point():
    rng = np.random.default_rng(0)
    r, p = 3, 2; k = r * p
    A = 0.25 * rng.standard_normal((r, k)); M = K.companion(A, r, p)
    Qt = np.zeros((k, k)); Qt[:r, :r] = np.eye(r)
    P = K.lyapunov_doubling(M, Qt)
    np.testing.assert_allclose(P, M @ P @ M.T + Qt, atol=1e-12)
