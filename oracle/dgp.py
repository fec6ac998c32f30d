"""Frozen synthetic data generator (SURVEY.md section 8d / BASELINE.md section 3).  ORACLE / bench input only.

    Lam_ij ~ N(0,1);  f_t = diag(a) f_{t-1} + eta_t, a_j ~ U(0.2,0.8), eta ~ N(0,I_r), burn-in 100;
    e_it ~ N(0, s2_i), s2_i ~ U(0.5,1.5);  x = f Lam' + e, column-standardised (population std).
RNG: numpy Philox, key = (SEED, replication id)  => panel b is identical whatever the GPU count.
The reference has no generator; this definition is the frozen one for every config C2/C3/C5.
"""
import numpy as np

SEED = 20260922


def simulate_panel(N, r, T, rep=0, seed=SEED, missing_frac=0.0, standardize=True):
    rng = np.random.Generator(np.random.Philox(key=[seed, rep]))
    Lam = rng.standard_normal((N, r))
    a = rng.uniform(0.2, 0.8, r)
    s2 = rng.uniform(0.5, 1.5, N)
    eta = rng.standard_normal((T + 100, r))
    e = rng.standard_normal((T, N)) * np.sqrt(s2)
    f = np.zeros(r); F = np.empty((T, r))
    for t in range(T + 100):
        f = a * f + eta[t]
        if t >= 100:
            F[t - 100] = f
    X = F @ Lam.T + e
    if missing_frac > 0:
        X[rng.uniform(size=X.shape) < missing_frac] = np.nan
    if standardize:
        mu = np.nanmean(X, axis=0); sd = np.nanstd(X, axis=0)
        X = (X - mu) / sd
    return X, dict(Lam=Lam, a=a, s2=s2, F=F)


def simulate_batch(B, N, r, T, rep0=0, **kw):
    """(B, T, N) C-order array: panel b is X[b] (T x N row-major)."""
    return np.stack([simulate_panel(N, r, T, rep=rep0 + b, **kw)[0] for b in range(B)])
