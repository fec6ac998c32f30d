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


# ------------------------------------------------------------------------------------------------------------------
# numpy restatement of the DEVICE generator (dynamic_factor_models_b200/csrc/dfm_kernels_rep.cuh): the same frozen DGP
# drawn from a counter-based Philox4x32-10 stream, counter = (element lo, element hi, rep lo, (rep hi << 8) | stream),
# key = seed.  Integer part bit-exact; the Box-Muller transform agrees with the device to libm rounding.  Test infra.
RNG_LAM, RNG_AR, RNG_S2, RNG_ETA, RNG_E, RNG_BIDX, RNG_BETA = range(7)
_M0, _M1, _W0, _W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10: counters as uint64 arrays holding 32-bit values; returns 4 uint64 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF; k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0; p1 = _M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & _MASK; n1 = p1 & _MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & _MASK; n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF; k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def rng_u2(seed, rep, stream, ctr):
    ctr = np.asarray(ctr, dtype=np.uint64)
    o = philox4x32_10(ctr & _MASK, ctr >> np.uint64(32), np.uint64(rep & 0xFFFFFFFF), np.uint64((((rep >> 32) << 8) | stream) & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, seed >> 32)
    u0 = (((o[0] >> np.uint64(5)) << np.uint64(26)) | (o[1] >> np.uint64(6))).astype(np.float64)
    u1 = (((o[2] >> np.uint64(5)) << np.uint64(26)) | (o[3] >> np.uint64(6))).astype(np.float64)
    return (u0 + 0.5) / 9007199254740992.0, (u1 + 0.5) / 9007199254740992.0


def rng_uniform(seed, rep, stream, e):
    return rng_u2(seed, rep, stream, e)[0]


def rng_normal(seed, rep, stream, e):
    e = np.asarray(e, dtype=np.uint64)
    u0, u1 = rng_u2(seed, rep, stream, e >> np.uint64(1))
    rad = np.sqrt(-2.0 * np.log(u0)); ang = 6.283185307179586476925286766559 * u1
    return np.where((e & np.uint64(1)) == 1, rad * np.sin(ang), rad * np.cos(ang))


def simulate_panel_device_stream(N, r, T, rep=0, seed=SEED):
    """The panel dfm_simulate_panels generates for replication id `rep` (T x N, standardised) and its true factors."""
    a = 0.2 + 0.6 * rng_uniform(seed, rep, RNG_AR, np.arange(r))
    eta = rng_normal(seed, rep, RNG_ETA, np.arange((T + 100) * r)).reshape(T + 100, r)
    f = np.zeros(r); F = np.empty((T, r))
    for t in range(T + 100):
        f = a * f + eta[t]
        if t >= 100:
            F[t - 100] = f
    Lam = rng_normal(seed, rep, RNG_LAM, np.arange(N * r)).reshape(N, r)
    sd = np.sqrt(0.5 + rng_uniform(seed, rep, RNG_S2, np.arange(N)))
    e = rng_normal(seed, rep, RNG_E, np.arange(N * T)).reshape(N, T).T          # element i*T + t
    X = e * sd
    for j in range(r):                                                          # same summation order as the kernel
        X = X + F[:, j:j + 1] * Lam[None, :, j]
    mean = X.mean(0)
    X = (X - mean) / np.sqrt(((X - mean) ** 2).mean(0))
    return X, dict(Lam=Lam, a=a, s2=sd ** 2, F=F)


def bootstrap_panel_device_stream(F0, resid, beta, lam, uar_coef, uar_ser, data, rep, seed=SEED, burn=50):
    """The draw dfm_bootstrap_panels generates for replication id `rep` (Tw x ns)."""
    Tw, r = F0.shape; ns, L = uar_coef.shape; K = beta.shape[0]; p = (K - 1) // r; nres = resid.shape[0]
    fs = np.empty((Tw, r)); fs[:p] = F0[:p]
    u = rng_uniform(seed, rep, RNG_BIDX, np.arange(Tw - p))
    idx = np.minimum((u * nres).astype(np.int64), nres - 1)
    for t in range(p, Tw):
        v = beta[0] + resid[idx[t - p]]
        for l in range(1, p + 1):
            for c in range(r):
                v = v + fs[t - l, c] * beta[1 + (l - 1) * r + c]
        fs[t] = v
    ok = ~np.isnan(lam).any(axis=1) & ~np.isnan(uar_ser)
    out = np.full((Tw, ns), np.nan)
    eta = rng_normal(seed, rep, RNG_BETA, np.arange(ns * (Tw + burn))).reshape(ns, Tw + burn)
    for i in np.flatnonzero(ok):
        ul = np.zeros(L)
        for t in range(Tw + burn):
            acc = uar_ser[i] * eta[i, t]
            for l in range(L):
                acc += uar_coef[i, l] * ul[l]
            ul[1:] = ul[:-1]; ul[0] = acc
            if t >= burn:
                v = acc
                for c in range(r):
                    v += fs[t - burn, c] * lam[i, c]
                out[t - burn, i] = v
    out[np.isnan(data)] = np.nan
    return out
