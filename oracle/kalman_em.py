"""FP64 spec of the state-space (Parametric) DFM estimator: Kalman filter + RTS smoother + EM.
ORACLE / TEST INFRASTRUCTURE ONLY.

*** PARITY UNPINNED ***  The reference contains no Kalman filter, smoother, parametric EM or
bootstrap (`struct Parametric <: EstimationMethod end` is an empty placeholder at
dfm_functions.ipynb:23; only estimate!(::NonParametric) exists, :530-543).  This file therefore
IS the specification of row a' of SURVEY.md section 8; it is validated by (i) a brute-force joint-Gaussian
computation on tiny problems, (ii) monotone log-likelihood, (iii) recovery of the DGP subspace --
see tests/test_oracle_kalman.py.

Model (the reference's own companion form, dfm_functions.ipynb:477-492, as transition equation):
    z_t = M z_{t-1} + [eta_t; 0],   eta_t ~ N(0, Q)      z_t = [f_t; f_{t-1}; ...; f_{t-p+1}],  k = r p
    x_t = Lam f_t + e_t,            e_t ~ N(0, diag(R))   (x standardized; missing = NaN)
    z_1 ~ N(0, P0)                  P0 fixed during EM (so the M-step below is the exact EM update)
M = [A_1 ... A_p ; I 0].  Parameters updated by EM: Lam (N x r), R (N), A = [A_1..A_p] (r x k), Q (r x r).

Filter update is in information form (only r x r / k x k systems):
    C_t = sum_{i obs} lam_i lam_i' / R_i     b_t = sum_{i obs} lam_i x_it / R_i     q_t = sum_{i obs} x_it^2 / R_i
"""
import numpy as np
from scipy.linalg import solve_triangular, cho_factor, cho_solve

LOG2PI = float(np.log(2.0 * np.pi))


def companion(A, r, p):
    k = r * p
    M = np.zeros((k, k)); M[:r] = A
    if p > 1:
        M[r:, :k - r] = np.eye(k - r)
    return M


def lyapunov_doubling(M, Qt, steps=12):
    """P = sum_j M^j Qt M'^j via doubling (2^steps terms).  Default prior P0 (stationary cov)."""
    P, Mj = Qt.copy(), M.copy()
    for _ in range(steps):
        P = P + Mj @ P @ Mj.T
        Mj = Mj @ Mj
    return 0.5 * (P + P.T)


def e_step(X, Lam, R, A, Q, P0, p):
    """One Kalman filter + RTS smoother pass.  Returns dict of smoothed moments and loglik."""
    T, N = X.shape
    r = Lam.shape[1]; k = r * p
    M = companion(A, r, p)
    Qt = np.zeros((k, k)); Qt[:r, :r] = Q
    use = ~np.isnan(Lam).any(axis=1) & ~np.isnan(R)
    obs = ~np.isnan(X) & use[None, :]
    X0 = np.where(obs, X, 0.0)
    Lam0 = np.where(use[:, None], Lam, 0.0)
    Rinv = np.where(use, 1.0 / np.where(use, R, 1.0), 0.0)
    W = Lam0 * Rinv[:, None]                                   # N x r
    B = X0 @ W                                                 # T x r    b_t
    qv = (X0 ** 2) @ Rinv                                      # T        q_t
    logR = np.where(use, np.log(np.where(use, R, 1.0)), 0.0)
    sumlogR = obs @ logR                                       # T
    nobs_t = obs.sum(axis=1)
    balanced = obs.all(axis=0)[use].all() if use.any() else True
    Cfull = Lam0.T @ W
    zp = np.zeros((T, k)); zf = np.zeros((T, k))
    Pp = np.zeros((T, k, k)); Pf = np.zeros((T, k, k))
    ll = 0.0
    Ir = np.eye(r)
    for t in range(T):
        if t == 0:
            zp[t] = 0.0; Pp[t] = P0
        else:
            zp[t] = M @ zf[t - 1]
            Pt = M @ Pf[t - 1] @ M.T + Qt
            Pp[t] = 0.5 * (Pt + Pt.T)
        if balanced:
            C = Cfull
        else:
            o = obs[t]
            C = Lam0[o].T @ W[o]
        L = np.linalg.cholesky(Pp[t][:r, :r])
        S = Ir + L.T @ C @ L
        Ls = np.linalg.cholesky(0.5 * (S + S.T))
        Tm = solve_triangular(L, Pp[t][:r, :], lower=True)     # r x k
        Wm = solve_triangular(Ls, Tm, lower=True)
        Pft = Pp[t] - Tm.T @ Tm + Wm.T @ Wm
        Pf[t] = 0.5 * (Pft + Pft.T)
        zpf = zp[t][:r]
        g = B[t] - C @ zpf
        zf[t] = zp[t] + Pf[t][:, :r] @ g
        logdetF = sumlogR[t] + 2.0 * np.log(np.diag(Ls)).sum()
        quad = qv[t] - 2.0 * zpf @ B[t] + zpf @ C @ zpf - g @ Pf[t][:r, :r] @ g
        ll += -0.5 * (nobs_t[t] * LOG2PI + logdetF + quad)
    zs = np.zeros((T, k)); Ps = np.zeros((T, k, k))
    zs[-1] = zf[-1]; Ps[-1] = Pf[-1]
    S00 = np.zeros((k, k)); S11 = np.zeros((r, k)); Sff2 = np.zeros((r, r))
    for t in range(T - 2, -1, -1):
        cf = cho_factor(Pp[t + 1], lower=True)
        J = cho_solve(cf, M @ Pf[t]).T                         # Pf M' Pp^-1
        zs[t] = zf[t] + J @ (zs[t + 1] - zp[t + 1])
        Pst = Pf[t] + J @ (Ps[t + 1] - Pp[t + 1]) @ J.T
        Ps[t] = 0.5 * (Pst + Pst.T)
        Pc = Ps[t + 1] @ J.T                                   # cov(z_{t+1}, z_t | T)
        S11 += np.outer(zs[t + 1][:r], zs[t]) + Pc[:r, :]
        S00 += np.outer(zs[t], zs[t]) + Ps[t]
        Sff2 += np.outer(zs[t + 1][:r], zs[t + 1][:r]) + Ps[t + 1][:r, :r]
    return dict(zs=zs, Ps=Ps, zf=zf, Pf=Pf, zp=zp, Pp=Pp, loglik=ll, S00=S00, S11=S11, Sff2=Sff2,
                obs=obs, use=use)


def m_step(X, es, r, p):
    """Exact EM update given smoothed moments (zero-mean model, fixed P0)."""
    T, N = X.shape
    zs, Ps, obs, use = es["zs"], es["Ps"], es["obs"], es["use"]
    Fs = zs[:, :r]
    E = Fs[:, :, None] * Fs[:, None, :] + Ps[:, :r, :r]        # T x r x r   E[f f' | T]
    X0 = np.where(obs, X, 0.0)
    Sxf = X0.T @ Fs                                            # N x r
    Sxx = (X0 ** 2).sum(axis=0)
    Ti = obs.sum(axis=0)
    Sff_i = np.einsum("ti,tab->iab", obs.astype(float), E)     # N x r x r
    Lam = np.full((N, r), np.nan); R = np.full(N, np.nan)
    for i in range(N):
        if use[i] and Ti[i] > 0:
            Lam[i] = np.linalg.solve(Sff_i[i], Sxf[i])
            R[i] = (Sxx[i] - 2.0 * Lam[i] @ Sxf[i] + Lam[i] @ Sff_i[i] @ Lam[i]) / Ti[i]
    A = np.linalg.solve(es["S00"], es["S11"].T).T              # S11 S00^-1
    Q = (es["Sff2"] - A @ es["S11"].T) / (T - 1)
    return Lam, R, A, 0.5 * (Q + Q.T)


def em_kalman(X, Lam, R, A, Q, p=1, P0=None, max_iter=50, tol=0.0):
    """EM loop.  Each iteration = E-step under current params (gives loglik of those params)
    followed by an M-step.  Stops after iteration j>=2 when
        |ll_j - ll_{j-1}| <= tol * (|ll_j| + |ll_{j-1}|) / 2 .
    Returns params after the last M-step, the smoothed moments of the last E-step, loglik path."""
    r = Lam.shape[1]; k = r * p
    if P0 is None:
        Qt = np.zeros((k, k)); Qt[:r, :r] = Q
        P0 = lyapunov_doubling(companion(A, r, p), Qt)
    lls = []
    es = None
    for it in range(1, max_iter + 1):
        es = e_step(X, Lam, R, A, Q, P0, p)
        lls.append(es["loglik"])
        Lam, R, A, Q = m_step(X, es, r, p)
        if it >= 2 and abs(lls[-1] - lls[-2]) <= tol * 0.5 * (abs(lls[-1]) + abs(lls[-2])):
            break
    return dict(Lam=Lam, R=R, A=A, Q=Q, P0=P0, F=es["zs"][:, :r], PsF=es["Ps"][:, :r, :r],
                loglik=np.array(lls), iters=len(lls), es=es)


def init_from_factors(Xs, F, p=1):
    """Initial (Lam, R, A, Q) from standardized panel Xs and factor estimates F (e.g. the ALS
    factors): per-series OLS on F (no const), residual variance, VAR(p) without constant."""
    T, N = Xs.shape; r = F.shape[1]
    Lam = np.full((N, r), np.nan); R = np.full(N, np.nan)
    for i in range(N):
        o = ~np.isnan(Xs[:, i])
        if o.sum() > r:
            G = F[o].T @ F[o]
            Lam[i] = np.linalg.solve(G, F[o].T @ Xs[o, i])
            e = Xs[o, i] - F[o] @ Lam[i]
            R[i] = e @ e / o.sum()
    Y = F[p:]
    Z = np.hstack([F[p - l:T - l] for l in range(1, p + 1)])
    A = np.linalg.solve(Z.T @ Z, Z.T @ Y).T
    E = Y - Z @ A.T
    Q = E.T @ E / (T - p)
    return Lam, R, A, 0.5 * (Q + Q.T)
