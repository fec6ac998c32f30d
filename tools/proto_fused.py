"""Numpy prototype of the FUSED fast path algorithm (p = 1, balanced panel): validates the math of
dfm_fast.cu against oracle/kalman_em.py before it is written in CUDA.  Dev tool, not product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import kalman_em as K, dfm_ref as R
from oracle.dgp import simulate_panel

LOG2PI = np.log(2 * np.pi)


def gj_inv(A):
    """Gauss-Jordan inverse of SPD A without pivoting + log det (what a warp does in registers)."""
    A = A.copy(); n = len(A); ld = 0.0
    for p in range(n):
        piv = A[p, p]; ld += np.log(piv); d = 1.0 / piv
        row = A[p].copy(); col = A[:, p].copy()
        A -= np.outer(col, row) * d
        A[p] = row * d; A[:, p] = -col * d; A[p, p] = d
    # the sign convention above yields inv with mixed signs; fix: standard GJ
    return A, ld


def gj_inv2(A):
    n = len(A); M = A.copy(); ld = 0.0
    for p in range(n):
        piv = M[p, p]; ld += np.log(piv); d = 1.0 / piv
        row = M[p].copy(); col = M[:, p].copy()
        for i in range(n):
            for j in range(n):
                if i != p and j != p: M[i, j] = M[i, j] - col[i] * row[j] * d
        for j in range(n):
            if j != p: M[p, j] = row[j] * d
        for i in range(n):
            if i != p: M[i, p] = -col[i] * d
        M[p, p] = d
    return M, ld


def e_step_fused(X, Lam, Rv, A, Q, P0, eps=1e-14):
    T, N = X.shape; r = Lam.shape[1]
    M = A
    W = Lam / Rv[:, None]; C = Lam.T @ W; slr = np.log(Rv).sum()
    B = X @ W; qv = (X ** 2) @ (1 / Rv)
    # ---- forward covariance chain with freeze
    Pf = []; Phi = []; ld = []; Pp = [P0]; J = []
    nE = T
    t = 0
    frozen_at = None
    while t < T:
        Pi, ldp = gj_inv2(Pp[t]); Pft, ldw_neg = gj_inv2(Pi + C)     # ldw_neg = log|Pi + C|
        Pf.append(Pft); ld.append(ldp + ldw_neg)
        G = Pft @ Pi; Phi.append(G @ M)
        if t >= 1: J.append(Pf[t - 1] @ M.T @ Pi)
        Pn = M @ Pft @ M.T + Q; Pn = 0.5 * (Pn + Pn.T); Pp.append(Pn)
        if frozen_at is not None and t == frozen_at + 1:
            nE = t + 1; break
        if frozen_at is None and np.abs(Pn - Pp[t]).max() <= eps * np.abs(Pp[t]).max():
            frozen_at = t
        t += 1
    frozen = nE < T
    Jinf = Pf[nE - 1] @ M.T @ gj_inv2(Pp[nE - 1])[0] if frozen else None   # Pi_inf = inv(Pp_{nE-1}) (== Pp_nE)
    getPf = lambda t: Pf[t] if t < nE else Pf[nE - 1]
    getPhi = lambda t: Phi[t] if t < nE else Phi[nE - 1]
    getld = lambda t: ld[t] if t < nE else ld[nE - 1]
    getPp = lambda t: Pp[t] if t < nE else Pp[nE - 1]
    getJ = lambda t: J[t] if t < nE - 1 else Jinf
    # ---- forward means
    zf = np.zeros((T, r))
    zf[0] = getPf(0) @ B[0]
    for t in range(1, T):
        zf[t] = getPhi(t) @ zf[t - 1] + getPf(t) @ B[t]
    # ---- loglik (parallel over t)
    ll = 0.0
    for t in range(T):
        zp = M @ zf[t - 1] if t else np.zeros(r)
        Czp = C @ zp
        quad = qv[t] - 2 * zp @ B[t] + zp @ Czp - (B[t] - Czp) @ (zf[t] - zp)
        ll += -0.5 * (N * LOG2PI + slr + getld(t) + quad)
    # ---- backward means
    zs = zf.copy()
    for t in range(T - 2, -1, -1):
        zs[t] = zf[t] + getJ(t) @ (zs[t + 1] - M @ zf[t])
    # ---- backward covariances with freeze; sums of covariance parts
    SP_all = np.zeros((r, r)); SP_00 = np.zeros((r, r)); SP_ff2 = np.zeros((r, r)); SP_11 = np.zeros((r, r))
    Ps_next = getPf(T - 1).copy()
    SP_all += Ps_next; SP_ff2 += Ps_next
    Ps_list = {T - 1: Ps_next}
    t = T - 2
    lo = nE - 1 if frozen else T      # frozen region for the triple is t >= lo
    while t >= 0:
        Jt = getJ(t)
        Ps = getPf(t) + Jt @ (Ps_next - getPp(t + 1)) @ Jt.T; Ps = 0.5 * (Ps + Ps.T)
        SP_11 += Ps_next @ Jt.T
        SP_all += Ps; SP_00 += Ps
        if t >= 1: SP_ff2 += Ps
        Ps_list[t] = Ps
        conv = frozen and t > lo and np.abs(Ps - Ps_next).max() <= eps * np.abs(Ps).max()
        Ps_next = Ps
        if conv:
            cnt = t - lo            # indices lo .. t-1 all equal Ps (and cross term Ps J')
            SP_all += cnt * Ps; SP_00 += cnt * Ps; SP_ff2 += (cnt if lo >= 1 else cnt - 1) * Ps
            SP_11 += cnt * (Ps @ Jinf.T)
            for s in range(lo, t): Ps_list[s] = Ps
            t = lo - 1
        else:
            t -= 1
    # ---- mean parts
    SffAll = zs.T @ zs + SP_all
    S00 = zs[:-1].T @ zs[:-1] + SP_00
    Sff2 = zs[1:].T @ zs[1:] + SP_ff2
    S11 = zs[1:].T @ zs[:-1] + SP_11
    return dict(zs=zs, ll=ll, SffAll=SffAll, S00=S00, S11=S11, Sff2=Sff2, nE=nE, Ps=Ps_list)


if __name__ == "__main__":
    for (N, r, T) in ((40, 3, 150), (200, 8, 500)):
        X, _ = simulate_panel(N, r, T, rep=1)
        Lam, Rv, A, Q = K.init_from_factors(X, R.pca_score(X, r), 1)
        P0 = K.lyapunov_doubling(A, Q)
        es = K.e_step(X, Lam, Rv, A, Q, P0, 1)
        fu = e_step_fused(X, Lam, Rv, A, Q, P0)
        E = es["zs"][:, :, None] * es["zs"][:, None, :] + es["Ps"]
        print(N, r, T, "nE", fu["nE"], "zs", np.abs(fu["zs"] - es["zs"]).max(), "ll", abs(fu["ll"] - es["loglik"]) / abs(es["loglik"]),
              "S00", np.abs(fu["S00"] - es["S00"]).max(), "S11", np.abs(fu["S11"] - es["S11"]).max(),
              "Sff2", np.abs(fu["Sff2"] - es["Sff2"]).max(), "SffAll", np.abs(fu["SffAll"] - E.sum(0)).max(),
              "Ps", max(np.abs(fu["Ps"][t] - es["Ps"][t]).max() for t in range(T)))
