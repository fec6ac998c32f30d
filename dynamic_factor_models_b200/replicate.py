"""Replication-level drivers -- row (e) of SURVEY.md section 8 and configs C4 / C5 of BASELINE.json.

Replications (Monte-Carlo panels, bootstrap draws) are independent: rank g of G owns the contiguous
shard dfm_shard_range(n_rep, g, G), every replication's random stream is keyed by its replication
id only (numpy Philox, key = (seed, id)) so results do not depend on the GPU count, the per-shard
work is one batched call into the CUDA library, and the path's single collective is one all-gather
of the per-replication records at the very end (torch.distributed: NCCL over NVLink on GPUs, gloo in
the CPU tests).  The reference has no bootstrap / Monte-Carlo / RNG code at all.
"""
import numpy as np

SEED = 20260922


def simulate_panel(N, r, T, rep=0, seed=SEED):
    """Frozen synthetic DGP of SURVEY.md 8d: Lam~N(0,1); f_t = diag(a) f_{t-1} + eta_t, a~U(.2,.8),
    burn-in 100; e_it~N(0,s2_i), s2_i~U(.5,1.5); column-standardised (population std)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, rep]))
    Lam = rng.standard_normal((N, r)); a = rng.uniform(0.2, 0.8, r); s2 = rng.uniform(0.5, 1.5, N)
    eta = rng.standard_normal((T + 100, r)); e = rng.standard_normal((T, N)) * np.sqrt(s2)
    f = np.zeros(r); F = np.empty((T, r))
    for t in range(T + 100):
        f = a * f + eta[t]
        if t >= 100:
            F[t - 100] = f
    X = F @ Lam.T + e
    return (X - X.mean(0)) / X.std(0)


def gather_records(rec_local, n_rep, rank, world, lib):
    """All-gather per-replication records (n_local, d) -> (n_rep, d) on every rank.  One collective."""
    if world == 1:
        return np.asarray(rec_local)
    import torch
    import torch.distributed as dist
    b, e = lib.shard_range(n_rep, rank, world)
    nmax = max(lib.shard_range(n_rep, g, world)[1] - lib.shard_range(n_rep, g, world)[0] for g in range(world))
    d = rec_local.shape[1]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.full((nmax, d), float("nan"), dtype=torch.float64, device=dev)
    buf[:e - b] = torch.from_numpy(np.ascontiguousarray(rec_local)).to(dev)
    out = torch.empty((world * nmax, d), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, nmax, d)
    return np.concatenate([out[g, :lib.shard_range(n_rep, g, world)[1] - lib.shard_range(n_rep, g, world)[0]] for g in range(world)])


def monte_carlo_em(lib, n_rep, N, r, T, p=1, em_iters=50, tol=0.0, rank=0, world=1, seed=SEED):
    """C5: EM on n_rep simulated panels.  Record per replication: [final loglik, iterations, status,
    trace R2 of the ALS start].  Returns the gathered (n_rep, 4) array (identical on all ranks)."""
    b, e = lib.shard_range(n_rep, rank, world)
    X = np.stack([simulate_panel(N, r, T, rep=i, seed=seed) for i in range(b, e)]) if e > b else np.empty((0, T, N))
    rec = np.empty((e - b, 4))
    if e > b:
        als = lib.estimate_factor(X, r, max_iter=1, compute_r2=False)
        Lam, R, A, Q = lib.em_init_from_factors(X, als["F"], p)
        em = lib.em_kalman(X, Lam, R, A, Q, p=p, max_iter=em_iters, tol=tol, want_PF=False)
        it = np.asarray(em["iters"]).reshape(-1)
        ll = np.asarray(em["loglik"]).reshape(e - b, em_iters)
        rec[:, 0] = ll[np.arange(e - b), it - 1]; rec[:, 1] = it; rec[:, 2] = np.asarray(em["status"]).reshape(-1)
        st = als["stats"] if isinstance(als["stats"], list) else [als["stats"]]
        rec[:, 3] = [1 - s["ssr"] / s["tss"] for s in st]
    return gather_records(rec, n_rep, rank, world, lib)


def bootstrap_panels(m, ids, seed=SEED, burn=50):
    """Residual bootstrap of a fitted non-parametric model `m` (api.DFMModel after estimate()):
    resample the factor-VAR residuals with replacement and rebuild f* through the VAR, draw the
    idiosyncratic AR(n_uarlag) processes from (uar_coef, uar_ser), x* = Lam f* + u*, original missing
    pattern re-imposed.  Returns (len(ids), T_w, ns)."""
    i0, i1 = m.initperiod, m.lastperiod
    v = m.factor_var_model
    F = m.factor[i0 - 1:i1]; Tw, r = F.shape; p = v.nlag
    resid = v.resid[i0 - 1:i1][p:]
    beta = v.betahat                                     # [const; lag1; ...; lagp] x r
    mask = np.isnan(m.data[i0 - 1:i1])
    lam = m.lambda_; ok = ~np.isnan(lam).any(axis=1) & ~np.isnan(m.uar_ser)
    ns = m.ns; L = m.n_uarlag
    out = np.full((len(ids), Tw, ns), np.nan)
    for k, rep in enumerate(ids):
        rng = np.random.Generator(np.random.Philox(key=[seed, int(rep)]))
        idx = rng.integers(0, len(resid), size=Tw - p)
        fs = np.empty((Tw, r)); fs[:p] = F[:p]
        for t in range(p, Tw):
            z = np.concatenate([[1.0]] + [fs[t - l] for l in range(1, p + 1)])
            fs[t] = z @ beta + resid[idx[t - p]]
        eta = rng.standard_normal((Tw + burn, ns))
        u = np.zeros((Tw + burn, ns))
        ac = np.where(ok[:, None], m.uar_coef, 0.0); ser = np.where(ok, m.uar_ser, 0.0)
        for t in range(Tw + burn):
            acc = ser * eta[t]
            for l in range(1, L + 1):
                if t - l >= 0:
                    acc = acc + ac[:, l - 1] * u[t - l]
            u[t] = acc
        x = fs @ np.where(ok[:, None], lam, 0.0).T + u[burn:]
        x[:, ~ok] = np.nan
        x[mask] = np.nan
        out[k] = x
    return out


def bootstrap_irf(lib, m, n_rep, H=24, rank=0, world=1, seed=SEED):
    """C4: bootstrap distribution of the factor-VAR impulse responses.  Every replication is
    re-estimated with the full non-parametric pipeline (ALS factors -> loadings -> VAR -> IRF), all
    batched on the device; factor signs are aligned with the original estimate.  Returns
    (irfs (n_rep, r, H, r), bands dict of 5/16/50/84/95 percentiles)."""
    b, e = lib.shard_range(n_rep, rank, world)
    r = m.nfac_t; p = m.factor_var_model.nlag
    rec = np.empty((e - b, r * H * r))
    if e > b:
        Xs = bootstrap_panels(m, range(b, e), seed)
        incl = m.inclcode == 1
        als = lib.estimate_factor(Xs[:, :, incl], r, nt_min=m.nt_min_factor_estimation, tol=m.tol, compute_r2=False)
        Fb = als["F"]                                                        # (n, Tw, r)
        F0 = m.factor[m.initperiod - 1:m.lastperiod]
        sg = np.sign(np.einsum("btr,tr->br", Fb, F0)); sg[sg == 0] = 1.0
        Fb = Fb * sg[:, None, :]
        var = lib.estimate_var(Fb, p, True)
        irf = lib.irf(var["M"], var["Q"], var["G"], H, list(range(r)))       # (n, r, H, r)
        rec[:] = irf.reshape(e - b, -1)
    allrec = gather_records(rec, n_rep, rank, world, lib)
    irfs = allrec.reshape(n_rep, r, H, r)
    bands = {q: np.percentile(irfs, q, axis=0) for q in (5, 16, 50, 84, 95)}
    return irfs, bands
