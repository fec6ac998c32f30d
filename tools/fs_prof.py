#!/usr/bin/env python
"""Section timers of k_em_filter_smooth (CTA 0) on a C3-shaped or C1-shaped EM run.  usage: fs_prof.py c3|c1 [iters]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynamic_factor_models_b200 as D
from dynamic_factor_models_b200 import Library, replicate
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = Library()
if which == "c3":
    N, r, T, p = 2000, 20, 2000, 1
    Xs = lib.simulate_panels(0, 1, N, r, T, 20260922)[0]
    F = lib.pca_score(Xs, r)
    Lam, R, A, Q = lib.em_init_from_factors(Xs, F, p)
else:
    p = 4
    z = np.load(os.path.join(ROOT, "tests", "golden", "hom_fac_1_panels.npz"))
    m = D.DFMModel(z["all_bpdata"], z["all_inclcode"], 20, 40, 3, 224, 0, 8, 1e-8, 4, 4)
    D.estimate(m, lib=lib)
    X = replicate.bootstrap_panels(m, range(296), lib=lib)[:, :, m.inclcode == 1]
    Xs, _, _ = lib.standardize(X)
    als = lib.estimate_factor(X, 8, compute_r2=False)
    Lam, R, A, Q = lib.em_init_from_factors(Xs, als["F"], p)
lib.em_kalman(Xs, Lam, R, A, Q, p=p, max_iter=1, want_PF=False, path=1)
lib.fs_prof(1)
lib.em_kalman(Xs, Lam, R, A, Q, p=p, max_iter=iters, want_PF=False, path=1)
v = lib.fs_prof(0)
names = {0: "fwd explicit (rest)", 10: "fwd predict", 16: "fwd copies L,TmT", 17: "fwd chol(L)", 18: "fwd S = I + L'CL", 19: "fwd chol(S) || TmT solve",
         11: "fwd WmT solve", 12: "fwd Pf+means", 13: "fwd ll/stores/freeze test", 22: "fwd run: Phi + u_t", 23: "fwd run: scan", 20: "bwd copy, chol(Pp) || Pf M'",
         21: "bwd solve L", 24: "bwd run: v_t", 25: "bwd run: scan", 26: "bwd run: Gram sums", 27: "scan: powers", 28: "scan: pass 1", 29: "scan: pass 2", 30: "scan: boundaries",
         1: "fwd run: zp + loglik", 2: "fwd serial frozen steps", 14: "bwd loads", 15: "bwd solve L'", 7: "bwd Ps gemms",
         3: "bwd sums/close of step", 4: "bwd run: close (PsF fill)", 6: "transition M-step"}
tot = sum(v[k] for k in names)
print(json.dumps({"config": which, "iters": iters, "explicit_fwd_steps": v[8], "explicit_bwd_steps": v[9],
                  "cycles_per_iter": {names[k]: round(v[k] / iters) for k in names}, "total_cycles_per_iter": round(tot / iters)}, indent=1))
