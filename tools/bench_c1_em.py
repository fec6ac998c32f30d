#!/usr/bin/env python
"""Timing of the state-space EM on C1/C4-shaped panels (T=222, N=139, r=8, VAR(4) state k=32, 5.7 % missing):
B bootstrap draws of the hom_fac_1 model, `iters` EM iterations, device-resident.  Prints panel-EM-iterations/s."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dynamic_factor_models_b200 as D
from dynamic_factor_models_b200 import Library, replicate

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
p = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lib = Library(path=os.environ.get("DFM_BENCH_LIB"))
z = np.load(os.path.join(ROOT, "tests", "golden", "hom_fac_1_panels.npz"))
m = D.DFMModel(z["all_bpdata"], z["all_inclcode"], 20, 40, 3, 224, 0, 8, 1e-8, 4, 4)
D.estimate(m, lib=lib)
X = replicate.bootstrap_panels(m, range(B), lib=lib)[:, :, m.inclcode == 1]
Xs, _, _ = lib.standardize(X)
als = lib.estimate_factor(X, 8, compute_r2=False)
Lam, R, A, Q = lib.em_init_from_factors(Xs, als["F"], p)
lib.em_kalman(Xs[:8], Lam[:8], R[:8], A[:8], Q[:8], p=p, max_iter=2, want_PF=False)
lib.profile(True)
t0 = time.perf_counter()
out = lib.em_kalman(Xs, Lam, R, A, Q, p=p, max_iter=iters, want_PF=False)
dt = time.perf_counter() - t0
prof = lib.profile_report()
print(json.dumps({"B": B, "iters": iters, "p": p, "seconds_incl_copies": dt, "panel_iters_per_s": B * iters / dt,
                  "status_ok": bool((np.asarray(out["status"]) == 0).all()),
                  "kernel_ms": {k: round(v[0], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6]}}))
