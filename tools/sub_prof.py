#!/usr/bin/env python
"""Section timers of k_subspace_eig2 (CTA 0).  usage: sub_prof.py c5|c4"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynamic_factor_models_b200 as D
from dynamic_factor_models_b200 import Library, replicate
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
lib = Library()
if which == "c5":
    X = lib.simulate_panels(0, 296, 200, 8, 500, 20260922)
else:
    z = np.load(os.path.join(ROOT, "tests", "golden", "hom_fac_1_panels.npz"))
    m = D.DFMModel(z["all_bpdata"], z["all_inclcode"], 20, 40, 3, 224, 0, 8, 1e-8, 4, 4)
    D.estimate(m, lib=lib)
    X = replicate.bootstrap_panels(m, range(296), lib=lib)[:, :, m.inclcode == 1]
lib.estimate_factor(X, 8, compute_r2=False, max_iter=1)
lib.fs_prof(1)
lib.estimate_factor(X, 8, compute_r2=False, max_iter=1)
v = lib.fs_prof(0)[48:]
names = ["S = V'V gemm + sym", "chol", "trsm", "first product", "H gemm", "jacobi", "rotate + product + residual", "power products (x2) + normalise"]
print(json.dumps({"config": which, "cycles": {n: v[i] for i, n in enumerate(names)}, "cycles_total": sum(v[:8]), "subspace cycles": v[8], "calls": v[9], "jacobi sweeps": v[10]}, indent=1))
