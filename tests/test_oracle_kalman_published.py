"""External pin of the Kalman-filter oracle: the local-level model on the Nile data, the worked
example of Durbin & Koopman (2012, "Time Series Analysis by State Space Methods", 2nd ed., ch. 2).

    x_t = mu_t + eps_t,  mu_{t+1} = mu_t + eta_t       ==  N = 1, r = 1, Lam = 1, A = 1, R = s2_eps, Q = s2_eta

Published: ML estimates s2_eps = 15099, s2_eta = 1469.1 (also the printed output of R's StructTS(Nile, "level")
example), maximised diffuse log-likelihood -632.54.  The oracle's filter starts from z_1 ~ N(0, P0); with the usual
big-P0 approximation of the diffuse prior (P0 = 1e7, as in D&K sec. 2.9) its log-likelihood equals the diffuse one plus
the (then parameter-free to 1e-3) contribution of the first observation, which is removed in closed form below.
The reference repo has no Kalman code at all, so this is the pin that is independent of this repo.
"""
import json
import os

import numpy as np
import pytest
from scipy.optimize import minimize

from oracle import kalman_em as K

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "nile_local_level.json")))
NILE = np.array(G["nile"], float)
PUB = G["published"]
P1 = 1e7


def first_term(s2e):
    return -0.5 * np.log(2 * np.pi) - 0.5 * np.log(P1 + s2e) - 0.5 * NILE[0] ** 2 / (P1 + s2e)


def estep(s2e, s2n):
    return K.e_step(NILE[:, None], np.array([[1.0]]), np.array([s2e]), np.array([[1.0]]), np.array([[s2n]]), np.array([[P1]]), 1)


def test_data_is_the_published_series():
    assert len(NILE) == 100 and abs(NILE.mean() - PUB["mean"]) < 1e-9 and NILE[0] == 1120 and NILE[-1] == 740


def test_loglik_at_published_mle():
    es = estep(PUB["sigma2_eps"], PUB["sigma2_eta"])
    assert abs((es["loglik"] - first_term(PUB["sigma2_eps"])) - PUB["loglik_diffuse"]) < 0.01


def test_published_estimates_maximise_the_oracle_likelihood():
    f = lambda th: -estep(np.exp(th[0]), np.exp(th[1]))["loglik"]
    r = minimize(f, [np.log(10000.0), np.log(3000.0)], method="Nelder-Mead", options=dict(xatol=1e-7, fatol=1e-10))
    s2e, s2n = np.exp(r.x)
    assert abs(s2e / PUB["sigma2_eps"] - 1) < 1e-3 and abs(s2n / PUB["sigma2_eta"] - 1) < 1e-3


def test_steady_state_matches_the_riccati_solution():
    """D&K sec. 2.11: P = P s2_eps / (P + s2_eps) + s2_eta  ->  P = (q + sqrt(q^2 + 4 q)) / 2 * s2_eps, q = s2_eta / s2_eps."""
    s2e, s2n = PUB["sigma2_eps"], PUB["sigma2_eta"]
    q = s2n / s2e
    Pbar = 0.5 * (q + np.sqrt(q * q + 4 * q)) * s2e
    es = estep(s2e, s2n)
    assert abs(es["Pp"][-1, 0, 0] / Pbar - 1) < 1e-10
    assert abs(es["Pf"][-1, 0, 0] / (Pbar * s2e / (Pbar + s2e)) - 1) < 1e-10
    # RTS smoother: the smoothed variance in the middle of the sample solves V = Pf + J^2 (V - Pbar), J = Pf / Pbar
    Pf = Pbar * s2e / (Pbar + s2e); J = Pf / Pbar
    V = (Pf - J * J * Pbar) / (1 - J * J)
    assert abs(es["Ps"][50, 0, 0] / V - 1) < 1e-8


def test_c_port_agrees_on_the_published_example():
    from oracle.c import kem
    out = kem.em_kalman_batch(NILE[None, :, None], np.ones((1, 1, 1)), np.full((1, 1), PUB["sigma2_eps"]), np.ones((1, 1, 1)),
                              np.full((1, 1, 1), PUB["sigma2_eta"]), p=1, P0=np.full((1, 1, 1), P1), max_iter=1)
    assert abs((out["loglik"][0, 0] - first_term(PUB["sigma2_eps"])) - PUB["loglik_diffuse"]) < 0.01
    es = estep(PUB["sigma2_eps"], PUB["sigma2_eta"])
    np.testing.assert_allclose(out["F"][0, :, 0], es["zs"][:, 0], rtol=1e-10)
