"""40-digit reference values of the Laplace objective L(theta) on the README model (covid_canada, Poisson IWP3
k = 30 + 6 weekday effects), made with mpmath in the BUILD container (committed output: covid_hp.json).

Why: cond(H) = 3.7e11 on this model, so two FP64 evaluations of 1/2 logdet H differ by ~cond * eps ~ 1e-5 from
each other; against an extended-precision value each implementation can be held to the north-star tolerance
(1e-8 relative) on its own.  Follows the same restatement as oracle/model.py / oracle/laplace.py
(/root/reference/src/BayesGP.cpp:133-168,219-249; TMB Laplace) with every operation in mpmath.
"""
import json
import os
import sys

import numpy as np
from mpmath import mp, mpf, matrix, exp, log, loggamma, pi, cholesky, lu_solve

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
mp.dps = 40


def hp_laplace(model, theta, W0, iters=4):
    n, p = model.A.shape
    A = [[mpf(float(v)) for v in row] for row in model.A]
    y = [mpf(float(v)) for v in model.y]
    th = mpf(float(theta[0]))
    et = exp(th)
    d = model.d[0]
    Pd = [mpf(float(v)) for v in model.P[0]]
    qdiag = [et * Pd[i] if i < d else mpf(float(model.qfix[i])) for i in range(p)]
    mu0 = [mpf(float(v)) for v in model.mu0]
    W = [mpf(float(v)) for v in W0]
    ll_const = -sum(loggamma(v + 1) for v in y)
    for it in range(iters + 1):
        eta = [sum(A[i][c] * W[c] for c in range(p)) for i in range(n)]
        mu = [exp(e) for e in eta]
        g = [-sum(A[i][c] * (y[i] - mu[i]) for i in range(n)) + qdiag[c] * (W[c] - mu0[c]) for c in range(p)]
        H = matrix(p, p)
        for a in range(p):
            for b in range(a + 1):
                s = sum(A[i][a] * mu[i] * A[i][b] for i in range(n) if A[i][a] != 0 and A[i][b] != 0)
                H[a, b] = s
                H[b, a] = s
            H[a, a] += qdiag[a]
        if it == iters:
            break
        step = lu_solve(H, matrix([-v for v in g]))
        W = [W[c] + step[c] for c in range(p)]
    gmax = max(abs(v) for v in g)
    ll = ll_const + sum(y[i] * eta[i] - mu[i] for i in range(n))
    lpW = -sum(qdiag[c] * (W[c] - mu0[c]) ** 2 for c in range(p)) / 2 + (d * th + mpf(float(model.logPdet[0]))) / 2
    phi = -log(mpf(float(model.alpha[0]))) / mpf(float(model.u[0]))
    lpT = log(phi / 2) - phi * exp(-th / 2) - th / 2
    f = -(ll + lpW + lpT)
    L = cholesky(H)
    logdet = 2 * sum(log(L[i, i]) for i in range(p))
    value = f + logdet / 2 - mpf(p) / 2 * log(2 * pi)
    return value, W, gmax


def main():
    from helpers import covid_model
    from oracle.laplace import LaplaceObjective
    model = covid_model()[0]
    off = LaplaceObjective(model)
    # the theta values the GPU tests use: tests/test_gpu_core.py CASES and the 4 GH nodes of the oracle fit
    og = np.load(os.path.join(HERE, "oracle_covid.npz"))
    thetas = [0.0, -3.2, -2.5] + [float(v) for v in np.ravel(og["nodes"])]
    out = {"source": "tests/golden/make_hp_covid.py (mpmath, 40 digits)", "theta": [], "value": [], "value_str": [],
           "gmax": [], "oracle_fp64_value": [], "mode": []}
    for t in thetas:
        v64 = off.fn(np.array([t]))
        val, W, gmax = hp_laplace(model, np.array([t]), off.last_par)
        out["theta"].append(t)
        out["value"].append(float(val))
        out["value_str"].append(mp.nstr(val, 25))
        out["gmax"].append(float(gmax))
        out["oracle_fp64_value"].append(float(v64))
        out["mode"].append([float(x) for x in W])
        print(t, mp.nstr(val, 20), "fp64 oracle diff", float(v64 - val), "gmax", float(gmax), flush=True)
    with open(os.path.join(HERE, "covid_hp.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
