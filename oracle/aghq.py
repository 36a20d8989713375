"""Outer inference: aghq::marginal_laplace_tmb + aghq::sample_marginal — ORACLE
restatement (test infrastructure).

Call sites in the reference: ``/root/reference/R/02_model_fit.R:284`` and
``:687-689``.  aghq (>= 0.4.1, DESCRIPTION:15), mvQuad, numDeriv and R's
``stats::optim`` are un-vendored dependencies absent from /root/reference;
their published algorithms are restated per SURVEY.md Appendix A.2-A.6.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np
from numpy.polynomial.hermite_e import hermegauss
from scipy.linalg import solve_triangular
from scipy.special import logsumexp

from .laplace import LaplaceObjective


# ----------------------------------------------------------------------------
# stats::optim(method = "BFGS")  ==  vmmin()   (A.2)
# ----------------------------------------------------------------------------
def vmmin(fn, gr, b0, maxit=100, abstol=-np.inf, reltol=np.sqrt(np.finfo(float).eps)):
    stepredn, acctol, reltest = 0.2, 1e-4, 10.0
    b = np.array(b0, dtype=np.float64)
    n = len(b)
    f = fn(b)
    if not np.isfinite(f):
        raise FloatingPointError("initial value in 'vmmin' is not finite")
    Fmin = f
    funcount = gradcount = 1
    g = np.asarray(gr(b), dtype=np.float64).copy()
    it = 1
    ilast = gradcount
    Bm = np.eye(n)
    while True:
        if ilast == gradcount:
            Bm = np.eye(n)
        X = b.copy()
        c = g.copy()
        t = -(Bm @ c)
        gradproj = float(t @ c)
        if gradproj < 0.0:
            steplength = 1.0
            accpoint = False
            while True:
                count = 0
                for i in range(n):
                    b[i] = X[i] + steplength * t[i]
                    if reltest + X[i] == reltest + b[i]:
                        count += 1
                if count < n:
                    f = fn(b)
                    funcount += 1
                    accpoint = bool(np.isfinite(f) and (f <= Fmin + gradproj * steplength * acctol))
                    if not accpoint:
                        steplength *= stepredn
                if count == n or accpoint:
                    break
            enough = (f > abstol) and abs(f - Fmin) > reltol * (abs(Fmin) + reltol)
            if not enough:
                count = n
                Fmin = f
            if count < n:
                Fmin = f
                g = np.asarray(gr(b), dtype=np.float64).copy()
                gradcount += 1
                it += 1
                t = steplength * t
                c = g - c
                D1 = float(t @ c)
                if D1 > 0:
                    Xv = Bm @ c
                    D2 = 1.0 + float(Xv @ c) / D1
                    Bm = Bm + (D2 * np.outer(t, t) - np.outer(Xv, t) - np.outer(t, Xv)) / D1
                else:
                    ilast = gradcount
            else:
                if ilast < gradcount:
                    count = 0
                    ilast = gradcount
        else:
            count = 0
            if ilast == gradcount:
                count = n
            else:
                ilast = gradcount
        if it >= maxit:
            break
        if gradcount - ilast > 2 * n:
            ilast = gradcount
        if count == n and ilast == gradcount:
            break
    return {"par": b, "value": Fmin, "fncount": funcount, "grcount": gradcount,
            "convergence": 0 if it < maxit else 1}


# ----------------------------------------------------------------------------
# mvQuad::createNIGrid(dim=S, type="GHe", level=k) product rule   (A.4)
# ----------------------------------------------------------------------------
def gh_rule(k):
    """Nodes of the probabilists' Hermite polynomial He_k and weights that
    integrate g(z) dz (the N(0,1) GH weight divided by phi(z))."""
    z, w = hermegauss(k)
    order = np.argsort(z)
    z, w = z[order], w[order]
    z = 0.5 * (z - z[::-1])                        # enforce exact symmetry
    w = 0.5 * (w + w[::-1])
    return z, w * np.exp(0.5 * z * z)


def gh_product_grid(S, k):
    """K = k^S nodes (K x S) and weights (K); first coordinate varies fastest."""
    z, w = gh_rule(k)
    idx = np.indices((k,) * S).reshape(S, -1)
    idx = idx[::-1]                                # expand.grid order
    return z[idx].T.copy(), np.prod(w[idx], axis=0)


@dataclass
class AGHQFit:
    """Fields of the ``c("marginallaplace","aghq")`` object that BayesGP reads
    (SURVEY.md section 8b)."""
    k: int
    mode: np.ndarray
    hessian: np.ndarray
    convergence: int
    nodes: np.ndarray              # K x S
    weights: np.ndarray            # K
    logpost: np.ndarray            # K   (= -ff.fn(theta_j))
    lognormconst: float
    logpost_normalized: np.ndarray
    modes: np.ndarray              # K x p  (modesandhessians$mode)
    hessians: np.ndarray           # K x p x p
    marginals: List[dict] = field(default_factory=list)
    opt: dict = field(default_factory=dict)


def rescaled_grid(mode, hess, k, order=None):
    """mvQuad::rescale(grid, m=mode, C=forceSymmetric(solve(H)), dec.type=2):
    theta_j = mode + L z_j, w_j = omega_j det(L), C = L L^T."""
    S = len(mode)
    order = list(range(S)) if order is None else list(order)
    C = np.linalg.inv(hess)
    C = np.triu(C) + np.triu(C, 1).T               # forceSymmetric (upper)
    C = C[np.ix_(order, order)]
    L = np.linalg.cholesky(C)
    z, w = gh_product_grid(S, k)
    nodes_perm = mode[order][None, :] + z @ L.T
    nodes = np.empty_like(nodes_perm)
    nodes[:, order] = nodes_perm
    return nodes, w * np.prod(np.diag(L)), L


def normalize_logpost(ff: LaplaceObjective, mode, hess, k, order=None):
    nodes, weights, L = rescaled_grid(mode, hess, k, order)
    logpost = np.array([-ff.fn(th) for th in nodes])
    lognormconst = float(logsumexp(logpost + np.log(weights)))
    return nodes, weights, logpost, lognormconst, L


def marginal_laplace_tmb(ff: LaplaceObjective, k: int, startingvalue, mode=None, hessian=None) -> AGHQFit:
    """aghq::marginal_laplace_tmb(ff, k, startingvalue) with default_control_tmb()
    (BFGS, numhessian via ff$he, product grid, marginals by "reuse")."""
    S = ff.m.S
    if mode is None:
        opt = vmmin(ff.fn, ff.gr, np.asarray(startingvalue, dtype=np.float64))
        mode = opt["par"].copy()
    else:
        opt = {"par": np.asarray(mode, dtype=np.float64), "convergence": 0}
        mode = opt["par"].copy()
    if hessian is None:
        hessian = ff.he(mode)
    hessian = np.atleast_2d(np.asarray(hessian, dtype=np.float64))
    nodes, weights, logpost, lognormconst, L = normalize_logpost(ff, mode, hessian, k)
    logpost_normalized = logpost - lognormconst
    K = len(weights)
    # marginals ("reuse")  (A.5)
    marginals = []
    z1, w1 = gh_rule(k)
    for j in range(S):
        if j == 0:
            nj, wj, lpn, Lj = nodes, weights, logpost_normalized, L
        else:
            order = [j] + [i for i in range(S) if i != j]
            nj, wj, lpj, lncj, Lj = normalize_logpost(ff, mode, hessian, k, order)
            lpn = lpj - lncj
        th = mode[j] + Lj[0, 0] * z1
        ww = w1 * Lj[0, 0]
        lm = np.empty(k)
        for q in range(k):
            sel = np.arange(K) % k == q             # first coordinate fastest
            lm[q] = logsumexp(lpn[sel] + np.log(wj[sel])) - np.log(ww[q])
        marginals.append({"theta": th, "logmargpost": lm, "w": ww})
    # per-node modes and Hessians
    modes = np.empty((K, ff.m.p))
    hessians = np.empty((K, ff.m.p, ff.m.p))
    for j in range(K):
        ff.fn(nodes[j])
        modes[j] = ff.last_par
        hessians[j] = ff.sp_hess()
    return AGHQFit(k=k, mode=mode, hessian=hessian, convergence=opt.get("convergence", 0), nodes=nodes,
                   weights=weights, logpost=logpost, lognormconst=lognormconst,
                   logpost_normalized=logpost_normalized, modes=modes, hessians=hessians,
                   marginals=marginals, opt=opt)


def theta_moments(fit: AGHQFit):
    """aghq::compute_moment on the normalised posterior: mean and sd per theta."""
    lam = fit.weights * np.exp(fit.logpost_normalized)
    mean = lam @ fit.nodes
    var = lam @ (fit.nodes - mean[None, :]) ** 2
    return mean, np.sqrt(var)


def node_probabilities(fit: AGHQFit):
    """lambda_j = w_j exp(logpost_normalized_j)  (A.6)."""
    return fit.weights * np.exp(fit.logpost_normalized)


def sample_marginal(fit: AGHQFit, Z: np.ndarray, node_idx: np.ndarray):
    """aghq::sample_marginal with the random inputs made explicit (A.6):
    ``W_m = mode_{j(m)} + R_{j(m)}^{-1} z_m`` with ``R_j = chol(H_j)`` upper.
    ``Z`` is p x M standard normal, ``node_idx`` (M) 0-based.  Returns p x M."""
    p, M = Z.shape
    out = np.empty((p, M))
    for j in np.unique(node_idx):
        sel = np.nonzero(node_idx == j)[0]
        H = fit.hessians[j]
        H = np.triu(H) + np.triu(H, 1).T
        R = np.linalg.cholesky(H).T
        out[:, sel] = fit.modes[j][:, None] + solve_triangular(R, Z[:, sel], lower=False)
    return out
