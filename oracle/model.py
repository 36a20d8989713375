"""Latent-Gaussian model container + TMB objective — ORACLE restatement.

Restates ``/root/reference/src/BayesGP.cpp:30-253`` (the objective template)
on the data layout built by ``/root/reference/R/02_model_fit.R:30-183,249-252``
(``tmbdat`` / ``tmbparams``).  Derivatives w.r.t. W that TMB obtains by AD are
written in closed form.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
from scipy.special import gammaln

FAMILY_GAUSSIAN, FAMILY_POISSON, FAMILY_BINOMIAL, FAMILY_NONE = 0, 1, 2, -2
FAMILY_CODES = {"Gaussian": 0, "Poisson": 1, "Binomial": 2, "none": -2}   # R/02_model_fit.R:9-28


@dataclass
class Model:
    """Mirror of ``tmbdat`` (R/02_model_fit.R:152-183).

    ``B[j]`` n x d_j, ``P[j]`` either a length-d_j vector (diagonal P, the IWP
    case R/01_utility.R:325-344) or a dense d_j x d_j matrix (sGP), ``X[j]``
    n x b_j boundary columns, ``Xf[i]`` n x c_i fixed-effect columns.
    ``u`` / ``alpha`` have one entry per theta (the Gaussian-noise theta last,
    R/02_model_fit.R:120-121).
    """
    family: int
    y: np.ndarray
    B: List[np.ndarray]
    P: List[np.ndarray]
    logPdet: List[float]
    u: List[float]
    alpha: List[float]
    X: List[np.ndarray]
    betaprec: List[float]
    betamean: List[float]
    Xf: List[np.ndarray]
    beta_fixed_prec: List[float]
    beta_fixed_mean: List[float]
    size: Optional[np.ndarray] = None
    A: np.ndarray = field(init=False, repr=False)

    def __post_init__(self):
        self.y = np.asarray(self.y, dtype=np.float64)
        n = self.n
        if self.family == FAMILY_BINOMIAL and self.size is None:
            self.size = np.ones(n)                       # R/02_model_fit.R:176-183
        blocks = list(self.B) + list(self.X) + list(self.Xf)
        # W = c(U_1..U_J, beta_1..beta_J, beta_fixed_0..)   src/BayesGP.cpp:76-127
        self.A = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float64).reshape(n, -1) for b in blocks], axis=1))
        self.d = [np.asarray(b).reshape(n, -1).shape[1] for b in self.B]
        self.betadim = [np.asarray(b).reshape(n, -1).shape[1] for b in self.X]
        self.fixdim = [np.asarray(b).reshape(n, -1).shape[1] for b in self.Xf]
        self.J = len(self.B)
        self.S = self.J + (1 if self.family == FAMILY_GAUSSIAN else 0)
        assert len(self.u) == self.S and len(self.alpha) == self.S
        self.p = self.A.shape[1]
        # prior mean mu0 and the theta-independent diagonal part of Q
        mu0 = np.zeros(self.p)
        qfix = np.zeros(self.p)
        o = sum(self.d)
        for bd, pr, mn in zip(self.betadim, self.betaprec, self.betamean):
            mu0[o:o + bd] = mn
            qfix[o:o + bd] = pr
            o += bd
        for fd, pr, mn in zip(self.fixdim, self.beta_fixed_prec, self.beta_fixed_mean):
            mu0[o:o + fd] = mn
            qfix[o:o + fd] = pr
            o += fd
        self.mu0, self.qfix = mu0, qfix
        self.u_off = np.concatenate([[0], np.cumsum(self.d)]).astype(int)
        if self.family == FAMILY_POISSON:
            self.ll_const = -np.sum(gammaln(self.y + 1.0))
        elif self.family == FAMILY_BINOMIAL:
            s, y = self.size, self.y
            lch = gammaln(s + 1.0) - gammaln(y + 1.0) - gammaln(s - y + 1.0)
            self.ll_const = float(np.sum(np.where(s > 1, lch, 0.0)))   # TMB dbinom_robust adds it only if size > 1
        elif self.family == FAMILY_GAUSSIAN:
            self.ll_const = -0.5 * n * np.log(2.0 * np.pi)
        else:
            self.ll_const = 0.0

    @property
    def n(self):
        return len(self.y)

    # ---- Q(theta) = blockdiag(e^{theta_j} P_j, betaprec, fixedprec)  (BayesGP.cpp:219-238)
    def Q(self, theta):
        Q = np.diag(self.qfix.copy())
        for j in range(self.J):
            a, b = self.u_off[j], self.u_off[j + 1]
            Pj = self.P[j]
            Q[a:b, a:b] = np.exp(theta[j]) * (np.diag(Pj) if Pj.ndim == 1 else Pj)
        return Q

    def Qmul(self, theta, v):
        out = self.qfix * v
        for j in range(self.J):
            a, b = self.u_off[j], self.u_off[j + 1]
            Pj = self.P[j]
            out[a:b] = np.exp(theta[j]) * (Pj * v[a:b] if Pj.ndim == 1 else Pj @ v[a:b])
        return out

    # ---- per-observation likelihood pieces (BayesGP.cpp:155-168,212-214)
    def lik(self, eta, theta):
        """Return (ll, r, w, c3): log-lik, d ll/d eta, -d2 ll/d eta2, d w/d eta."""
        y = self.y
        if self.family == FAMILY_GAUSSIAN:
            tau = np.exp(theta[self.S - 1])
            res = y - eta
            ll = self.ll_const + 0.5 * self.n * theta[self.S - 1] - 0.5 * tau * np.sum(res * res)
            return ll, tau * res, np.full(self.n, tau), np.zeros(self.n)
        if self.family == FAMILY_POISSON:
            mu = np.exp(eta)
            ll = self.ll_const + np.sum(y * eta - mu)
            return ll, y - mu, mu, mu
        if self.family == FAMILY_BINOMIAL:
            s = self.size
            # dbinom_robust: y*log p + (size-y)*log(1-p), p = logistic(eta)
            l1pe_neg = np.logaddexp(0.0, -eta)
            l1pe_pos = np.logaddexp(0.0, eta)
            ll = self.ll_const + np.sum(-y * l1pe_neg - (s - y) * l1pe_pos)
            pi = 1.0 / (1.0 + np.exp(-eta))
            v = s * pi * (1.0 - pi)
            return ll, y - s * pi, v, v * (1.0 - 2.0 * pi)
        z = np.zeros(self.n)
        return 0.0, z, z.copy(), z.copy()

    def log_prior_theta(self, theta):                     # BayesGP.cpp:241-246
        lp = 0.0
        for i in range(self.S):
            phi = -np.log(self.alpha[i]) / self.u[i]
            lp += np.log(0.5 * phi) - phi * np.exp(-0.5 * theta[i]) - 0.5 * theta[i]
        return lp

    def objective(self, W, theta, want="fgH"):
        """f(W, theta) = -(ll + lpW + lpT)  (BayesGP.cpp:249) and derivatives in W."""
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        eta = self.A @ W
        ll, r, w, c3 = self.lik(eta, theta)
        dW = W - self.mu0
        QdW = self.Qmul(theta, dW)
        lpW = -0.5 * float(dW @ QdW)
        for j in range(self.J):
            lpW += 0.5 * (self.d[j] * theta[j] + self.logPdet[j])
        f = -(ll + lpW + self.log_prior_theta(theta))
        out = {"f": f, "eta": eta, "r": r, "w": w, "c3": c3}
        if "g" in want:
            out["g"] = -(self.A.T @ r) + QdW
        if "H" in want:
            out["H"] = self.A.T @ (w[:, None] * self.A) + self.Q(theta)
        return out
