"""TMB-style Laplace objective ``ff`` — ORACLE restatement (test infrastructure).

Restates what ``TMB::MakeADFun(data, parameters, random="W")`` provides at the
call site ``/root/reference/R/02_model_fit.R:276-283`` (TMB itself is an
un-vendored dependency; algorithm per SURVEY.md Appendix A.1):

  ff.fn(theta)  = f(w_hat, theta) + 1/2 logdet H(w_hat, theta) - p/2 log(2 pi)
  ff.gr(theta)  = exact gradient of the above (A.1.3)
  ff.he(theta)  = numDeriv::jacobian(ff.gr, theta)          (R/02_model_fit.R:283)
  ff.last_par   = w_hat of the most recent fn/gr call       (ff$env$last.par[random])
  ff.spHess()   = H(w_hat, theta)                           (ff$env$spHess(random=TRUE))

The inner problem is strictly convex in W for families 0/1/2 with PD Q, so the
mode is unique; the inner solver is a damped Newton (step halving) started from
the previous mode (TMB warm start), converged when max|g| < 1e-8 (TMB
``newton`` default ``grad.tol``).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular

from .model import FAMILY_GAUSSIAN, Model

LOG_2PI = float(np.log(2.0 * np.pi))


class LaplaceObjective:
    def __init__(self, model: Model, grad_tol: float = 1e-8, step_tol: float = 1e-8, maxit: int = 100):
        self.m = model
        self.grad_tol = grad_tol
        self.step_tol = step_tol
        self.maxit = maxit
        self.par = np.zeros(model.S)                 # tmbparams theta = 0 (R/02_model_fit.R:249-252)
        self.last_par = np.zeros(model.p)            # W = 0
        self.last_theta = None
        self.last_H = None
        self.last_L = None                           # lower Cholesky factor of H
        self.n_fn = 0
        self.n_gr = 0
        self.newton_iters = 0
        self.converged = True

    # -- inner Newton -------------------------------------------------------
    def inner(self, theta, w0=None):
        m = self.m
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        w = (self.last_par if w0 is None else w0).copy()
        o = m.objective(w, theta, "fgH")
        if not np.isfinite(o["f"]):                  # bad warm start for this theta: restart from 0
            w = np.zeros(m.p)
            o = m.objective(w, theta, "fgH")
        self.converged = False
        if not (np.isfinite(o["f"]) and np.all(np.isfinite(o["g"]))):
            return w, o                              # TMB: NaN value, the outer optimiser backtracks
        for it in range(self.maxit):
            gmax = np.max(np.abs(o["g"]))
            if gmax < self.grad_tol:
                self.converged = True
                break
            try:
                c = cho_factor(o["H"], lower=True)
            except (np.linalg.LinAlgError, ValueError):
                break
            step = -cho_solve(c, o["g"])
            # TMB newton() also stops on the step size (step.tol = tol): with
            # cond(H) ~ 1e10 the FP64 gradient noise floor can sit above grad.tol.
            if np.max(np.abs(step)) < self.step_tol:
                self.converged = True
                break
            t = 1.0
            accepted = False
            for _ in range(40):
                o2 = m.objective(w + t * step, theta, "fgH")
                if np.isfinite(o2["f"]) and np.all(np.isfinite(o2["g"])) and (
                        o2["f"] <= o["f"] or np.max(np.abs(o2["g"])) < gmax):
                    accepted = True
                    break
                t *= 0.5
            if not accepted:
                break
            w = w + t * step
            o = o2
            self.newton_iters += 1
        return w, o

    def _eval(self, theta):
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        if self.last_theta is not None and np.array_equal(theta, self.last_theta) and self.last_L is not None:
            return self._cache
        w, o = self.inner(theta)
        if not self.converged:
            self._cache = (np.nan, w, o, None)
            self.last_theta, self.last_L = None, None
            return self._cache
        L = np.linalg.cholesky(o["H"])
        logdet = 2.0 * np.sum(np.log(np.diag(L)))
        val = o["f"] + 0.5 * logdet - 0.5 * self.m.p * LOG_2PI
        self.last_par, self.last_theta, self.last_H, self.last_L = w, theta.copy(), o["H"], L
        self._cache = (val, w, o, L)
        return self._cache

    def fn(self, theta):
        self.n_fn += 1
        return float(self._eval(theta)[0])

    def sp_hess(self):
        return self.last_H

    # -- exact gradient of the Laplace objective (A.1.3) --------------------
    def gr(self, theta):
        self.n_gr += 1
        m = self.m
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        val, w, o, L = self._eval(theta)
        if L is None:
            return np.full(m.S, np.nan)
        Hi = cho_solve((L, True), np.eye(m.p))
        grad = np.zeros(m.S)
        # leverage term v = A^T (c3 * q), q_i = a_i^T H^-1 a_i (zero for Gaussian)
        if np.any(o["c3"] != 0.0):
            Y = solve_triangular(L, m.A.T, lower=True)        # p x n, columns L^-1 a_i
            q = np.sum(Y * Y, axis=0)
            v = m.A.T @ (o["c3"] * q)
            Hiv = Hi @ v
        else:
            Hiv = np.zeros(m.p)
        for k in range(m.J):
            a, b = m.u_off[k], m.u_off[k + 1]
            Pk = m.P[k]
            U = w[a:b]
            PU = Pk * U if Pk.ndim == 1 else Pk @ U
            ek = np.exp(theta[k])
            phi = -np.log(m.alpha[k]) / m.u[k]
            dfdth = 0.5 * ek * float(U @ PU) - 0.5 * m.d[k] - 0.5 * phi * np.exp(-0.5 * theta[k]) + 0.5
            Hib = Hi[a:b, a:b]
            tr = float(np.sum(np.diag(Hib) * Pk)) if Pk.ndim == 1 else float(np.sum(Hib * Pk))
            implicit = -0.5 * ek * float(Hiv[a:b] @ PU)
            grad[k] = dfdth + 0.5 * ek * tr + implicit
        if m.family == FAMILY_GAUSSIAN:
            k = m.S - 1
            tau = np.exp(theta[k])
            res = m.y - o["eta"]
            phi = -np.log(m.alpha[k]) / m.u[k]
            dfdth = -0.5 * m.n + 0.5 * tau * float(res @ res) - 0.5 * phi * np.exp(-0.5 * theta[k]) + 0.5
            # 1/2 tr(Hi * tau A^T A) = 1/2 (p - tr(Hi Q))
            trHiQ = float(np.sum(Hi * m.Q(theta)))
            grad[k] = dfdth + 0.5 * (m.p - trHiQ)            # c3 = 0: no implicit term
        return grad

    # -- numDeriv::jacobian(ff$gr, theta), method "Richardson" (A.3) --------
    def he(self, theta, d=1e-4, eps=1e-4, r=4, v=2):
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        return richardson_jacobian(self.gr, theta, d=d, eps=eps, r=r, v=v)


def richardson_jacobian(func, x, d=1e-4, eps=1e-4, r=4, v=2):
    """numDeriv::jacobian(func, x, method="Richardson") with its defaults
    ``d=1e-4, eps=1e-4, r=4, v=2, zero.tol=sqrt(.Machine$double.eps/7e-7)``."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    zero_tol = np.sqrt(np.finfo(float).eps / 7e-7)
    f0 = np.asarray(func(x))
    h = np.abs(d * x) + eps * (np.abs(x) < zero_tol)
    a = np.zeros((len(f0), r, n))
    for k in range(r):
        for i in range(n):
            e = np.zeros(n)
            e[i] = h[i]
            a[:, k, i] = (np.asarray(func(x + e)) - np.asarray(func(x - e))) / (2.0 * h[i])
        h = h / v
    for mm in range(1, r):
        a = (a[:, 1:r - mm + 1, :] * (4.0 ** mm) - a[:, 0:r - mm, :]) / (4.0 ** mm - 1.0)
    return a[:, 0, :]
