"""Posterior summaries of a fitted model — the host-side functions the R layer runs on the object
``model_fit`` returns (``/root/reference/R/03_post_fit.R``):

* ``summary(fit)``            summary.FitResult (:2-42): the theta table of ``summary(mod)`` (mean, sd, 2.5 %, median,
                              97.5 %) and the fixed-effect table (1st Qu., Median, Mean, 3rd Qu., sd);
* ``var_density(fit, ...)``   var_density (:309-447, aghq branch): posterior and prior density of a smoothing /
                              family standard deviation, optionally on the predictive-SD (PSD) scale.

The theta marginals come out of the CUDA fit (``bgp_fit_get_marginal``); what is done to them here is aghq's
``compute_pdf_and_cdf`` / ``compute_quantiles`` (k-point tables, microseconds of host work): natural cubic spline
of the log marginal with linear continuation outside the nodes, 1000-point grid over the node range widened by half
its length on each side, ``cdf = cumsum(pdf * c(0, diff(theta)))``.  With the README's grid this reproduces the
quantiles printed at ``/root/reference/README.md:83-85`` (tests/test_gpu_summary.py).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np


def _natural_spline(x, y):
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    order = np.argsort(x)
    x, y = x[order], y[order]
    k = len(x)
    if k <= 2:
        raise ValueError("The number of quadrature points is too small, please use aghq_k >= 3.")
    h = np.diff(x)
    slope = np.diff(y) / h
    # tridiagonal system for the second derivatives, zero at both ends (Thomas algorithm)
    diag = 2.0 * (h[:-1] + h[1:])
    rhs = 6.0 * np.diff(slope)
    lower, upper = h[1:-1].copy(), h[1:-1].copy()
    for i in range(1, k - 2):
        w = lower[i - 1] / diag[i - 1]
        diag[i] -= w * upper[i - 1]
        rhs[i] -= w * rhs[i - 1]
    m = np.zeros(k)
    for i in range(k - 3, -1, -1):
        m[i + 1] = (rhs[i] - (upper[i] * m[i + 2] if i < k - 3 else 0.0)) / diag[i]
    d_lo = slope[0] - h[0] * (2.0 * m[0] + m[1]) / 6.0
    d_hi = slope[-1] + h[-1] * (m[-2] + 2.0 * m[-1]) / 6.0

    def f(xn):
        xn = np.asarray(xn, dtype=np.float64)
        i = np.clip(np.searchsorted(x, xn, side="right") - 1, 0, k - 2)
        a, b = xn - x[i], x[i + 1] - xn
        val = ((m[i] * b ** 3 + m[i + 1] * a ** 3) / (6.0 * h[i]) + (y[i] / h[i] - m[i] * h[i] / 6.0) * b
               + (y[i + 1] / h[i] - m[i + 1] * h[i] / 6.0) * a)
        val = np.where(xn < x[0], y[0] + d_lo * (xn - x[0]), val)
        return np.where(xn > x[-1], y[-1] + d_hi * (xn - x[-1]), val)

    return f


def fmm_spline(x, y, n):
    """stats::spline(x, y, n = n, method = "fmm"): the Forsythe-Malcolm-Moler cubic spline (end conditions from the
    cubics through the first and the last four points) evaluated at seq(min(x), max(x), length.out = n)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    m = len(x)
    xo = x[0] + (x[-1] - x[0]) * np.arange(n) / (n - 1.0) if n > 1 else np.array([x[0]])
    if m < 2:
        raise ValueError("spline needs at least two points")
    b, c, d = np.zeros(m), np.zeros(m), np.zeros(m)
    if m < 3:
        b[:] = (y[1] - y[0]) / (x[1] - x[0])
    else:
        d[0] = x[1] - x[0]
        c[1] = (y[1] - y[0]) / d[0]
        for i in range(1, m - 1):
            d[i] = x[i + 1] - x[i]
            b[i] = 2.0 * (d[i - 1] + d[i])
            c[i + 1] = (y[i + 1] - y[i]) / d[i]
            c[i] = c[i + 1] - c[i]
        b[0], b[m - 1] = -d[0], -d[m - 2]
        c[0] = c[m - 1] = 0.0
        if m > 3:
            c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0])
            c[m - 1] = c[m - 2] / (x[m - 1] - x[m - 3]) - c[m - 3] / (x[m - 2] - x[m - 4])
            c[0] = c[0] * d[0] * d[0] / (x[3] - x[0])
            c[m - 1] = -c[m - 1] * d[m - 2] * d[m - 2] / (x[m - 1] - x[m - 4])
        for i in range(1, m):                       # Gaussian elimination
            t = d[i - 1] / b[i - 1]
            b[i] -= t * d[i - 1]
            c[i] -= t * c[i - 1]
        c[m - 1] /= b[m - 1]                        # back substitution
        for i in range(m - 2, -1, -1):
            c[i] = (c[i] - d[i] * c[i + 1]) / b[i]
        b[m - 1] = (y[m - 1] - y[m - 2]) / d[m - 2] + d[m - 2] * (c[m - 2] + 2.0 * c[m - 1])
        for i in range(m - 1):
            b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i])
            d[i] = (c[i + 1] - c[i]) / d[i]
            c[i] = 3.0 * c[i]
        c[m - 1] = 3.0 * c[m - 1]
        d[m - 1] = d[m - 2]
    i = np.clip(np.searchsorted(x, xo, side="right") - 1, 0, m - 1)
    dx = xo - x[i]
    return xo, y[i] + dx * (b[i] + dx * (c[i] + dx * d[i]))


def integrate_xy(x, fx):
    """sfsmisc::integrate.xy(x, fx) over the whole range with its defaults (use.spline = TRUE): the fmm spline on
    max(1024, 3 n) points, then the trapezoid rule (/root/reference/R/02_model_fit.R:774)."""
    x, fx = np.asarray(x, dtype=np.float64), np.asarray(fx, dtype=np.float64)
    order = np.argsort(x, kind="stable")
    x, fx = x[order], fx[order]
    keep = np.concatenate([[True], np.diff(x) != 0])
    x, fx = x[keep], fx[keep]
    xs, ys = fmm_spline(x, fx, max(1024, 3 * len(x)))
    if xs[-1] < x[-1]:
        xs, ys = np.append(xs, x[-1]), np.append(ys, fx[-1])
    return float(np.sum(np.diff(xs) * (ys[1:] + ys[:-1]) / 2.0))


def compute_pdf_and_cdf(marginal, transformation: Optional[str] = None, ngrid: int = 1000):
    """aghq::compute_pdf_and_cdf(marginal, interpolation = 'spline').  ``transformation='sd'`` adds the columns for
    sigma = exp(-theta / 2) the way var_density asks for them (totheta = -2 log x)."""
    th = np.asarray(marginal["theta"], dtype=np.float64)
    lo, hi = float(th.min()), float(th.max())
    half = 0.5 * (hi - lo)
    grid = np.linspace(lo - half, hi + half, ngrid)
    pdf = np.exp(_natural_spline(th, marginal["logmargpost"])(grid))
    out = {"theta": grid, "pdf": pdf, "cdf": np.cumsum(pdf * np.concatenate([[0.0], np.diff(grid)]))}
    if transformation == "sd":
        sigma = np.exp(-grid / 2.0)
        out["transparam"] = sigma
        out["pdf_transparam"] = pdf * (2.0 / sigma)
    elif transformation is not None:
        raise ValueError("unknown transformation %r" % (transformation,))
    return out


def compute_quantiles(marginal, q=(0.025, 0.975)):
    """aghq::compute_quantiles: the last grid point whose cdf is still below q."""
    pc = compute_pdf_and_cdf(marginal)
    vals = []
    for qq in np.atleast_1d(q):
        below = np.flatnonzero(pc["cdf"] < qq)
        vals.append(pc["theta"][below[-1]] if below.size else np.nan)
    return np.array(vals)


def _theta_names(fit):
    names = ["theta(%s)" % t.name for t in fit.instances]
    S = fit.mod.normalized_posterior["nodesandweights"]["theta"].shape[1]
    return names + ["theta(family)"] * (S - len(names))


def summary(fit, echo: bool = True):
    """summary.FitResult: returns {"theta": table, "fixed": table}; prints them like the R method when ``echo``."""
    mod = fit.mod
    mean, sd = mod.theta_moments()
    theta_rows = {}
    for j, name in enumerate(_theta_names(fit)):
        lo, med, hi = compute_quantiles(mod.marginals[j], (0.025, 0.5, 0.975))
        theta_rows[name + ("" if name not in theta_rows else "#%d" % j)] = {
            "mean": float(mean[j]), "sd": float(sd[j]), "2.5%": float(lo), "median": float(med), "97.5%": float(hi)}
    fixed_rows = {}
    if fit.samps is not None and len(fit.fixed_samp_indexes) >= 1:
        samps = fit.samps["samps"]
        for name, idx in fit.fixed_samp_indexes.items():
            r = np.asarray(samps[idx, :], dtype=np.float64).reshape(-1)
            q1, med, q3 = np.quantile(r, [0.25, 0.5, 0.75], method="linear")      # summary.default: type 7
            fixed_rows[name] = {"1st Qu.": float(q1), "Median": float(med), "Mean": float(r.mean()),
                                "3rd Qu.": float(q3), "sd": float(r.std(ddof=1))}
    if echo:
        nw = mod.normalized_posterior["nodesandweights"]
        S = nw["theta"].shape[1]
        print("AGHQ on a %d dimensional posterior with  %s quadrature points\n" % (S, " ".join([str(mod.k)] * S)))
        print("The posterior mode is:", " ".join("%.7g" % v for v in mod.optresults["mode"]), "\n")
        print("The log of the normalizing constant/marginal likelihood is: %.7g \n" % mod.lognormconst)
        print("The covariance matrix used for the quadrature is...")
        print(np.linalg.inv(np.atleast_2d(mod.optresults["hessian"])))
        print("\nHere are some moments and quantiles for the log precision: ")
        for name, r in theta_rows.items():
            print("%-14s" % name, " ".join("%s %.7g" % (k2, v) for k2, v in r.items()))
        if fixed_rows:
            print("\nHere are some moments and quantiles for the fixed effects: \n")
            for name, r in fixed_rows.items():
                print("%-10s" % name, " ".join("%s %.8f" % (k2, v) for k2, v in r.items()))
    return {"theta": theta_rows, "fixed": fixed_rows}


def _theta_logprior(theta, alpha, u):
    """The Exponential prior on the SD, written in theta (src/BayesGP.cpp:241-246; R/03_post_fit.R:312-316)."""
    lam = -math.log(alpha) / u
    return math.log(lam / 2.0) - lam * np.exp(-theta / 2.0) - theta / 2.0


def compute_d_step_sGPsd(d, a):
    """R/01_utility.R:460-462."""
    return math.sqrt((1.0 / a ** 2) * (d / 2.0 - math.sin(2.0 * a * d) / (4.0 * a)))


def var_density(fit, component: Optional[str] = None, h: Optional[float] = None, family_prior=None):
    """var_density(object, component, h) for a model fitted with method = "aghq".

    Returns columns SD, post, prior (+ PSD, post.PSD, prior.PSD when ``h`` is given), ordered by SD."""
    mod = fit.mod
    if component is None:
        if fit.family != "Gaussian":
            raise ValueError("There is no family SD in the fitted model. Please indicate which component of the "
                             "var-parameter that you want to show in `component`.")
        marg = mod.marginals[len(fit.instances)]
        prior = family_prior or getattr(fit, "control_family", None) or {"u": 1.0, "alpha": 0.5}
        alpha, u, inst = prior.get("alpha", 0.5), prior.get("u", 1.0), None
    else:
        match = [(i, t) for i, t in enumerate(fit.instances) if t.name == component]
        if not match:
            raise ValueError("The specified component cannot be found in the fitted model, please check the name.")
        i, inst = match[0]
        marg = mod.marginals[i]
        alpha, u = inst.alpha, inst.u
    if len(marg["theta"]) <= 2:
        raise ValueError("The number of quadrature points is too small, please use aghq_k >= 3.")
    pc = compute_pdf_and_cdf(marg, transformation="sd")
    sigma = pc["transparam"]
    out = {"SD": sigma, "post": pc["pdf_transparam"],
           "prior": (2.0 / sigma) * np.exp(_theta_logprior(-2.0 * np.log(sigma), alpha, u))}
    if h is not None and inst is not None:
        if inst.kind == "IWP":
            p = inst.order
            corr = math.sqrt(h ** (2 * p - 1) / ((2 * p - 1) * math.factorial(p - 1) ** 2))
        elif inst.kind == "sGP":
            corr = sum(compute_d_step_sGPsd(h, j * inst.a) for j in range(1, inst.m + 1))
        else:
            raise ValueError("PSD is currently on defined on IWP and sGP, please specify h = NULL for other type of "
                             "random effect")
        out.update({"PSD": sigma * corr, "post.PSD": out["post"] / corr, "prior.PSD": out["prior"] / corr})
    order = np.argsort(out["SD"], kind="stable")
    return {k2: v[order] for k2, v in out.items()}
