"""TEST INFRASTRUCTURE (CPU oracle) — posterior summaries on top of an AGHQ fit.

Restates, for the checker only:
  * aghq::compute_pdf_and_cdf / compute_quantiles on a theta marginal (aghq >= 0.4.1, not on disk; call sites
    /root/reference/R/03_post_fit.R:330,343 and the table printed by summary(), /root/reference/R/03_post_fit.R:2-29):
    natural cubic spline through (theta_j, logmargpost_j) with linear continuation outside the nodes, a 1000-point
    grid over the node range widened by half its length on both sides, pdf = exp(spline),
    cdf = cumsum(pdf * c(0, diff(theta))), quantile q = the last grid point whose cdf is below q.
    PINNED: with the README's grid centre and scale this reproduces the printed 2.5 % / median / 97.5 % of
    theta(t) (-3.87922, -3.268308, -2.760093; /root/reference/README.md:83-85) to the printed digits
    (tests/test_oracle_summary.py).  The polynomial interpolation branch of aghq is not restated (unpinned).
  * var_density (/root/reference/R/03_post_fit.R:309-447, aghq branch): the marginal on the SD scale
    sigma = exp(-theta / 2), the Exponential prior of BayesGP.cpp:241-246 on that scale, and the PSD rescaling
    (IWP: :352-355; sGP: compute_d_step_sGPsd, /root/reference/R/01_utility.R:460-462).
  * the fixed-effect table of summary.FitResult (/root/reference/R/03_post_fit.R:30-41).
"""
from __future__ import annotations

import math

import numpy as np


def natural_spline(x, y):
    """splines::interpSpline(x, y) (natural) + predict(): returns f(xnew); linear outside [x_1, x_k]."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    o = np.argsort(x)
    x, y = x[o], y[o]
    k = len(x)
    if k < 3:
        raise ValueError("The number of quadrature points is too small, please use aghq_k >= 3.")
    h = np.diff(x)
    # second derivatives m_0 = m_{k-1} = 0
    A = np.zeros((k, k))
    rhs = np.zeros(k)
    A[0, 0] = A[-1, -1] = 1.0
    for i in range(1, k - 1):
        A[i, i - 1], A[i, i], A[i, i + 1] = h[i - 1], 2.0 * (h[i - 1] + h[i]), h[i]
        rhs[i] = 6.0 * ((y[i + 1] - y[i]) / h[i] - (y[i] - y[i - 1]) / h[i - 1])
    m = np.linalg.solve(A, rhs)
    d_lo = (y[1] - y[0]) / h[0] - h[0] * (2.0 * m[0] + m[1]) / 6.0
    d_hi = (y[-1] - y[-2]) / h[-1] + h[-1] * (m[-2] + 2.0 * m[-1]) / 6.0

    def f(xn):
        xn = np.asarray(xn, dtype=np.float64)
        i = np.clip(np.searchsorted(x, xn, side="right") - 1, 0, k - 2)
        t0, t1 = xn - x[i], x[i + 1] - xn
        inside = (m[i] * t1 ** 3 + m[i + 1] * t0 ** 3) / (6.0 * h[i]) + (y[i] / h[i] - m[i] * h[i] / 6.0) * t1 + (
            y[i + 1] / h[i] - m[i + 1] * h[i] / 6.0) * t0
        out = np.where(xn < x[0], y[0] + d_lo * (xn - x[0]), inside)
        return np.where(xn > x[-1], y[-1] + d_hi * (xn - x[-1]), out)

    return f


def compute_pdf_and_cdf(marginal, to_sd=False, ngrid=1000):
    """marginal: {"theta", "logmargpost"}.  Returns theta, pdf, cdf (and transparam, pdf_transparam when to_sd:
    transformation totheta = -2 log x, fromtheta = exp(-x / 2) as var_density passes it)."""
    th = np.asarray(marginal["theta"], dtype=np.float64)
    lo, hi = th.min(), th.max()
    ext = 0.5 * (hi - lo)
    grid = np.linspace(lo - ext, hi + ext, ngrid)
    pdf = np.exp(natural_spline(th, marginal["logmargpost"])(grid))
    cdf = np.cumsum(pdf * np.concatenate([[0.0], np.diff(grid)]))
    out = {"theta": grid, "pdf": pdf, "cdf": cdf}
    if to_sd:
        sd = np.exp(-grid / 2.0)
        out["transparam"] = sd
        out["pdf_transparam"] = pdf * np.abs(-2.0 / sd)      # |d totheta / d sigma|
    return out


def compute_quantiles(marginal, q=(0.025, 0.975)):
    pc = compute_pdf_and_cdf(marginal)
    res = []
    for qq in q:
        idx = np.flatnonzero(pc["cdf"] < qq)
        res.append(pc["theta"][idx.max()] if len(idx) else np.nan)
    return np.array(res)


def theta_summary_table(mod):
    """rows of `summary(mod)$summarytable`: mean, sd (quadrature moments), 2.5 %, median, 97.5 %."""
    from .aghq import theta_moments
    mean, sd = theta_moments(mod)
    rows = []
    for j, marg in enumerate(mod.marginals):
        ql, med, qu = compute_quantiles(marg, (0.025, 0.5, 0.975))
        rows.append({"mean": mean[j], "sd": sd[j], "2.5%": ql, "median": med, "97.5%": qu})
    return rows


def theta_logprior(theta, alpha, u):
    lam = -math.log(alpha) / u
    return math.log(lam / 2.0) - lam * np.exp(-theta / 2.0) - theta / 2.0


def psd_correction(kind, h, order=None, a=None, m=1):
    if kind == "IWP":
        p = order
        return math.sqrt(h ** (2 * p - 1) / ((2 * p - 1) * math.factorial(p - 1) ** 2))
    if kind == "sGP":
        return sum(math.sqrt((1.0 / (j * a) ** 2) * (h / 2.0 - math.sin(2.0 * j * a * h) / (4.0 * j * a))) for j in range(1, m + 1))
    raise ValueError("PSD is currently on defined on IWP and sGP, please specify h = NULL for other type of random effect")


def var_density(marginal, alpha, u, kind=None, h=None, order=None, a=None, m=1):
    pc = compute_pdf_and_cdf(marginal, to_sd=True)
    sd = pc["transparam"]
    out = {"SD": sd, "post": pc["pdf_transparam"], "prior": (2.0 / sd) * np.exp(theta_logprior(-2.0 * np.log(sd), alpha, u))}
    if h is not None:
        c = psd_correction(kind, h, order, a, m)
        out.update({"PSD": sd * c, "post.PSD": out["post"] / c, "prior.PSD": out["prior"] / c})
    o = np.argsort(out["SD"], kind="stable")
    return {k2: v[o] for k2, v in out.items()}


def fixed_effect_summary(rows):
    """rows: (#fixed x M) sample rows.  Columns of t(fixed_summary[c(2:5, 7), ]): 1st Qu., Median, Mean, 3rd Qu., sd."""
    rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
    q1, med, q3 = (np.quantile(rows, q, axis=1, method="linear") for q in (0.25, 0.5, 0.75))
    return {"1st Qu.": q1, "Median": med, "Mean": rows.mean(axis=1), "3rd Qu.": q3, "sd": rows.std(axis=1, ddof=1)}
