"""model_fit() / predict.FitResult() — ORACLE restatement (test infrastructure).

Follows ``/root/reference/R/02_model_fit.R:336-701`` (term construction,
defaults, W index maps, sampling call) and ``/root/reference/R/03_post_fit.R:
53-125,159-165,200-296`` (predict, sample_fixed_effect, compute_post_fun_*,
extract_mean_interval_given_samps).  The formula DSL is replaced by a plain
term list (the DSL is out of scope).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import factorial
from typing import Dict, List, Optional

import numpy as np

from . import basis
from .aghq import AGHQFit, marginal_laplace_tmb, node_probabilities, sample_marginal
from .laplace import LaplaceObjective
from .model import FAMILY_CODES, FAMILY_GAUSSIAN, Model


@dataclass
class Term:
    kind: str                     # "IWP" | "sGP" | "IID"
    name: str
    x: np.ndarray
    order: int = 0                # IWP
    knots: Optional[np.ndarray] = None
    k: Optional[int] = None
    initial_location: Optional[float] = None
    a: float = 0.0                # sGP
    m: int = 1
    region: Optional[np.ndarray] = None
    accuracy: float = 0.01
    boundary: bool = True
    u: float = 1.0                # sd.prior$param (R/02_model_fit.R:377)
    alpha: float = 0.5
    boundary_prec: float = 0.01   # R/02_model_fit.R:444-452
    boundary_mean: float = 0.0
    # filled by build
    observed_x: np.ndarray = field(default=None, repr=False)
    X: np.ndarray = field(default=None, repr=False)
    B: np.ndarray = field(default=None, repr=False)
    P: np.ndarray = field(default=None, repr=False)


def build_term(t: Term) -> Term:
    x = np.asarray(t.x, dtype=np.float64)
    if t.kind == "IWP":                                    # R/02_model_fit.R:415-463
        if t.initial_location is None:
            t.initial_location = float(x.min())
        xi = x - t.initial_location
        if t.knots is None:
            t.knots = basis.default_knots(xi, 5 if t.k is None else t.k)
        t.knots = np.asarray(t.knots, dtype=np.float64)
        t.observed_x = np.sort(xi)
        t.X = basis.global_poly_helper(xi, t.order)[:, 1:]
        t.B = basis.local_poly_helper(t.knots, xi, t.order)
        t.P = basis.compute_weights_precision(t.knots)
    elif t.kind == "sGP":                                  # R/02_model_fit.R:493-562
        if t.k is None:
            t.k = 30
        if t.initial_location is None:
            t.initial_location = float(x.min())
        xi = x - t.initial_location
        t.observed_x = np.sort(xi)
        if t.region is None:
            t.region = np.array([t.observed_x[0], t.observed_x[-1]])
        t.X = basis.global_poly_sGP(xi, t.a, t.m)
        # compute_B ignores `boundary` at fit time (A.8 quirk; R/01_utility.R:236)
        t.B = np.concatenate([basis.compute_B_sB(xi, i * t.a, t.k, t.region) for i in range(1, t.m + 1)], axis=1)
        t.P = basis.compute_P_sGP(t.a, t.k, t.m, t.region, t.accuracy)
    elif t.kind == "IID":                                  # R/01_utility.R:214-219,245-250
        lev, inv = np.unique(x, return_inverse=True)
        t.B = np.zeros((len(x), len(lev)))
        t.B[np.arange(len(x)), inv] = 1.0
        t.P = np.ones(len(lev))
        t.X = np.zeros((len(x), 0))
    else:
        raise ValueError(t.kind)
    return t


@dataclass
class FitResult:
    terms: List[Term]
    model: Model
    ff: LaplaceObjective
    mod: AGHQFit
    boundary_samp_indexes: Dict[str, np.ndarray]
    random_samp_indexes: Dict[str, np.ndarray]
    fixed_samp_indexes: Dict[str, int]
    family: str
    samps: Optional[np.ndarray] = None     # p x M


def build_model(y, terms: List[Term], fixed: Dict[str, np.ndarray], family="Gaussian", size=None,
                family_u=1.0, family_alpha=0.5, fixed_prec=None, fixed_mean=None):
    """tmbdat assembly (R/02_model_fit.R:30-183) + index maps (:627-675), 0-based."""
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    terms = [build_term(t) for t in terms]
    fam = FAMILY_CODES[family]
    B, P, logPdet, u, alpha, X, bprec, bmean = [], [], [], [], [], [], [], []
    for t in terms:
        if t.kind in ("IWP", "sGP"):
            X.append(t.X)
            bprec.append(t.boundary_prec)
            bmean.append(t.boundary_mean)
        B.append(t.B)
        P.append(t.P)
        if t.P.ndim == 1:
            logPdet.append(float(np.sum(np.log(t.P))))
        else:
            logPdet.append(float(np.linalg.slogdet(t.P)[1]))      # determinant(P)$modulus (:66)
        u.append(t.u)
        alpha.append(t.alpha)
    if fam == FAMILY_GAUSSIAN:
        u.append(family_u)
        alpha.append(family_alpha)
    fixed_prec = fixed_prec or {}
    fixed_mean = fixed_mean or {}
    Xf = [np.ones((n, 1))]
    names = ["intercept"]
    for nm, col in fixed.items():
        Xf.append(np.asarray(col, dtype=np.float64).reshape(n, 1))
        names.append(nm)
    fprec = [fixed_prec.get(nm, 0.01) for nm in names]
    fmean = [fixed_mean.get(nm, 0.0) for nm in names]
    model = Model(family=fam, y=y, B=B, P=P, logPdet=logPdet, u=u, alpha=alpha, X=X, betaprec=bprec,
                  betamean=bmean, Xf=Xf, beta_fixed_prec=fprec, beta_fixed_mean=fmean, size=size)
    rand_idx, bnd_idx, fix_idx = {}, {}, {}
    o = 0
    for t in terms:
        rand_idx[t.name] = np.arange(o, o + t.B.shape[1])
        o += t.B.shape[1]
    for t in terms:
        if t.kind in ("IWP", "sGP"):
            bnd_idx[t.name] = np.arange(o, o + t.X.shape[1])
            o += t.X.shape[1]
    for nm in names:
        fix_idx[nm] = o
        o += 1
    assert o == model.p
    return model, terms, rand_idx, bnd_idx, fix_idx


def model_fit(y, terms, fixed=None, family="Gaussian", aghq_k=4, size=None, M=3000, rng=None,
              Z=None, node_idx=None, **kw) -> FitResult:
    """model_fit(..., method="aghq")  (R/02_model_fit.R:336-701)."""
    model, terms, rand_idx, bnd_idx, fix_idx = build_model(y, terms, fixed or {}, family, size, **kw)
    ff = LaplaceObjective(model)
    mod = marginal_laplace_tmb(ff, aghq_k, np.zeros(model.S))
    res = FitResult(terms, model, ff, mod, bnd_idx, rand_idx, fix_idx, family)
    if M:
        if Z is None:
            rng = rng or np.random.default_rng(0)
            lam = node_probabilities(mod)
            node_idx = rng.choice(len(lam), size=M, p=lam / lam.sum())
            Z = rng.standard_normal((model.p, M))
        res.samps = sample_marginal(mod, Z, node_idx)
    return res


def sample_fixed_effect(fit: FitResult, variables):
    """R/03_post_fit.R:159-165: M x len(variables)."""
    return fit.samps[[fit.fixed_samp_indexes[v] for v in variables], :].T


def compute_post_fun_IWP(samps, global_samps, knots, refined_x, p, degree=0, intercept_samps=None):
    """R/03_post_fit.R:200-241 — returns G x M matrix (without the x column)."""
    if p <= degree:
        return None
    M = samps.shape[1]
    if global_samps is None:
        global_samps = np.zeros((p - 1, M))
    if intercept_samps is None:
        intercept_samps = np.zeros((1, M))
    gs = np.concatenate([intercept_samps.reshape(1, M), global_samps.reshape(-1, M)], axis=0)
    Bm = basis.local_poly_helper(knots, refined_x, p - degree)
    X = basis.global_poly_helper(refined_x, p)[:, :p - degree]
    for i in range(1, X.shape[1] + 1):
        X[:, i - 1] = (factorial(i + degree - 1) / factorial(i - 1)) * X[:, i - 1]
    return X @ gs[degree:p, :] + Bm @ samps


def compute_post_fun_sGP(samps, global_samps, k, refined_x, a, region, m, boundary=True, intercept_samps=None):
    """R/03_post_fit.R:261-276."""
    M = samps.shape[1]
    Bm = basis.compute_B_sB_helper(refined_x, a, k, m, region, boundary, None)
    X = np.concatenate([np.ones((len(refined_x), 1)), basis.global_poly_helper_sGP(refined_x, a, m)], axis=1)
    if intercept_samps is None:
        intercept_samps = np.zeros((1, M))
    if global_samps is None:
        global_samps = np.zeros((2 * m, M))
    gs = np.concatenate([intercept_samps.reshape(1, M), global_samps], axis=0)
    return X @ gs + Bm @ samps


def quantile7(sorted_rows, q):
    """stats::quantile type 7 on rows already sorted ascending (A.7)."""
    M = sorted_rows.shape[1]
    index = 1.0 + (M - 1) * q
    lo = int(np.floor(index))
    hi = int(np.ceil(index))
    h = index - lo
    return (1.0 - h) * sorted_rows[:, lo - 1] + h * sorted_rows[:, hi - 1]


def extract_mean_interval_given_samps(F, level=0.95):
    """R/03_post_fit.R:287-296 — returns (plower, pupper, mean)."""
    alpha = 1.0 - level
    srt = np.sort(F, axis=1)
    return quantile7(srt, alpha / 2.0), quantile7(srt, level + alpha / 2.0), F.mean(axis=1)


def predict(fit: FitResult, variable, newx=None, degree=0, include_intercept=True, only_samples=False):
    """predict.FitResult  (R/03_post_fit.R:53-125).  Returns dict with x and
    either samples (G x M) or plower/pupper/mean."""
    samps = fit.samps
    term = next(t for t in fit.terms if t.name == variable)
    gsamps = samps[fit.boundary_samp_indexes[variable], :] if variable in fit.boundary_samp_indexes else None
    csamps = samps[fit.random_samp_indexes[variable], :]
    rx = term.observed_x if newx is None else np.sort(np.asarray(newx, dtype=np.float64) - term.initial_location)
    isamps = samps[[fit.fixed_samp_indexes["intercept"]], :] if include_intercept else None
    if term.kind == "IWP":
        F = compute_post_fun_IWP(csamps, gsamps, term.knots, rx, term.order, degree, isamps)
    elif term.kind == "sGP":
        F = compute_post_fun_sGP(csamps, gsamps, term.k, rx, term.a, term.region, term.m, term.boundary, isamps)
    else:
        raise ValueError("predict supports IWP and sGP terms")
    out = {"x": rx + term.initial_location}
    if F is None:
        return None
    if only_samples:
        out["samples"] = F
        return out
    lo, hi, mean = extract_mean_interval_given_samps(F)
    out.update(plower=lo, pupper=hi, mean=mean)
    return out


# ----------------------------------------------------------------------------
# model_fit_loop (R/02_model_fit.R:725-778)
# ----------------------------------------------------------------------------
def fmm_spline_eval(x, y, xo):
    """stats::spline(method = "fmm") evaluated at xo.  The FMM end conditions make the third derivative of the spline at
    each end equal to that of the cubic through the four nearest points: solved here as a dense linear system for the
    second derivatives (independent of the tridiagonal recurrences the product uses)."""
    x, y, xo = np.asarray(x, float), np.asarray(y, float), np.asarray(xo, float)
    m = len(x)
    if m < 3:
        return y[0] + (y[-1] - y[0]) / (x[-1] - x[0]) * (xo - x[0])
    h = np.diff(x)
    # unknowns: second derivatives M_0..M_{m-1}
    A = np.zeros((m, m))
    r = np.zeros(m)
    for i in range(1, m - 1):
        A[i, i - 1], A[i, i], A[i, i + 1] = h[i - 1], 2.0 * (h[i - 1] + h[i]), h[i]
        r[i] = 6.0 * ((y[i + 1] - y[i]) / h[i] - (y[i] - y[i - 1]) / h[i - 1])
    if m > 3:
        def third_derivative(xx, yy):          # of the cubic through four points = 6 * leading coefficient
            return 6.0 * np.polyfit(xx - xx[0], yy, 3)[0]
        A[0, 0], A[0, 1] = -1.0 / h[0], 1.0 / h[0]
        r[0] = third_derivative(x[:4], y[:4])
        A[m - 1, m - 2], A[m - 1, m - 1] = -1.0 / h[-1], 1.0 / h[-1]
        r[m - 1] = third_derivative(x[-4:], y[-4:])
    else:                                        # three points: the FMM conditions degenerate to M_0 = M_1 = M_2 ... (c[1] = c[n] = 0)
        A[0, 0], A[0, 1] = -1.0 / h[0], 1.0 / h[0]
        A[m - 1, m - 2], A[m - 1, m - 1] = -1.0 / h[-1], 1.0 / h[-1]
    M = np.linalg.solve(A, r)
    i = np.clip(np.searchsorted(x, xo, side="right") - 1, 0, m - 2)
    a, b = xo - x[i], x[i + 1] - xo
    return ((M[i] * b ** 3 + M[i + 1] * a ** 3) / (6.0 * h[i]) + (y[i] / h[i] - M[i] * h[i] / 6.0) * b
            + (y[i + 1] / h[i] - M[i + 1] * h[i] / 6.0) * a)


def integrate_xy(x, fx):
    """sfsmisc::integrate.xy with its defaults: fmm spline on max(1024, 3 n) points + trapezoid."""
    x, fx = np.asarray(x, float), np.asarray(fx, float)
    o = np.argsort(x, kind="stable")
    x, fx = x[o], fx[o]
    n = max(1024, 3 * len(x))
    xs = x[0] + (x[-1] - x[0]) * np.arange(n) / (n - 1.0)
    ys = fmm_spline_eval(x, fx, xs)
    if xs[-1] < x[-1]:
        xs, ys = np.append(xs, x[-1]), np.append(ys, fx[-1])
    return float(np.sum(np.diff(xs) * (ys[1:] + ys[:-1]) / 2.0))


def model_fit_loop(loop_values, fit_args, prior_func=None):
    """R/02_model_fit.R:725-778 (sequential branch): log_ml per loop value, posterior normalised by integrate.xy."""
    loop_values = np.asarray(loop_values, float)
    log_ml = np.array([model_fit(**dict(fit_args(float(v)), M=0)).mod.lognormconst for v in loop_values])
    prior = np.ones(len(loop_values)) if prior_func is None else np.asarray(prior_func(loop_values), float) * np.ones(len(loop_values))
    lj = log_ml + np.log(prior)
    lj = lj - lj.max()
    post = np.exp(lj)
    return {"var": loop_values, "post": post / integrate_xy(loop_values, post), "log_ml": log_ml}
