"""Seeded synthetic workloads of BASELINE.json and the host-side setup both
arms of the benchmark share.  Only covariates / responses are generated here; design matrices are
built by whoever consumes them (the CUDA library from x, or the oracle with its own constructors).
"""
from __future__ import annotations

import numpy as np


def c3_data(n=1_000_000, seed=20243):
    """C3: synthetic Poisson, one IWP3 term with k = 300 knots + intercept (p = 302)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 1.0, n)
    eta = 1.0 + np.sin(2 * np.pi * x) + 0.5 * np.cos(6 * np.pi * x)
    y = rng.poisson(np.exp(eta)).astype(np.float64)
    return x, y


def iwp_knots(x, k):
    """R/02_model_fit.R:429-441: initial_location = min(x), knots = seq(0, max(x - x0), length.out = k)."""
    x0 = float(np.min(x))
    xi = x - x0
    return x0, np.unique(np.sort(np.linspace(xi.min(), xi.max(), k)))


def gh_nodes(k):
    """Probabilists' Gauss-Hermite nodes (mvQuad "GHe"), ascending."""
    from numpy.polynomial.hermite_e import hermegauss
    z, _ = hermegauss(k)
    z = np.sort(z)
    return 0.5 * (z - z[::-1])


def locate_mode_1d(fn, lo, hi, iters=28):
    """Untimed setup helper: golden-section minimiser of a 1-D Laplace objective plus a
    central-difference curvature, used only to centre the benchmark's quadrature nodes."""
    gr = (np.sqrt(5.0) - 1.0) / 2.0
    a, b = lo, hi
    c, d = b - gr * (b - a), a + gr * (b - a)
    fc, fd = fn(np.array([c])), fn(np.array([d]))
    for _ in range(iters):
        if fc < fd:
            b, d, fd = d, c, fc
            c = b - gr * (b - a)
            fc = fn(np.array([c]))
        else:
            a, c, fc = c, d, fd
            d = a + gr * (b - a)
            fd = fn(np.array([d]))
    m = 0.5 * (a + b)
    h = 0.05
    f0, fp, fm = fn(np.array([m])), fn(np.array([m + h])), fn(np.array([m - h]))
    curv = (fp - 2 * f0 + fm) / (h * h)
    return m, (1.0 / np.sqrt(curv) if curv > 0 else 0.25)
