"""Exactness of the per-row summary of predict (extract_mean_interval_given_samps, R/03_post_fit.R:287-296):
the two type-7 quantiles must be the exact order statistics of the row the device produced, on the candidate /
bucket path and on the radix fallback alike.

Rows are shaped through the public call: an order-1 IWP evaluated inside its first knot interval gives
F[g, :] = intercept_samps + (x_g - knot_0) * coef[0, :], so any sample distribution can be put in a row.
`only_samples=True` returns the device's own F next to the summary, which makes the check bit-level
(no GEMM rounding between the two sides).
"""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _type7(F, prob):
    """stats::quantile type 7 on each row of F, R's own interpolation expression."""
    M = F.shape[1]
    srt = np.sort(F, axis=1)
    index = 1.0 + (M - 1) * prob
    lo = math.floor(index)
    h = index - lo
    lo_i = min(max(int(lo) - 1, 0), M - 1)
    hi_i = min(lo_i + 1, M - 1)
    x_lo, x_hi = srt[:, lo_i], srt[:, hi_i]
    if h == 0.0:
        return x_lo
    return np.where(x_hi != x_lo, (1.0 - h) * x_lo + h * x_hi, x_lo)


def _rows(kind, M, rng):
    """(intercept samples, slope samples) of the row family `kind`."""
    z = rng.standard_normal(M)
    if kind == "gaussian":
        return z, 0.7 * rng.standard_normal(M)
    if kind == "offset":                       # tiny spread on a large level: variance by cancellation would be lost
        return 1.0e6 + 1.0e-3 * z, 1.0e-3 * rng.standard_normal(M)
    if kind == "ties":                         # a few distinct values only
        return np.round(3.0 * z), np.zeros(M)
    if kind == "constant":
        return np.full(M, 2.5), np.zeros(M)
    if kind == "heavy":                        # cuts at mean +- z sd miss the ranks
        return rng.standard_cauchy(M), np.zeros(M)
    if kind == "skewed":
        return rng.exponential(1.0, M) - 5.0, 0.1 * rng.standard_normal(M)
    if kind == "bimodal":
        return np.where(rng.random(M) < 0.5, -1.0, 1.0) + 1.0e-3 * z, np.zeros(M)
    if kind == "signed_zero":
        return np.where(rng.random(M) < 0.5, -0.0, 0.0) + np.where(rng.random(M) < 0.1, z, 0.0), np.zeros(M)
    raise ValueError(kind)


KINDS = ["gaussian", "offset", "ties", "constant", "heavy", "skewed", "bimodal", "signed_zero"]


def _check(M, G, level, kind, seed):
    import bayesgp_b200 as bg
    rng = np.random.default_rng(seed)
    icpt, slope = _rows(kind, M, rng)
    knots = np.array([0.0, 1.0, 2.0])
    coef = np.vstack([slope, np.zeros(M)])
    xg = np.linspace(0.05, 0.95, G)
    out = bg.compute_post_fun_IWP(coef, None, knots, xg, 1, 0, icpt, level=level, only_samples=True)
    F = np.ascontiguousarray(out["samples"])
    assert F.shape == (G, M)
    alpha = 1.0 - level
    lo, hi = _type7(F, alpha / 2.0), _type7(F, level + alpha / 2.0)
    # bit for bit: the device spells the interpolation with R's two products and one sum, no FMA
    bad_lo = out["plower"] != lo
    bad_hi = out["pupper"] != hi
    assert not bad_lo.any(), (kind, M, level, np.flatnonzero(bad_lo)[:5], out["plower"][bad_lo][:3], lo[bad_lo][:3])
    assert not bad_hi.any(), (kind, M, level, np.flatnonzero(bad_hi)[:5], out["pupper"][bad_hi][:3], hi[bad_hi][:3])
    scale = np.maximum(np.abs(F).max(axis=1), 1e-300)
    if np.isfinite(F).all():
        assert np.max(np.abs(out["mean"] - F.mean(axis=1)) / scale) < 1e-13


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("M", [64, 65, 1000, 3001, 10000])
def test_quantiles_are_exact_order_statistics(kind, M):
    for level in (0.95, 0.5, 0.999):
        _check(M, 37, level, kind, seed=1000 + M)


@pytest.mark.parametrize("kind", ["gaussian", "heavy", "ties"])
def test_rows_too_long_for_shared_memory(kind):
    _check(30000, 9, 0.95, kind, seed=5)
    _check(100001, 5, 0.9, kind, seed=6)


def test_tiny_sample_counts():
    for M in (1, 2, 3, 7, 63):
        _check(M, 5, 0.95, "gaussian", seed=M)


def test_fast_path_and_radix_path_agree_bitwise():
    """BGP_SELECT_RADIX=1 forces the fallback for every row; both paths must give the same bits."""
    code = r"""
import numpy as np, bayesgp_b200 as bg
rng = np.random.default_rng(3)
M, G = 10000, 200
coef = np.vstack([0.7 * rng.standard_normal(M), np.zeros(M)])
out = bg.compute_post_fun_IWP(coef, None, np.array([0.0, 1.0, 2.0]), np.linspace(0.05, 0.95, G), 1, 0,
                              rng.standard_normal(M), level=0.95)
print(out["plower"].tobytes().hex()); print(out["pupper"].tobytes().hex()); print(out["mean"].tobytes().hex())
"""
    root = os.path.join(os.path.dirname(__file__), "..")
    res = []
    for radix in ("", "1"):
        env = dict(os.environ)
        env.pop("BGP_SELECT_RADIX", None)
        if radix:
            env["BGP_SELECT_RADIX"] = radix
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        res.append(r.stdout)
    assert res[0] == res[1]
