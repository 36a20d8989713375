"""GPU tests at BASELINE.json's full size (C3: n = 1e6, IWP3 k = 300, p = 302): one direct comparison with the
oracle at a size it finishes in seconds, and size-independent properties of the hot path at n = 1e6
(row-order invariance, linearity over observation blocks, batch == single, determinism)."""
import os

import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu


def _c3(n):
    from bayesgp_b200.workloads import c3_data, iwp_knots
    x, y = c3_data(n)
    x0, knots = iwp_knots(x, 300)
    return x, y, x0, knots


def _build(x, y, x0, knots, sort=True):
    from bayesgp_b200.objective import LaplaceObjective
    old = os.environ.get("BGP_NO_SORT")
    if not sort:
        os.environ["BGP_NO_SORT"] = "1"
    try:
        ff = LaplaceObjective(y=y, family="Poisson")
        ff.add_iwp(x, x0, knots, 3)
        ff.add_fixed(np.ones(len(y)))
        ff.finalize()
    finally:
        if not sort:
            if old is None:
                del os.environ["BGP_NO_SORT"]
            else:
                os.environ["BGP_NO_SORT"] = old
    return ff


def test_c3_200k_matches_oracle():
    """Same generator as the bench at n = 2e5 (the oracle needs ~10 s): Laplace value 1e-8, mode / Hessian 1e-6."""
    from oracle.fit import Term, build_model
    from oracle.laplace import LaplaceObjective as OFF
    x, y, x0, knots = _c3(200_000)
    model = build_model(y, [Term("IWP", "x", x, order=3, k=300)], {}, family="Poisson")[0]
    off = OFF(model)
    ff = _build(x, y, x0, knots)
    try:
        assert ff.p == model.p == 302
        for theta in (np.array([-9.0]), np.array([-10.5])):
            want = off.fn(theta)
            got, _, w, H = ff._eval(theta, want_hess=True)
            assert abs(got - want) <= 1e-8 * abs(want), (theta, got, want)
            assert relerr(w, off.last_par) < 1e-6
            assert relerr(H, off.sp_hess()) < 1e-6
    finally:
        ff.close()


@pytest.fixture(scope="module")
def c3_full():
    x, y, x0, knots = _c3(1_000_000)
    ff = _build(x, y, x0, knots)
    yield x, y, x0, knots, ff
    ff.close()


def test_full_size_row_order_invariance(c3_full):
    """f, g, H are sums over observations: the zero-pattern sort (and the skipping of empty cells it enables)
    must not change them beyond summation-order rounding."""
    x, y, x0, knots, ff = c3_full
    ff2 = _build(x, y, x0, knots, sort=False)
    try:
        hf, hf2 = ff.hessian_flops(), ff2.hessian_flops()
        assert hf["structural"] < 0.5 * hf["dense"] and hf2["structural"] > 0.9 * hf2["dense"]
        rng = np.random.default_rng(3)
        W = 0.02 * rng.standard_normal(ff.p)
        theta = np.array([-10.0])
        f1, g1, H1 = ff.objective(W, theta, want_grad=True, want_hess=True)
        f2, g2, H2 = ff2.objective(W, theta, want_grad=True, want_hess=True)
        assert abs(f1 - f2) <= 1e-12 * abs(f2)
        assert relerr(g1, g2) < 1e-10
        assert relerr(H1, H2) < 1e-11
        assert np.array_equal(H1, H1.T)
        v1, v2 = ff.fn(np.array([-10.5])), ff2.fn(np.array([-10.5]))
        assert abs(v1 - v2) <= 1e-11 * abs(v2)
    finally:
        ff2.close()


def test_full_size_linearity_over_observation_blocks(c3_full):
    """H(all) - Q = (H(first half) - Q) + (H(second half) - Q) and the same for g and the log-likelihood part."""
    x, y, x0, knots, ff = c3_full
    n = len(y)
    h = n // 2
    # same knots / location for the halves so the designs are row blocks of the full design
    fa = _build(x[:h], y[:h], x0, knots)
    fb = _build(x[h:], y[h:], x0, knots)
    try:
        rng = np.random.default_rng(4)
        W = 0.02 * rng.standard_normal(ff.p)
        theta = np.array([-10.0])
        f, g, H = ff.objective(W, theta, want_grad=True, want_hess=True)
        f_a, g_a, H_a = fa.objective(W, theta, want_grad=True, want_hess=True)
        f_b, g_b, H_b = fb.objective(W, theta, want_grad=True, want_hess=True)
        # prior terms (Q, Q (W - mu0), -lpW - lpT) appear once in the full model and once in each half
        f0, g0, H0 = _prior_only(W, theta, x0, knots)
        assert relerr(H_a + H_b - H0, H) < 1e-11
        assert relerr(g_a + g_b - g0, g) < 1e-10
        assert abs((f_a + f_b - f0) - f) <= 1e-11 * abs(f)
    finally:
        fa.close()
        fb.close()


def _prior_only(W, theta, x0, knots):
    """f, g, H of the same model with the likelihood switched off (family "none", src/BayesGP.cpp:212-214)."""
    from bayesgp_b200.objective import LaplaceObjective
    x = x0 + np.linspace(0.0, float(knots[-1]), 64)
    y = np.zeros(64)
    f0 = LaplaceObjective(y=y, family="none")
    f0.add_iwp(x, x0, knots, 3)
    f0.add_fixed(np.ones(len(y)))
    f0.finalize()
    try:
        return f0.objective(W, theta, want_grad=True, want_hess=True)
    finally:
        f0.close()


def test_full_size_factor_reuse_is_within_its_certificate(c3_full):
    """log det from the last Newton factor vs the recomputed one: |difference| <= 1e-10 |L| (library default)."""
    x, y, x0, knots, ff = c3_full
    thetas = np.array([[-10.9], [-10.7], [-10.5], [-10.3], [-10.1]])
    ff.set_start(None)
    c0 = ff.counters()
    v_reuse, m_reuse, H_reuse, _ = ff.fn_batch(thetas, want_modes=True, want_hess=True)
    c1 = ff.counters()
    ff.set_factor_reuse(False)
    try:
        ff.set_start(None)
        v_exact, m_exact, H_exact, _ = ff.fn_batch(thetas, want_modes=True, want_hess=True)
    finally:
        ff.set_factor_reuse(True)
    assert c1["factor_reuses"] - c0["factor_reuses"] >= 1          # the shortcut was actually exercised
    assert np.max(np.abs(v_reuse - v_exact) / np.abs(v_exact)) <= 1e-10
    assert relerr(m_reuse, m_exact) < 1e-8
    assert relerr(H_reuse, H_exact) < 1e-6


def test_full_size_batch_equals_single_and_is_deterministic(c3_full):
    x, y, x0, knots, ff = c3_full
    thetas = np.array([[-10.8], [-10.5], [-10.2]])
    ff.set_start(None)
    v1, m1, H1, _ = ff.fn_batch(thetas, want_modes=True, want_hess=True)
    ff.set_start(None)
    v2, m2, H2, _ = ff.fn_batch(thetas, want_modes=True, want_hess=True)
    assert np.array_equal(v1, v2) and np.array_equal(m1, m2) and np.array_equal(H1, H2)      # bit-reproducible
    ff.set_start(None)
    singles = [ff.fn(t) for t in thetas]
    assert np.max(np.abs(np.array(singles) - v1) / np.abs(v1)) < 1e-12
    # the gradient of the Laplace objective against a central difference of its values
    g = ff.gr(thetas[1])
    # five-point stencil: the values (~1e6 in magnitude) carry ~1e-8 of summation noise, so the step has to be wide
    # (noise / step ~ 1e-5) and the truncation error of a wide step has to be of fourth order
    eps = 2e-3
    ff.set_factor_reuse(False)       # a difference of close values needs them to ~1e-11 relative, not 1e-10
    try:
        f = {k: ff.fn(thetas[1] + k * eps) for k in (-2, -1, 1, 2)}
    finally:
        ff.set_factor_reuse(True)
    fd = (f[-2] - 8.0 * f[-1] + 8.0 * f[1] - f[2]) / (12.0 * eps)
    assert abs(g[0] - fd) <= 1e-5 * max(1.0, abs(fd)), (g, fd)


def test_c4_full_size_value_mode_hessian_match_oracle():
    """BASELINE C4 at its full size (Binomial, n = 1e6, IWP2 k = 440 + sGP k = 20, p = 497; the generator of
    scripts/run_config.py): one Laplace evaluation against the oracle — value 1e-8, mode and Hessian 1e-6, gradient
    2e-7 of its largest entry.  About a minute of host time for the dense NumPy side."""
    import bayesgp_b200 as bg
    from bayesgp_b200 import api
    from oracle import fit as ofit
    from oracle.laplace import LaplaceObjective as OFF
    n = 1_000_000
    rng = np.random.default_rng(20244)
    x1, x2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    eta = -0.3 + np.sin(2 * np.pi * x1) + 0.6 * np.cos(2 * np.pi * 5 * x2)
    size = 1.0 + rng.poisson(9, n)
    y = rng.binomial(size.astype(int), 1 / (1 + np.exp(-eta))).astype(np.float64)
    kn = np.linspace(0, x1.max() - x1.min(), 440)
    mk = lambda T: [T("IWP", "x1", x1, order=2, knots=kn, initial_location=float(x1.min())),
                    T("sGP", "x2", x2, a=2 * np.pi * 5, k=20, m=1, region=np.array([0.0, 1.0]), accuracy=0.01)]
    ff = api.build_objective(y, mk(bg.Term), {}, "Binomial", size)[0]
    try:
        assert ff.p == 497 and ff.n == n
        theta = np.array([-3.4, 3.7])                 # next to the posterior mode of this data set
        got, g, w, H = ff._eval(theta, want_grad=True, want_hess=True)
        model = ofit.build_model(y, mk(ofit.Term), {}, "Binomial", size)[0]
        off = OFF(model)
        # the dense host side starts NEXT to the device's mode (a cold start would cost it five more Newton iterations of
        # ~5e11 flops each) and converges by its own test
        off.last_par = w * (1.0 + 1e-3)
        want = off.fn(theta)
        assert off.newton_iters >= 1
        assert abs(got - want) <= 1e-8 * abs(want), (got, want)
        assert relerr(w, off.last_par) < 1e-6
        assert relerr(H, off.sp_hess()) < 1e-6
        gw = off.gr(theta)
        assert np.max(np.abs(g - gw)) <= 2e-7 * max(1.0, np.max(np.abs(gw))), (g, gw)
    finally:
        ff.close()
