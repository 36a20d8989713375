"""CPU: unit checks of the oracle's building blocks against independent implementations."""
import numpy as np
import pytest

from oracle import basis
from oracle.aghq import gh_product_grid, gh_rule, vmmin
from oracle.fit import quantile7
from oracle.laplace import richardson_jacobian


def test_local_poly_vectorised_equals_literal_loop():
    rng = np.random.default_rng(0)
    for knots in (np.linspace(0, 2, 7), np.array([-1.5, -0.7, -0.2, 0.0, 0.4, 1.1]), np.array([-2.0, -1.0, -0.5, 0.0])):
        x = rng.uniform(knots.min() - 0.2, knots.max() + 0.2, 40)
        x[:3] = knots[:3]                     # hit the knots exactly
        for p in (1, 2, 3, 4):
            a = basis.local_poly_helper(knots, x, p, loop=True)
            b = basis.local_poly_helper(knots, x, p, loop=False)
            assert a.shape == b.shape
            assert np.allclose(a, b, rtol=1e-14, atol=0)


def test_reference_documented_example_shape():
    # man/local_poly_helper.Rd example: knots = c(0, .2, .4, .6, .8), refined_x = seq(0, .8, by = .1), p = 2
    D = basis.local_poly_helper(np.array([0, 0.2, 0.4, 0.6, 0.8]), np.arange(0, 0.81, 0.1), 2)
    assert D.shape == (9, 4)
    assert D[0].tolist() == [0, 0, 0, 0]
    assert abs(D[1, 0] - 0.5 * 0.1 ** 2) < 1e-15            # inside the first interval: x^2 / 2
    assert abs(D[4, 0] - (0.2 * 0.2 + 0.5 * 0.2 ** 2)) < 1e-15


def test_osplines_integrate_to_truncated_powers():
    # phi_i(x) = [(x - k_i)_+^q - (x - k_{i+1})_+^q] / q!
    knots = np.linspace(0, 1, 6)
    x = np.linspace(0, 1.3, 50)
    for q in (1, 2, 3):
        D = basis.get_local_poly(knots, x, q)
        for i in range(5):
            want = (np.maximum(x - knots[i], 0) ** q - np.maximum(x - knots[i + 1], 0) ** q) / np.math.factorial(q) \
                if hasattr(np, "math") else None
            from math import factorial
            want = (np.maximum(x - knots[i], 0) ** q - np.maximum(x - knots[i + 1], 0) ** q) / factorial(q)
            assert np.allclose(D[:, i], want, rtol=1e-12, atol=1e-15)


def test_bspline_basis_against_scipy():
    from scipy.interpolate import BSpline
    k, region = 12, (0.0, 2.0)
    t = basis.bspline_knots(region, k)
    x = np.concatenate([np.linspace(0, 2, 101), [0.0, 2.0, 1.0]])
    for deriv in (0, 1, 2):
        mine = basis.bspline_basis(x, region, k, deriv, drop_first_two=False)
        for j in range(k):
            c = np.zeros(k)
            c[j] = 1.0
            sp = BSpline(t, c, 3, extrapolate=False)
            ref = sp.derivative(deriv)(x) if deriv else sp(x)
            ref = np.nan_to_num(ref)
            inner = (x > 0) & (x < 2)
            assert np.allclose(mine[inner, j], ref[inner], rtol=1e-10, atol=1e-10), (deriv, j)
    B = basis.bspline_basis(x, region, k, 0, drop_first_two=False)
    assert np.allclose(B.sum(1), 1.0)                        # partition of unity incl. both end points
    assert basis.bspline_basis(x, region, k).shape[1] == k - 2


def test_sgp_precision_is_symmetric_psd():
    Q = basis.compute_Q_sB(2 * np.pi, 8, np.array([0.0, 1.0]), accuracy=0.01)
    assert Q.shape == (18, 18)
    assert np.array_equal(Q, Q.T)
    assert np.linalg.eigvalsh(Q).min() > -1e-8 * np.abs(Q).max()


def test_gauss_hermite_rule():
    for k in (1, 3, 4, 7, 15):
        z, w = gh_rule(k)
        assert np.allclose(z, -z[::-1])
        # weights integrate g(z) dz: sum w phi(z) z^(2j) = (2j-1)!! for 2j < 2k
        phi = np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
        assert abs(np.sum(w * phi) - 1.0) < 1e-13
        if k >= 3:
            assert abs(np.sum(w * phi * z ** 2) - 1.0) < 1e-12
            assert abs(np.sum(w * phi * z ** 4) - 3.0) < 1e-11
    z, w = gh_rule(1)
    assert z[0] == 0 and abs(w[0] - np.sqrt(2 * np.pi)) < 1e-15     # k = 1 is the plain Laplace approximation
    nodes, weights = gh_product_grid(2, 3)
    assert nodes.shape == (9, 2) and np.allclose(nodes[:3, 1], nodes[0, 1])   # first coordinate fastest


def test_vmmin_reproduces_R_optim_documented_example():
    f = lambda b: float(100 * (b[1] - b[0] ** 2) ** 2 + (1 - b[0]) ** 2)
    g = lambda b: np.array([-400 * b[0] * (b[1] - b[0] ** 2) - 2 * (1 - b[0]), 200 * (b[1] - b[0] ** 2)])
    res = vmmin(f, g, np.array([-1.2, 1.0]))
    assert res["convergence"] == 0
    assert np.allclose(res["par"], [1.0, 1.0], atol=1e-3)    # R: optim(c(-1.2,1), fr, grr, method="BFGS") -> 1, 1
    # ?optim example output in R: $counts function 110 gradient 43, $value 9.594956e-18
    assert (res["fncount"], res["grcount"]) == (110, 43)
    assert abs(res["value"] - 9.594956e-18) < 1e-23


def test_richardson_jacobian_is_exact_on_quadratics():
    A = np.array([[2.0, 0.3], [-0.1, 1.5]])
    J = richardson_jacobian(lambda x: A @ x + 0.5, np.array([0.7, -1.3]))
    assert np.allclose(J, A, rtol=1e-9)
    J0 = richardson_jacobian(lambda x: A @ x, np.array([0.0, 0.0]))          # |x| < zero.tol branch (h = eps)
    assert np.allclose(J0, A, rtol=1e-9)


def test_quantile_type7():
    rng = np.random.default_rng(3)
    F = rng.standard_normal((5, 3000))
    srt = np.sort(F, axis=1)
    for q in (0.025, 0.975, 0.5):
        assert np.allclose(quantile7(srt, q), np.quantile(F, q, axis=1, method="linear"), rtol=1e-14)
    # M = 3000: index 75.975 / 2925.025
    assert abs((1 + 2999 * 0.025) - 75.975) < 1e-12


def test_laplace_gradient_matches_finite_differences():
    from helpers import synth_poisson
    from oracle.laplace import LaplaceObjective
    model = synth_poisson(n=3000, k=12)[0]
    ff = LaplaceObjective(model)
    th = np.array([2.0])
    g = ff.gr(th)
    h = 1e-5
    fd = (ff.fn(th + h) - ff.fn(th - h)) / (2 * h)
    assert abs(g[0] - fd) < 1e-5 * max(1.0, abs(fd))


def test_binomial_gaussian_gradients_match_finite_differences():
    from helpers import synth_binomial_sgp, synth_gaussian
    from oracle.laplace import LaplaceObjective
    for build, th in ((lambda: synth_binomial_sgp(n=1500), np.array([1.0, -0.5])),
                      (lambda: synth_gaussian(n=1200), np.array([2.0, 1.0]))):
        ff = LaplaceObjective(build()[0])
        g = ff.gr(th)
        for i in range(2):
            e = np.zeros(2)
            e[i] = 1e-5
            fd = (ff.fn(th + e) - ff.fn(th - e)) / 2e-5
            assert abs(g[i] - fd) < 2e-5 * max(1.0, abs(fd)), (i, g, fd)


def test_oracle_against_40_digit_reference():
    """The FP64 oracle itself vs the mpmath values (tests/golden/make_hp_covid.py): the oracle's own rounding
    noise on the ill-conditioned README model is measured, not assumed."""
    import json
    import os
    from helpers import GOLDEN, covid_model
    from oracle.laplace import LaplaceObjective
    hp = json.load(open(os.path.join(GOLDEN, "covid_hp.json")))
    off = LaplaceObjective(covid_model()[0])
    for theta, want, mode in zip(hp["theta"], hp["value"], hp["mode"]):
        got = off.fn(np.array([theta]))
        assert abs(got - want) <= 1e-8 * abs(want), (theta, got - want)
        assert np.max(np.abs(off.last_par - np.array(mode))) <= 1e-6 * np.max(np.abs(mode))
