"""CPU check of the identities behind the product's O-spline moment path (csrc/ospline.cu), restated in
oracle/moments.py: eta, B^T r, B^T diag(w) B and D^T diag(w) B from knot-interval moments equal the dense products with
the design of get_local_poly (R/01_utility.R:346-364, oracle/basis.py) — orders 1..4, uneven knots, observations on
knots, left of the first and beyond the last knot."""
import numpy as np
import pytest

from oracle import basis, moments


@pytest.mark.parametrize("P", [1, 2, 3, 4])
@pytest.mark.parametrize("uneven", [False, True])
def test_moment_identities_match_dense_design(P, uneven):
    rng = np.random.default_rng(10 * P + uneven)
    K, n = 14, 900
    knots = np.linspace(0.0, 1.0, K + 1)
    if uneven:
        knots = np.concatenate([[0.0], np.cumsum(rng.uniform(0.3, 1.7, K))])
        knots /= knots[-1]
    z = rng.uniform(-0.05, 1.08, n)                       # some left of the first knot, some beyond the last
    z[:K + 1] = knots                                       # exactly on every knot
    B = basis.get_local_poly(knots, z, P)
    assert np.allclose(B, basis.get_local_poly_loop(knots, z, P), rtol=1e-14, atol=0)      # vectorised form vs the literal double loop: a few ulps
    D = np.column_stack([np.ones(n), z, rng.standard_normal(n)])
    U = rng.standard_normal(K)
    r = rng.standard_normal(n)
    w = np.exp(rng.normal(0, 1, n))
    scale = lambda a: np.max(np.abs(a))
    # eta as a piecewise polynomial
    eta = moments.eta_from_coefficients(knots, z, P, U)
    assert np.max(np.abs(eta - B @ U)) <= 1e-13 * scale(B @ U)
    R, V, X = moments.interval_moments(knots, z, P, r, w, D)
    # B^T r and the {dense x spline} block
    g = moments.apply_transpose(knots, P, R)
    assert np.max(np.abs(g - B.T @ r)) <= 1e-12 * scale(B.T @ r)
    for c in range(D.shape[1]):
        hdb = moments.apply_transpose(knots, P, X[:, c, :])
        want = B.T @ (w * D[:, c])
        assert np.max(np.abs(hdb - want)) <= 1e-12 * scale(want)
    # B^T diag(w) B, entry by entry relative to sqrt(H_ii H_jj)
    H = moments.hessian_block(knots, P, V)
    want = B.T @ (w[:, None] * B)
    d = np.sqrt(np.diag(want))
    nz = d > 0
    assert np.max(np.abs(H - want)[np.ix_(nz, nz)] / np.outer(d[nz], d[nz])) < 1e-12
    assert np.array_equal(H, H.T)


def test_locate_conventions():
    """z <= first knot: no column; z on a knot belongs to the interval on its left (x <= knots[i] is zero for column i,
    R/01_utility.R:351-353); z beyond the last knot: interval K, every column a tail."""
    knots = np.array([0.0, 0.5, 1.0])
    J, u = moments.locate(knots, np.array([-0.2, 0.0, 0.25, 0.5, 0.75, 1.0, 1.3]))
    assert J.tolist() == [0, 0, 0, 0, 1, 1, 2]
    assert np.allclose(u, [0.0, 0.0, 0.25, 0.5, 0.25, 0.5, 0.3])
