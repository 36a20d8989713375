"""GPU parity of the O-spline moment path (csrc/ospline.cu): models with one IWP term evaluate eta, A^T r and
A^T diag(w) A from knot-interval moments instead of the dense design.  Each case is compared three ways on the same
inputs: against the NumPy oracle (dense restatement of src/BayesGP.cpp:133-252 on get_local_poly's design,
R/01_utility.R:346-364), against the product's own dense DMMA path (bgp_model_set_ospline(m, 0)), and through the
Laplace objective ff$fn / ff$gr / mode / Hessian at the north_star tolerances (1e-8 / 1e-6 relative).
Edge cases: orders 1..4, knots on both sides of the reference location, observations beyond the last knot, exactly on
knots and on the reference location, empty knot intervals, fixed effects up to the 8-column limit, the three families."""
import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu


def make_case(order, family, n=6000, k=24, nfixed=0, two_sided=False, seed=1, gap=False, beyond=False):
    from oracle.fit import Term
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, n)
    if gap:
        x = np.where((x > 0.4) & (x < 0.6), x * 0.4, x)            # several empty knot intervals
    knots = None
    x0 = None
    if two_sided:
        x0 = 0.37
        knots = np.linspace(-0.37, 0.63, k)
        knots = np.sort(np.append(knots, 0.0)) if not np.any(knots == 0.0) else knots
    elif beyond:
        x0 = 0.0
        knots = np.linspace(0.0, 0.8, k)                           # observations in (0.8, 1] lie beyond the last knot
    # observations exactly on knots / on the reference location
    if knots is not None:
        x[:5] = (knots[[0, 1, len(knots) // 2, -2, -1]] + x0)
        x[5] = x0
    fixed = {"z%d" % i: rng.standard_normal(n) for i in range(nfixed)}
    eta = 0.3 + np.sin(2 * np.pi * x) + sum(0.2 * v for v in fixed.values())
    size = None
    if family == "Poisson":
        y = rng.poisson(np.exp(eta)).astype(np.float64)
    elif family == "Binomial":
        size = 1.0 + rng.poisson(6, n)
        y = rng.binomial(size.astype(int), 1 / (1 + np.exp(-eta))).astype(np.float64)
    else:
        y = eta + 0.4 * rng.standard_normal(n)

    def terms():
        return [Term("IWP", "x", x.copy(), order=order, k=k, knots=None if knots is None else knots.copy(),
                     initial_location=x0)]
    return y, terms, fixed, family, size


def build_both(case):
    from bayesgp_b200.api import build_objective
    from oracle.fit import build_model
    y, terms, fixed, family, size = case
    model = build_model(y, terms(), fixed, family=family, size=size)[0]
    ff = build_objective(y, terms(), fixed, family=family, size=size)[0]
    return model, ff


CASES = [
    dict(order=3, family="Poisson"),
    dict(order=1, family="Poisson", nfixed=1),
    dict(order=2, family="Binomial", nfixed=2),
    dict(order=4, family="Poisson", k=16),
    dict(order=3, family="Gaussian", nfixed=1),
    dict(order=3, family="Poisson", two_sided=True),
    dict(order=2, family="Binomial", two_sided=True, nfixed=3),
    dict(order=3, family="Poisson", beyond=True, gap=True),
    dict(order=3, family="Poisson", nfixed=5),            # 2 + 1 + 5 = 8 dense columns: the limit
    dict(order=4, family="Gaussian", nfixed=4, k=12),     # 3 + 1 + 4 = 8
]


@pytest.mark.parametrize("kw", CASES, ids=lambda kw: "-".join("%s%s" % (k[0], v) for k, v in kw.items()))
def test_objective_against_oracle_and_dense_path(kw):
    model, ff = build_both(make_case(seed=31 + len(str(kw)), **kw))
    try:
        assert ff.ospline() == (True, True)
        assert ff.p == model.p
        rng = np.random.default_rng(9)
        S = model.S
        for rep in range(2):
            W = 0.05 * rng.standard_normal(model.p)
            theta = np.array([-2.0 - rep] + ([0.7] if S == 2 else []))
            o = model.objective(W, theta, "fgH")
            ff.set_ospline(True)
            f, g, H = ff.objective(W, theta, want_grad=True, want_hess=True)
            assert abs(f - o["f"]) <= 1e-11 * abs(o["f"])
            assert relerr(g, o["g"]) < 1e-10 and relerr(H, o["H"]) < 1e-10
            assert np.array_equal(H, H.T)
            ff.set_ospline(False)
            fd, gd, Hd = ff.objective(W, theta, want_grad=True, want_hess=True)
            assert abs(f - fd) <= 1e-13 * abs(fd)
            assert relerr(g, gd) < 1e-12 and relerr(H, Hd) < 1e-12
            # entry by entry, relative to the entry's own scale (the moment path has no cancellation to hide behind
            # the matrix norm): sqrt(H_ii H_jj) bounds |H_ij| for the likelihood part
            d = np.sqrt(np.abs(np.diag(Hd)))
            assert np.max(np.abs(H - Hd) / np.outer(d, d)) < 1e-11
    finally:
        ff.close()


@pytest.mark.parametrize("kw", [CASES[0], CASES[2], CASES[4], CASES[5], CASES[7]],
                         ids=["pois3", "binom2", "gauss3", "two-sided", "beyond-gap"])
def test_laplace_objective(kw):
    from oracle.laplace import LaplaceObjective as OFF
    model, ff = build_both(make_case(seed=77, **kw))
    off = OFF(model)
    try:
        S = model.S
        for th0 in (-1.0, -3.5):
            th = np.array([th0] + ([0.5] if S == 2 else []))
            want = off.fn(th)
            ff.set_ospline(True)
            got, _, w, Hm = ff._eval(th, want_hess=True)
            assert abs(got - want) <= 1e-8 * abs(want), (th, got, want)
            assert relerr(w, off.last_par) < 1e-6 and relerr(Hm, off.sp_hess()) < 1e-6
            gw, gg = off.gr(th), ff.gr(th)                    # leverages from per-interval quadratic forms
            assert np.max(np.abs(gw - gg)) <= 2e-7 * max(1.0, np.max(np.abs(gw))), (th, gw, gg)
            assert ff.ospline() == (True, True)
            ff.set_ospline(2)                                 # moment path, leverages from the dense design rows
            g2 = ff.gr(th)
            assert np.max(np.abs(g2 - gg)) <= 1e-9 * max(1.0, np.max(np.abs(g2))), (th, g2, gg)
            ff.set_ospline(False)
            g0 = ff.gr(th)
            assert np.max(np.abs(g0 - gg)) <= 1e-9 * max(1.0, np.max(np.abs(g0))), (th, g0, gg)
            ff.set_start(None)
            gd = ff._eval(th, want_hess=True)
            assert abs(got - gd[0]) <= 1e-10 * abs(want)
            assert relerr(w, gd[2]) < 1e-7
        # repeated evaluations are bit-identical (fixed reduction order)
        ff.set_ospline(True)
        ff.set_start(None)
        a = ff._eval(th, want_hess=True)
        ff.set_start(None)
        b = ff._eval(th, want_hess=True)
        assert a[0] == b[0] and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    finally:
        ff.close()


def test_model_fit_matches_dense_path_and_oracle():
    """model_fit() end to end on the moment path: log normalising constant, theta mode, modes and Hessians per node."""
    from bayesgp_b200.api import build_objective, marginal_laplace_tmb
    from oracle.fit import build_model
    from oracle.laplace import LaplaceObjective as OFF
    from oracle.aghq import marginal_laplace_tmb as oracle_mlt
    case = make_case(order=3, family="Poisson", n=8000, k=30, nfixed=1, seed=5)
    y, terms, fixed, family, size = case
    model = build_model(y, terms(), fixed, family=family, size=size)[0]
    want = oracle_mlt(OFF(model), 5, np.zeros(1))
    res = {}
    for on in (True, False):
        ff = build_objective(y, terms(), fixed, family=family, size=size)[0]
        try:
            ff.set_ospline(on)
            mod = marginal_laplace_tmb(ff, 5, np.zeros(1))
            mh = mod.modesandhessians
            res[on] = (mod.lognormconst, np.array(mod.optresults["mode"]), np.array(mh["mode"]), np.array(mh["H"]))
            mod.close()
        finally:
            ff.close()
    assert abs(res[True][0] - want.lognormconst) <= 1e-8 * abs(want.lognormconst)
    assert abs(res[True][0] - res[False][0]) <= 1e-9 * abs(res[False][0])
    assert relerr(res[True][1], np.array(want.mode)) < 1e-5
    assert relerr(res[True][2], res[False][2]) < 1e-6 and relerr(res[True][3], res[False][3]) < 1e-6


def test_skewed_intervals_and_unsorted_knots():
    """90 % of 200k observations in one knot interval: that interval is cut into far more than 64 pieces (the
    cooperative branch of the interval reduction); unsorted knots: not eligible, dense path, same answer as the
    oracle either way."""
    from bayesgp_b200.api import build_objective
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(11)
    n = 200_000
    x = np.where(rng.uniform(size=n) < 0.9, rng.uniform(0.500, 0.539, n), rng.uniform(0, 1, n))
    y = rng.poisson(np.exp(0.4 + np.sin(2 * np.pi * x))).astype(np.float64)
    knots = np.linspace(0.0, 1.0, 26)
    mk = lambda kn: [Term("IWP", "x", x.copy(), order=3, knots=kn.copy(), initial_location=0.0)]
    model = build_model(y, mk(knots), {}, family="Poisson")[0]
    ff = build_objective(y, mk(knots), {}, family="Poisson")[0]
    try:
        assert ff.ospline() == (True, True)
        W = 0.05 * rng.standard_normal(model.p)
        theta = np.array([-1.5])
        o = model.objective(W, theta, "fgH")
        f, g, H = ff.objective(W, theta, want_grad=True, want_hess=True)
        assert abs(f - o["f"]) <= 1e-11 * abs(o["f"])
        assert relerr(g, o["g"]) < 1e-10 and relerr(H, o["H"]) < 1e-10
        f2, g2, H2 = ff.objective(W, theta, want_grad=True, want_hess=True)
        assert f == f2 and np.array_equal(g, g2) and np.array_equal(H, H2)
    finally:
        ff.close()
    shuffled = knots.copy()
    shuffled[[3, 7]] = shuffled[[7, 3]]
    ff = build_objective(y[:5000], [Term("IWP", "x", x[:5000].copy(), order=3, knots=shuffled, initial_location=0.0)], {},
                         family="Poisson")[0]
    try:
        assert ff.ospline() == (False, False)
    finally:
        ff.close()


def test_lanes_match_one_lane_and_oracle():
    """Evaluation lanes (bgp_model_set_lanes): the nodes of a batch / a quadrature grid dealt to concurrent contexts.
    Same values, modes and Hessians as the one-lane run (to the inner solve's tolerance), same fit as the oracle;
    failing nodes keep their NaN, one lane's failure does not disturb the others."""
    from bayesgp_b200.api import build_objective, marginal_laplace_tmb
    from oracle.fit import build_model
    from oracle.laplace import LaplaceObjective as OFF
    from oracle.aghq import marginal_laplace_tmb as oracle_mlt
    y, terms, fixed, family, size = make_case(order=3, family="Poisson", n=20000, k=40, nfixed=1, seed=8)
    model = build_model(y, terms(), fixed, family=family, size=size)[0]
    ff = build_objective(y, terms(), fixed, family=family, size=size)[0]
    try:
        thetas = np.linspace(-4.0, 1.0, 13)[:, None]
        ff.set_lanes(1)
        ff.set_start(None)
        v1, m1, H1, _ = ff.fn_batch(thetas, want_modes=True, want_hess=True)
        for lanes in (2, 4, 6):
            ff.set_lanes(lanes)
            assert ff.lanes() == lanes
            ff.set_start(None)
            v, mm, HH, it = ff.fn_batch(thetas, want_modes=True, want_hess=True)
            assert np.max(np.abs(v - v1) / np.abs(v1)) < 1e-10
            assert relerr(mm, m1) < 1e-7 and relerr(HH, H1) < 1e-7
            assert it >= len(thetas)
        off = OFF(model)
        for j in (0, 6, 12):
            want = off.fn(thetas[j])
            assert abs(v[j] - want) <= 1e-8 * abs(want)
            assert relerr(mm[j], off.last_par) < 1e-6
        # a node outside the posterior fails alone
        ff.set_lanes(4)
        bad = np.vstack([thetas, [[-80.0]]])
        ff.set_start(None)
        vb = ff.fn_batch(bad, want_modes=False)[0]
        assert np.max(np.abs(vb[:13] - v1) / np.abs(v1)) < 1e-10
        # the whole fit through the lanes
        want = oracle_mlt(OFF(model), 7, np.zeros(1))
        mod = marginal_laplace_tmb(ff, 7, np.zeros(1))
        assert abs(mod.lognormconst - want.lognormconst) <= 1e-8 * abs(want.lognormconst)
        mh = mod.modesandhessians
        assert relerr(mh["mode"], want.modes) < 1e-6 and relerr(mh["H"], want.hessians) < 1e-6
        mod.close()
    finally:
        ff.close()


def test_tiny_and_left_of_first_knot():
    """40 observations on 6 knots (fewer pieces than resident warps), and observations left of the first knot of an
    all-positive knot sequence (every spline column zero there, R/01_utility.R:351)."""
    from bayesgp_b200.api import build_objective
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(4)
    for n, knots, lo in ((40, np.linspace(0.0, 1.0, 7), 0.0), (3000, np.linspace(0.2, 1.0, 12), -0.3)):
        x = rng.uniform(lo, 1.0, n)
        y = rng.poisson(np.exp(0.2 + x)).astype(np.float64)
        mk = lambda: [Term("IWP", "x", x.copy(), order=2, knots=knots.copy(), initial_location=0.0)]
        model = build_model(y, mk(), {}, family="Poisson")[0]
        ff = build_objective(y, mk(), {}, family="Poisson")[0]
        try:
            assert ff.ospline() == (True, True)
            W = 0.1 * rng.standard_normal(model.p)
            theta = np.array([-1.0])
            o = model.objective(W, theta, "fgH")
            f, g, H = ff.objective(W, theta, want_grad=True, want_hess=True)
            assert abs(f - o["f"]) <= 1e-11 * abs(o["f"])
            assert relerr(g, o["g"]) < 1e-10 and relerr(H, o["H"]) < 1e-10
        finally:
            ff.close()


@pytest.mark.parametrize("family", ["Gaussian", "Binomial"])
def test_lanes_two_dimensional_grid(family):
    """Lanes on an S = 2 grid (Gaussian: the noise theta) and on the Binomial family: fit with 4 lanes against the fit
    with one lane and against the oracle."""
    from bayesgp_b200.api import build_objective, marginal_laplace_tmb
    from oracle.fit import build_model
    from oracle.laplace import LaplaceObjective as OFF
    from oracle.aghq import marginal_laplace_tmb as oracle_mlt
    y, terms, fixed, fam, size = make_case(order=2, family=family, n=6000, k=16, nfixed=1, seed=21)
    model = build_model(y, terms(), fixed, family=fam, size=size)[0]
    S = model.S
    want = oracle_mlt(OFF(model), 5, np.zeros(S))
    res = {}
    for lanes in (1, 4):
        ff = build_objective(y, terms(), fixed, family=fam, size=size)[0]
        try:
            ff.set_lanes(lanes)
            mod = marginal_laplace_tmb(ff, 5, np.zeros(S))
            mh = mod.modesandhessians
            res[lanes] = (mod.lognormconst, np.array(mh["mode"]), np.array(mh["H"]),
                          np.array(mod.normalized_posterior["nodesandweights"]["logpost"]))
            mod.close()
        finally:
            ff.close()
    assert abs(res[4][0] - want.lognormconst) <= 1e-8 * abs(want.lognormconst)
    assert abs(res[4][0] - res[1][0]) <= 1e-10 * abs(res[1][0])
    assert relerr(res[4][3], res[1][3]) < 1e-10
    assert relerr(res[4][1], res[1][1]) < 1e-6 and relerr(res[4][2], res[1][2]) < 1e-6
    assert relerr(res[4][1], want.modes) < 1e-6


def test_eligibility():
    """Two smoothing terms, order above 4, more than 8 dense columns, caller-supplied designs: dense path only."""
    from bayesgp_b200 import BgpError, make_objective
    from bayesgp_b200.api import build_objective
    from helpers import synth_poisson, tmbdata_from_oracle
    from oracle.fit import Term
    rng = np.random.default_rng(3)
    n = 2000
    x1, x2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    y = rng.poisson(np.exp(0.2 + np.sin(3 * x1))).astype(np.float64)
    for terms, fixed in (([Term("IWP", "a", x1, order=2, k=10), Term("IWP", "b", x2, order=2, k=10)], {}),
                         ([Term("IWP", "a", x1, order=5, k=10)], {}),
                         ([Term("IWP", "a", x1, order=3, k=10)], {"z%d" % i: rng.standard_normal(n) for i in range(6)})):
        ff = build_objective(y, terms, fixed, family="Poisson")[0]
        try:
            assert ff.ospline() == (False, False)
            with pytest.raises(BgpError):
                ff.set_ospline(True)
        finally:
            ff.close()
    ff = make_objective(tmbdata_from_oracle(synth_poisson(n=3000, k=12)[0]))
    try:
        assert ff.ospline() == (False, False)
    finally:
        ff.close()
