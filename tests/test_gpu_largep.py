"""GPU parity for latent dimensions above 512 (the kernel instances BASELINE.json's C5 runs on): the likelihood pass
with one observation per consumer warp on 10..16 column groups, the Cholesky with 16-wide cluster panels, the
Hessian work list near its tile limit.  Same comparisons as the small fixtures (tests/test_gpu_core.py):
objective f / g / H at a random W (1e-11 / 1e-10), ff$fn (1e-8 relative), mode and Hessian at the mode (1e-6),
ff$gr (2e-7 relative to its largest entry) — CUDA through the C ABI against the NumPy oracle on the same inputs.
Then the two large BASELINE configurations at their full latent size and full quadrature grid, observations
reduced so that the oracle finishes in a minute: C4 (Binomial, IWP2 + sGP, p = 497, 7^2 nodes) and
C5 (Poisson, three IWP3 terms, p = 1006, 5^3 nodes)."""
import numpy as np
import pytest

from helpers import relerr, tmbdata_from_oracle

pytestmark = pytest.mark.gpu


def three_term_model(k, n, family, seed):
    """C5's generator (scripts/run_config.py) at n observations: three IWP3 terms with k knots each + intercept,
    p = 3 (k + 1) + 1."""
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(seed)
    xs = [rng.uniform(0, 1, n) for _ in range(3)]
    eta = 0.5 + np.sin(2 * np.pi * xs[0]) + 0.4 * np.sin(3 * np.pi * xs[1]) + 0.6 * np.sin(2.5 * np.pi * xs[2] + 1.0)
    terms = [Term("IWP", "x%d" % (i + 1), xs[i], order=3, k=k) for i in range(3)]
    if family == "Poisson":
        y = rng.poisson(np.exp(eta)).astype(np.float64)
        return build_model(y, terms, {}, family="Poisson")[0]
    size = 1.0 + rng.poisson(9, n)
    y = rng.binomial(size.astype(int), 1 / (1 + np.exp(-(eta - 1.0)))).astype(np.float64)
    return build_model(y, terms, {}, family="Binomial", size=size)[0]


def compare_with_oracle(model, thetas):
    from bayesgp_b200 import make_objective
    from oracle.laplace import LaplaceObjective as OFF
    off = OFF(model)
    ff = make_objective(tmbdata_from_oracle(model))
    try:
        assert ff.p == model.p
        rng = np.random.default_rng(5)
        W = 0.02 * rng.standard_normal(model.p)
        o = model.objective(W, thetas[0], "fgH")
        f, g, H = ff.objective(W, thetas[0], want_grad=True, want_hess=True)
        assert abs(f - o["f"]) <= 1e-11 * abs(o["f"])                      # src/BayesGP.cpp:133-168,219-252
        assert relerr(g, o["g"]) < 1e-10 and relerr(H, o["H"]) < 1e-10
        assert np.array_equal(H, H.T)
        for th in thetas:
            want = off.fn(th)
            got, _, w, Hm = ff._eval(th, want_hess=True)
            assert np.isfinite(want)
            assert abs(got - want) <= 1e-8 * abs(want), (th, got, want)       # north_star: 1e-8 relative
            assert relerr(w, off.last_par) < 1e-6 and relerr(Hm, off.sp_hess()) < 1e-6
            gw, gg = off.gr(th), ff.gr(th)
            assert np.max(np.abs(gw - gg)) <= 2e-7 * max(1.0, np.max(np.abs(gw))), (th, gw, gg, getattr(ff, "last_warning", None))
        # the batch entry point (what the quadrature grid runs through) at the same latent size
        ff.set_start(None)
        vals, modes, Hs, _ = ff.fn_batch(np.array(thetas), want_modes=True, want_hess=True)
        for j, th in enumerate(thetas):
            want = off.fn(th)
            assert abs(vals[j] - want) <= 1e-8 * abs(want)
            assert relerr(modes[j], off.last_par) < 1e-6 and relerr(Hs[j], off.sp_hess()) < 1e-6
    finally:
        ff.close()


@pytest.mark.parametrize("k,family", [(199, "Poisson"), (265, "Poisson"), (334, "Poisson"), (334, "Binomial"),
                                      (232, "Binomial")])
def test_three_terms_above_512_columns(k, family):
    """p = 601, 799, 1006 (Poisson) and 1006, 700 (Binomial): 10, 13, 16, 16 and 11 column groups."""
    model = three_term_model(k, 20000, family, seed=900 + k)
    assert model.p == 3 * (k + 1) + 1 and model.p > 512
    compare_with_oracle(model, [np.array([-3.0, -3.5, -4.0]), np.array([-3.4, -3.2, -4.3])])


def test_gaussian_above_512_columns():
    """Gaussian family (noise theta last, sumsq path of the likelihood pass, c3 = 0) at p = 601, S = 4."""
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(77)
    n, k = 20000, 199
    xs = [rng.uniform(0, 1, n) for _ in range(3)]
    eta = 0.5 + np.sin(2 * np.pi * xs[0]) + 0.4 * np.sin(3 * np.pi * xs[1]) + 0.6 * np.sin(2.5 * np.pi * xs[2] + 1.0)
    y = eta + 0.3 * rng.standard_normal(n)
    model = build_model(y, [Term("IWP", "x%d" % (i + 1), xs[i], order=3, k=k) for i in range(3)], {}, family="Gaussian")[0]
    assert model.p == 601 and model.S == 4
    compare_with_oracle(model, [np.array([-3.0, -3.5, -4.0, 2.0]), np.array([-3.4, -3.2, -4.3, 2.4])])


def _fit_both(oargs, pargs, k, mode, hessian):
    """oracle and product AGHQ objects on the same grid centre / scale (aghq's `optresults` argument)."""
    import bayesgp_b200 as bg
    from oracle import fit as ofit
    from oracle.aghq import marginal_laplace_tmb as o_mlt
    from oracle.laplace import LaplaceObjective as OFF
    model = ofit.build_model(oargs["y"], oargs["terms"], {}, oargs["family"], oargs.get("size"))[0]
    off = OFF(model)
    opt = {"mode": np.asarray(mode, dtype=np.float64), "hessian": np.asarray(hessian, dtype=np.float64)}
    omod = o_mlt(off, k, np.zeros(model.S), mode=opt["mode"], hessian=opt["hessian"])
    pfit = bg.model_fit(pargs["y"], pargs["terms"], {}, family=pargs["family"], size=pargs.get("size"), aghq_k=k, M=0,
                        optresults=opt)
    return model, omod, pfit


def _assert_fit_parity(model, omod, pfit):
    mod = pfit.mod
    nw = mod.normalized_posterior["nodesandweights"]
    assert mod.p == model.p and mod.K == len(omod.weights)
    assert relerr(nw["theta"], omod.nodes) < 1e-12 and relerr(nw["weights"], omod.weights) < 1e-12
    assert abs(mod.lognormconst - omod.lognormconst) <= 1e-8 * abs(omod.lognormconst)            # north_star
    assert np.max(np.abs(nw["logpost"] - omod.logpost)) <= 1e-8 * np.max(np.abs(omod.logpost))
    mh = mod.modesandhessians
    for j in range(mod.K):
        assert relerr(mh["mode"][j], omod.modes[j]) < 1e-6
        assert relerr(mh["H"][j], omod.hessians[j]) < 1e-6
    for j in range(mod.S):
        assert relerr(mod.marginals[j]["theta"], omod.marginals[j]["theta"]) < 1e-10
        assert np.max(np.abs(mod.marginals[j]["logmargpost"] - omod.marginals[j]["logmargpost"])) \
            < max(1e-5, 1e-8 * abs(omod.lognormconst))


def test_c4_shape_full_latent_size_and_grid():
    """BASELINE C4 as scripts/run_config.py builds it (Binomial, IWP2 k = 440 + sGP k = 20, p = 497), the full 7^2
    grid and its marginals, n = 6000."""
    import bayesgp_b200 as bg
    from oracle import fit as ofit
    rng = np.random.default_rng(20244)
    n = 6000
    x1, x2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    eta = -0.3 + np.sin(2 * np.pi * x1) + 0.6 * np.cos(2 * np.pi * 5 * x2)
    size = 1.0 + rng.poisson(9, n)
    y = rng.binomial(size.astype(int), 1 / (1 + np.exp(-eta))).astype(np.float64)
    mk = lambda T: [T("IWP", "x1", x1, order=2, k=440),
                    T("sGP", "x2", x2, a=2 * np.pi * 5, k=20, m=1, region=np.array([0.0, 1.0]), accuracy=0.01)]
    oargs = dict(y=y, terms=mk(ofit.Term), family="Binomial", size=size)
    pargs = dict(y=y, terms=mk(bg.Term), family="Binomial", size=size)
    model, omod, pfit = _fit_both(oargs, pargs, 7, [2.0, -1.0], [[3.0, 0.4], [0.4, 1.5]])
    try:
        assert model.p == 497 and pfit.mod.K == 49
        _assert_fit_parity(model, omod, pfit)
    finally:
        pfit.close()


def test_c5_shape_full_latent_size_and_grid():
    """BASELINE C5's latent structure (Poisson, three IWP3 k = 334 terms, p = 1006, S = 3) and its full 5^3 = 125 node
    grid, n = 2000.  The oracle evaluates the main grid (125 Laplace evaluations, about a minute of CPU); the product
    runs the whole marginal_laplace_tmb (main grid + the two re-ordered marginal grids = 375 evaluations).  Compared:
    nodes, weights, every node's log posterior, the normalising constant, the first marginal (a function of the
    main grid only), modes and Hessians at six nodes (corners, centre)."""
    import bayesgp_b200 as bg
    from oracle import fit as ofit
    from oracle.aghq import gh_rule, normalize_logpost
    from oracle.laplace import LaplaceObjective as OFF
    from scipy.special import logsumexp
    rng = np.random.default_rng(20245)
    n = 2000
    xs = [rng.uniform(0, 1, n) for _ in range(3)]
    eta = 0.5 + np.sin(2 * np.pi * xs[0]) + 0.4 * np.sin(3 * np.pi * xs[1]) + 0.6 * np.sin(2.5 * np.pi * xs[2] + 1.0)
    y = rng.poisson(np.exp(eta)).astype(np.float64)
    mk = lambda T: [T("IWP", "x%d" % (i + 1), xs[i], order=3, k=334) for i in range(3)]
    mode = np.array([-3.0, -2.0, -2.5])
    hess = np.array([[2.0, 0.1, 0.05], [0.1, 1.5, 0.1], [0.05, 0.1, 1.8]])
    k = 5
    model = ofit.build_model(y, mk(ofit.Term), {}, "Poisson")[0]
    off = OFF(model)
    nodes, weights, logpost, lnc, L = normalize_logpost(off, mode, hess, k)
    pfit = bg.model_fit(y, mk(bg.Term), {}, family="Poisson", aghq_k=k, M=0, optresults={"mode": mode, "hessian": hess})
    try:
        mod = pfit.mod
        assert model.p == 1006 and mod.p == 1006 and mod.K == 125 and mod.S == 3
        nw = mod.normalized_posterior["nodesandweights"]
        assert relerr(nw["theta"], nodes) < 1e-12 and relerr(nw["weights"], weights) < 1e-12
        assert np.max(np.abs(nw["logpost"] - logpost)) <= 1e-8 * np.max(np.abs(logpost))
        assert abs(mod.lognormconst - lnc) <= 1e-8 * abs(lnc)                                     # north_star
        z1, w1 = gh_rule(k)
        lpn = logpost - lnc
        want = np.array([logsumexp((lpn + np.log(weights))[np.arange(125) % k == q]) - np.log(w1[q] * L[0, 0])
                         for q in range(k)])
        assert relerr(mod.marginals[0]["theta"], mode[0] + L[0, 0] * z1) < 1e-10
        assert np.max(np.abs(mod.marginals[0]["logmargpost"] - want)) < max(1e-5, 1e-8 * abs(lnc))
        for j in range(1, 3):                       # the re-ordered grids: finite, and they integrate to one
            mj = mod.marginals[j]
            assert np.all(np.isfinite(mj["logmargpost"]))
            assert abs(np.sum(np.exp(mj["logmargpost"]) * mj["w"]) - 1.0) < 1e-8
        mh = mod.modesandhessians
        for j in (0, 4, 62, 100, 120, 124):
            off.fn(nodes[j])
            assert relerr(mh["mode"][j], off.last_par) < 1e-6
            assert relerr(mh["H"][j], off.sp_hess()) < 1e-6
    finally:
        pfit.close()
