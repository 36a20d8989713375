"""CPU: posterior summaries.  The restatement of aghq::compute_pdf_and_cdf /
compute_quantiles is pinned on the theta quantiles the README prints (/root/reference/README.md:83-85); the
host-side functions of the product (bayesgp_b200/post_fit.py, pure numpy on k-point tables) are checked against the
oracle's on the same tables."""
import json
import math
import os

import numpy as np
import pytest

from helpers import GOLDEN, covid_terms


@pytest.fixture(scope="module")
def readme():
    with open(os.path.join(GOLDEN, "readme_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def readme_mod(readme):
    from oracle.aghq import marginal_laplace_tmb
    from oracle.fit import build_model
    from oracle.laplace import LaplaceObjective
    y, terms, fixed = covid_terms()
    model = build_model(y, terms, fixed, family="Poisson")[0]
    return marginal_laplace_tmb(LaplaceObjective(model), 4, [0.0], mode=np.array([readme["theta_mode"]]),
                                hessian=np.array([[1.0 / readme["quad_cov"]]]))


def test_theta_quantiles_match_the_readme_printout(readme, readme_mod):
    from oracle.summary import theta_summary_table
    row = theta_summary_table(readme_mod)[0]
    # the quantiles are points of a 1000-point grid (spacing 2.6e-3): agreeing to the printed 7 digits means the
    # same interpolant, the same grid and the same cdf rule; the README's centre / scale carry 7 digits themselves
    for key in ("2.5%", "median", "97.5%"):
        assert abs(row[key] - readme["theta_quantiles"][key]) < 2e-6, (key, row[key])
    assert abs(row["mean"] - readme["theta_mean"]) < 5e-6 and abs(row["sd"] - readme["theta_sd"]) < 5e-6


def test_natural_spline_against_scipy_and_linear_outside():
    from scipy.interpolate import CubicSpline
    from oracle.summary import natural_spline
    from bayesgp_b200.post_fit import _natural_spline
    rng = np.random.default_rng(1)
    for k in (3, 4, 7, 15):
        x = np.sort(rng.uniform(-3, 2, k))
        y = rng.standard_normal(k)
        ref = CubicSpline(x, y, bc_type="natural")
        xs = np.linspace(x[0], x[-1], 101)
        for f in (natural_spline(x, y), _natural_spline(x, y)):
            assert np.max(np.abs(f(xs) - ref(xs))) < 1e-12
            assert np.allclose(f(x), y, atol=1e-13)
            out = np.array([x[0] - 2.0, x[0] - 1.0, x[0] - 0.5])           # straight line outside the nodes
            v = f(out)
            assert abs((v[1] - v[0]) - 2.0 * (v[2] - v[1])) < 1e-12
            assert abs((v[2] - v[1]) / 0.5 - ref(x[0], 1)) < 1e-10
    with pytest.raises(ValueError):
        natural_spline([0.0, 1.0], [0.0, 1.0])
    with pytest.raises(ValueError):
        _natural_spline([0.0, 1.0], [0.0, 1.0])


def test_product_host_functions_equal_the_oracle_on_the_same_tables(readme_mod):
    from oracle import summary as osum
    from bayesgp_b200 import post_fit as psum
    rng = np.random.default_rng(2)
    tables = [readme_mod.marginals[0]]
    for k in (3, 5, 9):
        th = np.sort(rng.uniform(-6, 1, k))
        tables.append({"theta": th, "logmargpost": -0.5 * ((th + 2.5) / 0.7) ** 2 + 0.1 * rng.standard_normal(k)})
    for tab in tables:
        a, b = osum.compute_pdf_and_cdf(tab, to_sd=True), psum.compute_pdf_and_cdf(tab, transformation="sd")
        for key in ("theta", "pdf", "cdf", "transparam", "pdf_transparam"):
            assert np.allclose(a[key], b[key], rtol=1e-11, atol=0)
        q = (0.025, 0.25, 0.5, 0.975)
        assert np.array_equal(osum.compute_quantiles(tab, q), psum.compute_quantiles(tab, q))


def test_var_density_scales_and_priors(readme_mod):
    from oracle.summary import psd_correction, theta_logprior, var_density
    from bayesgp_b200.post_fit import compute_d_step_sGPsd
    marg = readme_mod.marginals[0]
    vd = var_density(marg, alpha=0.5, u=1.0, kind="IWP", h=2.0, order=3)
    assert np.all(np.diff(vd["SD"]) > 0)
    # densities on the SD scale integrate like the theta-scale cdf does (change of variables), prior is a density
    tot = np.sum(0.5 * (vd["post"][1:] + vd["post"][:-1]) * np.diff(vd["SD"]))
    assert abs(tot - 1.0) < 5e-3
    lam = -math.log(0.5) / 1.0
    assert np.allclose(vd["prior"], lam * np.exp(-lam * vd["SD"]), rtol=1e-12)        # Exponential(lambda) on sigma
    c = math.sqrt(2.0 ** 5 / (5 * math.factorial(2) ** 2))                            # R/03_post_fit.R:352-355
    assert abs(psd_correction("IWP", 2.0, order=3) - c) < 1e-15
    assert np.allclose(vd["PSD"], vd["SD"] * c) and np.allclose(vd["post.PSD"], vd["post"] / c)
    a = 2 * math.pi
    want = sum(math.sqrt((1 / (j * a) ** 2) * (0.3 / 2 - math.sin(2 * j * a * 0.3) / (4 * j * a))) for j in (1, 2))
    assert abs(psd_correction("sGP", 0.3, a=a, m=2) - want) < 1e-15
    assert abs(compute_d_step_sGPsd(0.3, a) + compute_d_step_sGPsd(0.3, 2 * a) - want) < 1e-15
    assert abs(theta_logprior(0.0, 0.5, 1.0) - (math.log(lam / 2) - lam)) < 1e-15


def test_fixed_effect_table():
    from oracle.summary import fixed_effect_summary
    rng = np.random.default_rng(3)
    rows = rng.standard_normal((3, 3000)) * np.array([[1.0], [0.1], [5.0]]) + np.array([[0.0], [2.0], [-1.0]])
    tab = fixed_effect_summary(rows)
    assert np.allclose(tab["Median"], np.median(rows, axis=1))
    assert np.allclose(tab["1st Qu."], np.percentile(rows, 25, axis=1))
    assert np.allclose(tab["sd"], rows.std(axis=1, ddof=1))
