"""CPU: pin the oracle against the only known-answer values the reference ships for this path, the
rendered README fit (/root/reference/README.md:71-96, copied into tests/golden/readme_golden.json)."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, covid_terms


@pytest.fixture(scope="module")
def readme():
    with open(os.path.join(GOLDEN, "readme_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def covid_fit():
    from oracle.fit import model_fit
    y, terms, fixed = covid_terms()
    rng = np.random.default_rng(20241)
    return model_fit(y, terms, fixed, family="Poisson", aghq_k=4, M=3000, rng=rng)


def test_latent_dimension(readme, covid_fit):
    assert covid_fit.model.p == readme["latent_dim"] == 38
    assert covid_fit.model.S == 1


def test_lognormconst_matches_printed_value(readme, covid_fit):
    # README prints -4322.531 (7 significant digits)
    assert abs(covid_fit.mod.lognormconst - readme["lognormconst"]) < 1.5e-3


def test_mode_within_bfgs_stop_tolerance(readme, covid_fit):
    # same vmmin procedure from theta = 0: the README prints -3.245926, the oracle stops at -3.2459393
    assert abs(covid_fit.mod.mode[0] - readme["theta_mode"]) < 5e-5


def test_theta_moments_given_readme_grid(readme, covid_fit):
    """Using the README's own grid centre and scale, the oracle reproduces the printed lognormconst and the
    theta posterior mean / sd to printed precision."""
    from oracle.aghq import marginal_laplace_tmb, theta_moments
    from oracle.laplace import LaplaceObjective
    ff = LaplaceObjective(covid_fit.model)
    mod = marginal_laplace_tmb(ff, 4, [0.0], mode=np.array([readme["theta_mode"]]),
                               hessian=np.array([[1.0 / readme["quad_cov"]]]))
    assert abs(mod.lognormconst - readme["lognormconst"]) < 6e-4
    mean, sd = theta_moments(mod)
    assert abs(mean[0] - readme["theta_mean"]) < 5e-6
    assert abs(sd[0] - readme["theta_sd"]) < 5e-6


def test_fixed_effect_moments_within_monte_carlo_error(readme, covid_fit):
    from oracle.fit import sample_fixed_effect
    names = list(readme["fixed"])
    fe = sample_fixed_effect(covid_fit, names)
    for j, nm in enumerate(names):
        g = readme["fixed"][nm]
        se = g["sd"] / np.sqrt(readme["M"])
        assert abs(fe[:, j].mean() - g["mean"]) < 6 * se * np.sqrt(2), nm
        assert abs(fe[:, j].std(ddof=1) / g["sd"] - 1.0) < 0.08, nm


def test_frozen_oracle_outputs(covid_fit):
    """The committed oracle_covid.npz (made by tests/golden/make_golden.py) still matches the live oracle."""
    z = np.load(os.path.join(GOLDEN, "oracle_covid.npz"))
    assert np.allclose(z["theta_mode"], covid_fit.mod.mode, rtol=0, atol=1e-9)
    assert abs(float(z["lognormconst"]) - covid_fit.mod.lognormconst) < 1e-7
    assert np.allclose(z["modes"], covid_fit.mod.modes, rtol=1e-7, atol=1e-9)


def test_fixed_effect_quartiles_within_monte_carlo_error(readme, covid_fit):
    """The fixed-effect table of summary.FitResult (README.md:88-96: 1st Qu., Median, 3rd Qu.): sample quartiles
    of 3000 draws, so agreement is up to the Monte-Carlo error of a sample quantile,
    se = sqrt(q (1 - q) / M) / density."""
    from oracle.fit import sample_fixed_effect
    from oracle.summary import fixed_effect_summary
    names = list(readme["fixed_quartiles"])
    tab = fixed_effect_summary(sample_fixed_effect(covid_fit, names).T)
    for j, nm in enumerate(names):
        g, sd = readme["fixed_quartiles"][nm], readme["fixed"][nm]["sd"]
        se_q = np.sqrt(0.25 * 0.75 / readme["M"]) / (0.3178 / sd)
        se_m = np.sqrt(0.25 / readme["M"]) / (0.3989 / sd)
        assert abs(tab["1st Qu."][j] - g["q1"]) < 6 * se_q * np.sqrt(2), nm
        assert abs(tab["3rd Qu."][j] - g["q3"]) < 6 * se_q * np.sqrt(2), nm
        assert abs(tab["Median"][j] - g["median"]) < 6 * se_m * np.sqrt(2), nm
