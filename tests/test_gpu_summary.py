"""GPU: summary() and var_density() of a fit produced by the CUDA path, against the README
printout (/root/reference/README.md:75-96) and against the oracle's summaries of the oracle's own fit."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, covid_terms

pytestmark = pytest.mark.gpu


def _covid():
    cc = np.load(os.path.join(GOLDEN, "covid_canada.npz"))
    return cc["new_deaths"], cc["t"], {f"weekdays{i}": cc[f"weekdays{i}"] for i in range(1, 7)}


@pytest.fixture(scope="module")
def readme():
    with open(os.path.join(GOLDEN, "readme_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def fit(readme):
    import bayesgp_b200 as bg
    y, t, fixed = _covid()
    f = bg.model_fit(y, [bg.Term("IWP", "t", t, order=3, k=30)], fixed, family="Poisson", aghq_k=4, M=3000, seed=11,
                     optresults={"mode": np.array([readme["theta_mode"]]),
                                 "hessian": np.array([[1.0 / readme["quad_cov"]]])})
    yield f
    f.close()


def test_summary_reproduces_the_readme_tables(readme, fit, capsys):
    import bayesgp_b200 as bg
    out = bg.summary(fit)
    printed = capsys.readouterr().out
    assert "AGHQ on a 1 dimensional posterior with  4 quadrature points" in printed
    assert "Here are some moments and quantiles for the log precision" in printed
    row = out["theta"]["theta(t)"]
    assert abs(row["mean"] - readme["theta_mean"]) < 5e-6 and abs(row["sd"] - readme["theta_sd"]) < 5e-6
    for key in ("2.5%", "median", "97.5%"):
        assert abs(row[key] - readme["theta_quantiles"][key]) < 2e-6, (key, row[key])
    # fixed effects: sample quartiles within Monte-Carlo error of 3000 draws (different random draws than R's):
    # se of a sample quantile = sqrt(q (1 - q) / M) / density, density of a normal at its quartile = 0.3178 / sd
    for name, g in readme["fixed_quartiles"].items():
        r = out["fixed"][name]
        sd = readme["fixed"][name]["sd"]
        se_q = np.sqrt(0.25 * 0.75 / readme["M"]) / (0.3178 / sd)
        se_m = np.sqrt(0.25 / readme["M"]) / (0.3989 / sd)
        assert abs(r["1st Qu."] - g["q1"]) < 6 * se_q * np.sqrt(2), name
        assert abs(r["3rd Qu."] - g["q3"]) < 6 * se_q * np.sqrt(2), name
        assert abs(r["Median"] - g["median"]) < 6 * se_m * np.sqrt(2), name
        assert abs(r["Mean"] - readme["fixed"][name]["mean"]) < 6 * sd / np.sqrt(readme["M"]) * np.sqrt(2), name
        assert abs(r["sd"] / sd - 1.0) < 0.08, name


def test_var_density_matches_the_oracle(readme, fit):
    import bayesgp_b200 as bg
    from oracle import summary as osum
    from oracle.aghq import marginal_laplace_tmb
    from oracle.fit import build_model
    from oracle.laplace import LaplaceObjective
    y, terms, fixed = covid_terms()
    model = build_model(y, terms, fixed, family="Poisson")[0]
    omod = marginal_laplace_tmb(LaplaceObjective(model), 4, [0.0], mode=np.array([readme["theta_mode"]]),
                                hessian=np.array([[1.0 / readme["quad_cov"]]]))
    want = osum.var_density(omod.marginals[0], alpha=0.5, u=1.0, kind="IWP", h=1.5, order=3)
    got = bg.var_density(fit, component="t", h=1.5)
    assert set(got) == {"SD", "post", "prior", "PSD", "post.PSD", "prior.PSD"}
    for key in got:
        # the log marginal carries the ~1e-5 absolute noise of the ill-conditioned README model
        assert np.allclose(got[key], want[key], rtol=1e-4, atol=1e-12), key
    with pytest.raises(ValueError):
        bg.var_density(fit)                      # Poisson: no family SD
    with pytest.raises(ValueError):
        bg.var_density(fit, component="nope")


def test_gaussian_family_sd_density():
    import bayesgp_b200 as bg
    rng = np.random.default_rng(5)
    n = 1500
    x = rng.uniform(0, 1, n)
    y = np.sin(4 * x) + 0.3 * rng.standard_normal(n)
    fit = bg.model_fit(y, [bg.Term("IWP", "x", x, order=2, k=15)], None, family="Gaussian", aghq_k=5, M=200, seed=1,
                       control_family={"u": 1.0, "alpha": 0.5})
    try:
        s = bg.summary(fit, echo=False)
        assert list(s["theta"]) == ["theta(x)", "theta(family)"]
        vd = bg.var_density(fit)                 # family SD
        peak = vd["SD"][np.argmax(vd["post"])]
        assert 0.25 < peak < 0.36                # the noise SD the data were generated with is 0.3
        tot = np.sum(0.5 * (vd["post"][1:] + vd["post"][:-1]) * np.diff(vd["SD"]))
        assert abs(tot - 1.0) < 2e-2
    finally:
        fit.close()
