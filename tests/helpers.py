"""Shared test helpers: seeded synthetic inputs and oracle <-> product marshalling."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def covid_terms():
    from oracle.fit import Term
    cc = np.load(os.path.join(GOLDEN, "covid_canada.npz"))
    fixed = {f"weekdays{i}": cc[f"weekdays{i}"] for i in range(1, 7)}
    return cc["new_deaths"], [Term("IWP", "t", cc["t"], order=3, k=30)], fixed


def covid_model():
    from oracle.fit import build_model
    y, terms, fixed = covid_terms()
    return build_model(y, terms, fixed, family="Poisson")


def synth_poisson(n=20000, k=40, order=3, seed=20243):
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, n)
    eta = 1.0 + np.sin(2 * np.pi * x) + 0.5 * np.cos(6 * np.pi * x)
    y = rng.poisson(np.exp(eta)).astype(np.float64)
    return build_model(y, [Term("IWP", "x", x, order=order, k=k)], {}, family="Poisson")


def synth_binomial_sgp(n=8000, seed=20244):
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(seed)
    x1 = rng.uniform(0, 1, n)
    x2 = rng.uniform(0, 1, n)
    eta = -0.3 + np.sin(2 * np.pi * x1) + 0.6 * np.cos(2 * np.pi * 5 * x2)
    size = 1.0 + rng.poisson(9, n)
    y = rng.binomial(size.astype(int), 1 / (1 + np.exp(-eta))).astype(np.float64)
    terms = [Term("IWP", "x1", x1, order=2, k=20),
             Term("sGP", "x2", x2, a=2 * np.pi * 5, k=12, m=1, region=np.array([0.0, 1.0]), accuracy=0.01)]
    return build_model(y, terms, {}, family="Binomial", size=size)


def synth_gaussian(n=5000, seed=20242):
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, n)
    z = rng.standard_normal(n)
    y = 2.0 * x + np.sin(4 * x) + 0.3 * z + 0.5 * rng.standard_normal(n)
    return build_model(y, [Term("IWP", "x", x, order=2, k=25)], {"z": z}, family="Gaussian")


def tmbdata_from_oracle(model):
    """oracle.model.Model -> bayesgp_b200.TMBData (same numbers, R column-major layout)."""
    from bayesgp_b200 import TMBData
    fam = {0: "Gaussian", 1: "Poisson", 2: "Binomial", -2: "none"}[model.family]
    return TMBData(model.y, fam, model.B, model.P, model.logPdet, model.u, model.alpha, model.X, model.betaprec,
                   model.betamean, model.Xf, model.beta_fixed_prec, model.beta_fixed_mean, model.size)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))
