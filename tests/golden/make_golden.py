"""Regenerate the committed fixtures under tests/golden/ (run in the BUILD container,
where /root/reference is mounted; the GPU box never reads /root/reference).

  covid_canada.npz / sim1data.npz : the two datasets the reference ships
      (/root/reference/data/*.rda, decoded with oracle/rdata.py because R is absent).
  readme_golden.json              : the known-answer values printed in
      /root/reference/README.md:71-96 (the only goldens the reference has on this path).
  oracle_covid.npz                : oracle outputs on the README model, used by the GPU
      parity tests as a frozen cross-check of the live oracle.

R, TMB and aghq are not installed here, so no reference-*generated* vectors can be made;
see oracle/__init__.py ("Parity pin").
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.rdata import read_rda  # noqa: E402

REF = "/root/reference"


def main():
    cc = read_rda(os.path.join(REF, "data/covid_canada.rda"))["covid_canada"]
    np.savez_compressed(os.path.join(HERE, "covid_canada.npz"), **{k: np.asarray(v) for k, v in cc.items()})
    s1 = read_rda(os.path.join(REF, "data/sim1data.rda"))["sim1data"]
    np.savez_compressed(os.path.join(HERE, "sim1data.npz"), **{k: np.asarray(v) for k, v in s1.items()})
    readme = {
        "source": "/root/reference/README.md:71-96",
        "call": "model_fit(new_deaths ~ weekdays1..6 + f(t, model='IWP', order=3, k=30), covid_canada, method='aghq', family='Poisson')",
        "latent_dim": 38,
        "aghq_k": 4,
        "theta_mode": -3.245926,
        "lognormconst": -4322.531,
        "quad_cov": 0.07936619,
        "theta_mean": -3.271182,
        "theta_sd": 0.2785344,
        "fixed": {
            "intercept": {"mean": -5.40444709, "sd": 0.66061232},
            "weekdays1": {"mean": 0.09374558, "sd": 0.01198239},
            "weekdays2": {"mean": 0.07921671, "sd": 0.01188838},
            "weekdays3": {"mean": 0.12672077, "sd": 0.01150235},
            "weekdays4": {"mean": 0.12547251, "sd": 0.01181344},
            "weekdays5": {"mean": 0.05001256, "sd": 0.01213118},
            "weekdays6": {"mean": -0.15125835, "sd": 0.01336132},
        },
        "M": 3000,
    }
    with open(os.path.join(HERE, "readme_golden.json"), "w") as fh:
        json.dump(readme, fh, indent=1)

    # frozen oracle outputs on the README model
    from oracle.fit import Term, model_fit
    fixed = {f"weekdays{i}": cc[f"weekdays{i}"] for i in range(1, 7)}
    rng = np.random.default_rng(20241)
    fit = model_fit(cc["new_deaths"], [Term("IWP", "t", cc["t"], order=3, k=30)], fixed, family="Poisson",
                    aghq_k=4, M=64, rng=rng)
    np.savez_compressed(
        os.path.join(HERE, "oracle_covid.npz"),
        theta_mode=fit.mod.mode, theta_hessian=fit.mod.hessian, nodes=fit.mod.nodes, weights=fit.mod.weights,
        logpost=fit.mod.logpost, lognormconst=fit.mod.lognormconst, modes=fit.mod.modes,
        hess_diag=np.stack([np.diag(h) for h in fit.mod.hessians]),
    )
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
