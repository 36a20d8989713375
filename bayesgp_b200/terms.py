"""Host-side term setup mirroring the R layer above the drop-in boundary
(``/root/reference/R/02_model_fit.R:358-616``: defaults, knots, initial_location, priors) and,
for terms whose design cannot be generated on the device yet (sGP, IID), the dense blocks the R
layer would hand to ``get_result_by_method`` (``/root/reference/R/01_utility.R:67-272``).

This is setup code (runs once per fit, like the R constructors it mirrors), written with
vectorised numpy.  IWP terms carry no host matrices at all: their B / X / P are built on the
GPU from the covariate (``bgp_model_add_iwp``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


@dataclass
class Term:
    """One ``f(smoothing_var, model = ...)`` term (R/01_utility.R:3-15, S4 slots :34-56)."""
    kind: str                       # "IWP" | "sGP" | "IID"
    name: str
    x: np.ndarray
    order: int = 0
    knots: Optional[np.ndarray] = None
    k: Optional[int] = None
    initial_location: Optional[float] = None
    a: float = 0.0
    m: int = 1
    region: Optional[np.ndarray] = None
    accuracy: float = 0.01
    boundary: bool = True
    u: float = 1.0                  # sd.prior$param defaults (R/02_model_fit.R:377)
    alpha: float = 0.5
    boundary_prec: float = 0.01     # boundary.prior defaults (R/02_model_fit.R:444-452)
    boundary_mean: float = 0.0
    observed_x: np.ndarray = field(default=None, repr=False)
    n_basis: int = 0
    n_boundary: int = 0


def prepare_term(t: Term) -> Term:
    x = np.asarray(t.x, dtype=np.float64)
    if t.kind == "IWP":
        if t.k is not None and t.k < 3:
            raise ValueError("Error: parameter <k> in the random effect part should be >= 3.")
        if t.order is None or t.order < 1:
            raise ValueError("Error: Parameter <order> in the random effect part should be >= 1.")
        if t.initial_location is None:
            t.initial_location = float(x.min())
        xi = x - t.initial_location
        if t.knots is None:
            t.knots = np.unique(np.sort(np.linspace(xi.min(), xi.max(), 5 if t.k is None else t.k)))
        t.knots = np.asarray(t.knots, dtype=np.float64)
        t.observed_x = np.sort(xi)
        kn = t.knots
        nneg = len(np.unique(np.where(kn < 0, -kn, 0.0))) - 1 if kn.min() < 0 else 0
        npos = (len(kn) - 1) if kn.min() >= 0 else (len(np.unique(np.where(kn > 0, kn, 0.0))) - 1 if kn.max() > 0 else 0)
        t.n_basis = nneg + npos
        t.n_boundary = t.order - 1
    elif t.kind == "sGP":
        if t.k is None:
            t.k = 30
        if t.k < 3:
            raise ValueError("Error: parameter <k> in the random effect part should be >= 3.")
        if t.k < 4:         # accepted by R/02_model_fit.R:511-514, refused by fda inside Compute_B_sB (R/01_utility.R:179-183)
            raise ValueError("sGP with k = 3: fda::create.bspline.basis needs nbasis >= norder = 4")
        if t.a < 0:
            raise ValueError("Error: Parameter <a> in the random effect part should be positive.")
        if t.initial_location is None:
            t.initial_location = float(x.min())
        xi = x - t.initial_location
        t.observed_x = np.sort(xi)
        if t.region is None:
            t.region = np.array([t.observed_x[0], t.observed_x[-1]])
        t.region = np.asarray(t.region, dtype=np.float64)
        t.n_basis = 3 * (t.k - 2) * t.m
        t.n_boundary = 2 * t.m
    elif t.kind == "IID":
        t.n_basis = len(np.unique(x))
        t.n_boundary = 0
    else:
        raise ValueError("unknown model class %r" % t.kind)
    return t


# ---- dense blocks for the terms that are not generated on the device ------------------------------------
# (IWP and sGP designs, the IWP precision and the sGP precision Compute_Q_sB are all built on the device:
#  bgp_model_add_iwp / bgp_model_add_sgp_auto / bgp_sgp_precision, csrc/basis.cu)
def iid_design(t: Term):
    lev, inv = np.unique(np.asarray(t.x), return_inverse=True)
    B = np.zeros((len(inv), len(lev)))
    B[np.arange(len(inv)), inv] = 1.0
    return B, np.ones(len(lev))
