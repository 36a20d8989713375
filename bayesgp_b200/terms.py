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
def _bspline_all(x, lo, hi, k, deriv=0):
    """Cubic B-spline basis of fda::create.bspline.basis(c(lo, hi), nbasis = k, norder = 4) and its
    derivatives, vectorised Cox-de Boor on the clamped knot vector; (len(x), k)."""
    norder = 4
    breaks = np.linspace(lo, hi, k - norder + 2)
    t = np.concatenate([np.full(norder - 1, lo), breaks, np.full(norder - 1, hi)])
    nt = len(t)
    x = np.asarray(x, dtype=np.float64)
    B = np.zeros((len(x), nt - 1))
    last = np.max(np.nonzero(t[1:] > t[:-1])[0])
    for j in range(nt - 1):
        if t[j + 1] > t[j]:
            B[:, j] = (x >= t[j]) & ((x < t[j + 1]) | ((j == last) & (x <= t[j + 1])))
    for m in range(2, norder - deriv + 1):
        Bn = np.zeros((len(x), nt - m))
        for j in range(nt - m):
            d1, d2 = t[j + m - 1] - t[j], t[j + m] - t[j + 1]
            if d1 > 0:
                Bn[:, j] += (x - t[j]) / d1 * B[:, j]
            if d2 > 0:
                Bn[:, j] += (t[j + m] - x) / d2 * B[:, j + 1]
        B = Bn
    for m in range(norder - deriv + 1, norder + 1):
        Bn = np.zeros((len(x), nt - m))
        for j in range(nt - m):
            d1, d2 = t[j + m - 1] - t[j], t[j + m] - t[j + 1]
            if d1 > 0:
                Bn[:, j] += (m - 1) / d1 * B[:, j]
            if d2 > 0:
                Bn[:, j] -= (m - 1) / d2 * B[:, j + 1]
        B = Bn
    return B * ((x >= lo) & (x <= hi))[:, None]


def sgp_design(t: Term):
    """B = cbind over harmonics of [B cos, B sin, B]; X = cbind(cos, sin) (R/01_utility.R:177-195,224-239,301-312).
    Fit-time B always drops the first two B-splines (boundary = TRUE is hard-wired there, A.8)."""
    xi = np.asarray(t.x, dtype=np.float64) - t.initial_location
    lo, hi = float(t.region.min()), float(t.region.max())
    Bm = _bspline_all(xi, lo, hi, t.k)[:, 2:]
    Bs, Xs = [], []
    for i in range(1, t.m + 1):
        c, s = np.cos(i * t.a * xi)[:, None], np.sin(i * t.a * xi)[:, None]
        Bs += [Bm * c, Bm * s, Bm]
        Xs += [c, s]
    return np.concatenate(Bs, axis=1), np.concatenate(Xs, axis=1)


def sgp_precision(t: Term):
    """Compute_Q_sB per harmonic, block-diagonal (R/01_utility.R:67-174,255-272)."""
    lo, hi = float(t.region.min()), float(t.region.max())
    nx = int(np.floor((hi - lo) / t.accuracy + 1e-10)) + 1
    x = lo + t.accuracy * np.arange(nx)
    B0 = _bspline_all(x, lo, hi, t.k, 0)[:, 2:]
    B1 = _bspline_all(x, lo, hi, t.k, 1)[:, 2:]
    B2 = _bspline_all(x, lo, hi, t.k, 2)[:, 2:]
    wI = np.diff(np.concatenate([[0.0], x]))[:, None]
    blocks = []
    for i in range(1, t.m + 1):
        a = i * t.a
        c, s = np.cos(a * x)[:, None], np.sin(a * x)[:, None]
        Bc, B1c, B2c, Bs, B1s, B2s = B0 * c, B1 * c, B2 * c, B0 * s, B1 * s, B2 * s

        def ip(U, V):
            return U.T @ (wI * V)

        def ss(Mx):
            return Mx + Mx.T

        T00, T10, T11, T20, T21, T22 = ip(Bc, Bc), ip(B1c, Bc), ip(B1c, B1c), ip(B2c, Bc), ip(B2c, B1c), ip(B2c, B2c)
        L00, L10, L11, L20, L21, L22 = ip(Bs, Bs), ip(B1s, Bs), ip(B1s, B1s), ip(B2s, Bs), ip(B2s, B1s), ip(B2s, B2s)
        I00, I10, I11, I20, I21, I22 = ip(Bs, Bc), ip(B1s, Bc), ip(B1s, B1c), ip(B2s, Bc), ip(B2s, B1c), ip(B2s, B2c)
        BB, B2B2, BB2 = ip(B0, B0), ip(B2, B2), ip(B0, B2)
        BS, BC, BS1, BC1, BS2, BC2 = ip(B0, Bs), ip(B0, Bc), ip(B0, B1s), ip(B0, B1c), ip(B0, B2s), ip(B0, B2c)
        B2S, B2C, B2S1, B2C1, B2S2, B2C2 = ip(B2, Bs), ip(B2, Bc), ip(B2, B1s), ip(B2, B1c), ip(B2, B2s), ip(B2, B2c)
        Gm = np.block([[T00, I00.T, BC.T], [I00, L00, BS.T], [BC, BS, BB]])
        C11 = T22 - 2 * a * ss(I21) - a ** 2 * ss(T20) + 2 * a ** 3 * ss(I10) + 4 * a ** 2 * L11 + a ** 4 * T00
        C22 = L22 + 2 * a * ss(I21) - a ** 2 * ss(L20) - 2 * a ** 3 * ss(I10) + 4 * a ** 2 * T11 + a ** 4 * L00
        C12 = (I22 + 2 * a * T21 - a ** 2 * ss(I20) - 2 * a * L21.T - 4 * a ** 2 * I11 + 2 * a ** 3 * L10
               - 2 * a ** 3 * T10.T + a ** 4 * I00)
        C13 = B2C2.T - 2 * a * B2S1.T - a ** 2 * B2C.T
        C23 = B2S2.T + 2 * a * B2C1.T - a ** 2 * B2S.T
        Cm = np.block([[C11, C12, C13], [C12.T, C22, C23], [C13.T, C23.T, B2B2]])
        M11 = T20.T - 2 * a * I10.T - a ** 2 * T00
        M12 = I20.T + 2 * a * T10.T - a ** 2 * I00
        M21 = I20.T - 2 * a * L10.T - a ** 2 * I00
        M22 = L20.T + 2 * a * I10.T - a ** 2 * L00
        M31 = BC2 - 2 * a * BS1 - a ** 2 * BC
        M32 = BS2 + 2 * a * BC1 - a ** 2 * BS
        Mm = np.block([[M11, M12, B2C.T], [M21, M22, B2S.T], [M31, M32, BB2]])
        Q = a ** 4 * Gm + Cm + a ** 2 * ss(Mm)
        blocks.append(np.triu(Q) + np.triu(Q, 1).T)
    n = sum(b.shape[0] for b in blocks)
    P = np.zeros((n, n))
    o = 0
    for b in blocks:
        P[o:o + len(b), o:o + len(b)] = b
        o += len(b)
    return P


def iid_design(t: Term):
    lev, inv = np.unique(np.asarray(t.x), return_inverse=True)
    B = np.zeros((len(inv), len(lev)))
    B[np.arange(len(inv)), inv] = 1.0
    return B, np.ones(len(lev))
