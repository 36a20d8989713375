"""Basis / penalty constructors — ORACLE restatement (test infrastructure).

Follows ``/root/reference/R/01_utility.R``:
  * ``get_local_poly``            :346-364   O-spline design (IWP)
  * ``local_poly_helper``         :378-401   negative / positive knot split
  * ``global_poly_helper``        :413-419   1, x, x^2, ...
  * ``compute_weights_precision`` :325-344   diag(diff(knots))
  * ``Compute_B_sB``              :177-195   sGP design  [B cos, B sin, B]
  * ``Compute_B_sB_helper``       :198-208
  * ``global_poly_helper_sGP``    :430-440
  * ``Compute_Q_sB``              :67-174    sGP precision by Riemann sums
The cubic B-spline basis the sGP terms use comes from the un-vendored ``fda``
package (``create.bspline.basis(rangeval, nbasis=k, norder=4[, dropind=c(1,2)])``
+ ``eval.basis``): equally spaced breaks, ``k - 2`` of them, 4-fold boundary
knots, Cox-de Boor recursion, right-continuous except at the right end point
where the left limit is used.  Restated here in ``bspline_basis``.
"""
from __future__ import annotations

from math import factorial

import numpy as np


# ----------------------------------------------------------------------------
# IWP: O-splines
# ----------------------------------------------------------------------------
def get_local_poly_loop(knots, refined_x, p):
    """Literal double loop of R/01_utility.R:346-364 (small inputs only)."""
    knots = np.asarray(knots, dtype=np.float64)
    x = np.asarray(refined_x, dtype=np.float64)
    dif = np.diff(knots)
    nn, n = len(x), len(knots)
    D = np.zeros((nn, n - 1))
    for j in range(nn):
        for i in range(n - 1):
            if x[j] <= knots[i]:
                D[j, i] = 0.0
            elif x[j] <= knots[i + 1] and x[j] >= knots[i]:
                D[j, i] = (1.0 / factorial(p)) * (x[j] - knots[i]) ** p
            else:
                s = 0.0
                for k in range(1, p + 1):
                    s += (dif[i] ** k) * ((x[j] - knots[i + 1]) ** (p - k)) / (factorial(k) * factorial(p - k))
                D[j, i] = s
    return D


def get_local_poly(knots, refined_x, p):
    """Vectorised form of ``get_local_poly`` (same branch structure and the
    same summation order inside the tail polynomial)."""
    knots = np.asarray(knots, dtype=np.float64)
    xa = np.asarray(refined_x, dtype=np.float64)
    if xa.size > 65536:      # bound the temporaries: same arithmetic, row blocks
        out = np.empty((xa.size, len(knots) - 1))
        for s in range(0, xa.size, 65536):
            out[s:s + 65536] = get_local_poly(knots, xa[s:s + 65536], p)
        return out
    x = xa[:, None]
    k0 = knots[None, :-1]
    k1 = knots[None, 1:]
    dif = np.diff(knots)[None, :]
    inner = (1.0 / factorial(p)) * np.power(np.maximum(x - k0, 0.0), p)
    tail = np.zeros((x.shape[0], k0.shape[1]))
    xm = x - k1
    for k in range(1, p + 1):
        tail = tail + (dif ** k) * np.power(xm, p - k) / (factorial(k) * factorial(p - k))
    D = np.where(x <= k0, 0.0, np.where(x <= k1, inner, tail))
    return D


def _neg_knots(knots):
    knots = np.asarray(knots, dtype=np.float64)
    return np.unique(np.sort(np.where(knots < 0, -knots, 0.0)))


def _pos_knots(knots):
    knots = np.asarray(knots, dtype=np.float64)
    return np.unique(np.sort(np.where(knots > 0, knots, 0.0)))


def local_poly_helper(knots, refined_x, p=2, loop=False):
    """R/01_utility.R:378-401."""
    glp = get_local_poly_loop if loop else get_local_poly
    knots = np.asarray(knots, dtype=np.float64)
    x = np.asarray(refined_x, dtype=np.float64)
    if knots.min() >= 0:
        return glp(knots, x, p)
    if knots.max() <= 0:
        xn = np.where(x < 0, -x, 0.0)
        return glp(_neg_knots(knots), xn, p)
    xn = np.where(x < 0, -x, 0.0)
    D1 = glp(_neg_knots(knots), xn, p)
    xp = np.where(x > 0, x, 0.0)
    D2 = glp(_pos_knots(knots), xp, p)
    return np.concatenate([D1, D2], axis=1)


def global_poly_helper(x, p=2):
    """R/01_utility.R:413-419: columns x^0 .. x^(p-1)."""
    x = np.asarray(x, dtype=np.float64)
    return np.stack([x ** i for i in range(p)], axis=1)


def compute_weights_precision(knots):
    """R/01_utility.R:325-344 — returns the *diagonal* of the (diagonal) P."""
    knots = np.asarray(knots, dtype=np.float64)
    if knots.min() >= 0:
        return np.diff(knots)
    if knots.max() < 0:
        return np.diff(_neg_knots(knots))
    return np.concatenate([np.diff(_neg_knots(knots)), np.diff(_pos_knots(knots))])


def default_knots(x_init, k):
    """R/02_model_fit.R:435-442: unique(sort(seq(min, max, length.out=k)))."""
    return np.unique(np.sort(np.linspace(x_init.min(), x_init.max(), k)))


# ----------------------------------------------------------------------------
# sGP: seasonal B-spline basis
# ----------------------------------------------------------------------------
def bspline_knots(region, k, norder=4):
    lo, hi = float(np.min(region)), float(np.max(region))
    breaks = np.linspace(lo, hi, k - norder + 2)
    return np.concatenate([np.full(norder - 1, lo), breaks, np.full(norder - 1, hi)])


def bspline_basis(x, region, k, deriv=0, norder=4, drop_first_two=True):
    """fda::eval.basis(x, create.bspline.basis(range(region), nbasis=k, norder=4,
    dropind=c(1,2)), Lfdobj=deriv) — (len(x), k-2) (or k if not dropping)."""
    x = np.asarray(x, dtype=np.float64)
    t = bspline_knots(region, k, norder)
    nb = k
    lo, hi = t[0], t[-1]
    # order-1 (piecewise constant) functions on the non-empty spans;
    # the right end point belongs to the last non-empty span (left limit).
    nt = len(t)
    B = np.zeros((len(x), nt - 1))
    for j in range(nt - 1):
        if t[j + 1] > t[j]:
            if t[j + 1] == hi:
                B[:, j] = (x >= t[j]) & (x <= t[j + 1])
            else:
                B[:, j] = (x >= t[j]) & (x < t[j + 1])
    # raise to order `norder - deriv` by Cox-de Boor
    for m in range(2, norder - deriv + 1):
        Bn = np.zeros((len(x), nt - m))
        for j in range(nt - m):
            d1 = t[j + m - 1] - t[j]
            d2 = t[j + m] - t[j + 1]
            term = 0.0
            if d1 > 0:
                term = term + (x - t[j]) / d1 * B[:, j]
            if d2 > 0:
                term = term + (t[j + m] - x) / d2 * B[:, j + 1]
            Bn[:, j] = term
        B = Bn
    # apply the derivative recursion `deriv` times
    for m in range(norder - deriv + 1, norder + 1):
        Bn = np.zeros((len(x), nt - m))
        for j in range(nt - m):
            d1 = t[j + m - 1] - t[j]
            d2 = t[j + m] - t[j + 1]
            term = 0.0
            if d1 > 0:
                term = term + (m - 1) / d1 * B[:, j]
            if d2 > 0:
                term = term - (m - 1) / d2 * B[:, j + 1]
            Bn[:, j] = term
        B = Bn
    assert B.shape[1] == nb
    inside = (x >= lo) & (x <= hi)
    B = B * inside[:, None]
    return B[:, 2:] if drop_first_two else B


def compute_B_sB(x, a, k, region, boundary=True):
    """R/01_utility.R:177-195: cbind(B*cos(ax), B*sin(ax), B)."""
    x = np.asarray(x, dtype=np.float64)
    Bm = bspline_basis(x, region, k, 0, drop_first_two=boundary)
    c = np.cos(a * x)[:, None]
    s = np.sin(a * x)[:, None]
    return np.concatenate([Bm * c, Bm * s, Bm], axis=1)


def compute_B_sB_helper(refined_x, a, k, m, region, boundary=True, initial_location=None):
    """R/01_utility.R:198-208 (note: subtracts min(refined_x) when
    initial_location is NULL — predict passes NULL, R/03_post_fit.R:263)."""
    x = np.asarray(refined_x, dtype=np.float64)
    if initial_location is None:
        initial_location = x.min()
    x = x - initial_location
    return np.concatenate([compute_B_sB(x, (i * a), k, region, boundary) for i in range(1, m + 1)], axis=1)


def global_poly_sGP(x, a, m):
    """R/01_utility.R:301-312: cbind(cos(i a x), sin(i a x)) for i = 1..m."""
    x = np.asarray(x, dtype=np.float64)
    cols = []
    for i in range(1, m + 1):
        cols += [np.cos(i * a * x), np.sin(i * a * x)]
    return np.stack(cols, axis=1)


def global_poly_helper_sGP(refined_x, a, m, initial_location=None):
    """R/01_utility.R:430-440."""
    x = np.asarray(refined_x, dtype=np.float64)
    if initial_location is None:
        initial_location = x.min()
    return global_poly_sGP(x - initial_location, a, m)


def _seq_by(lo, hi, by):
    n = int(np.floor((hi - lo) / by + 1e-10)) + 1
    return lo + by * np.arange(n)


def compute_Q_sB(a, k, region, accuracy=0.01, boundary=True):
    """R/01_utility.R:67-174 — dense symmetric 3(k-2) x 3(k-2) precision."""
    lo, hi = float(np.min(region)), float(np.max(region))
    x = _seq_by(lo, hi, accuracy)
    B0 = bspline_basis(x, region, k, 0, drop_first_two=boundary)
    B1 = bspline_basis(x, region, k, 1, drop_first_two=boundary)
    B2 = bspline_basis(x, region, k, 2, drop_first_two=boundary)
    c = np.cos(a * x)[:, None]
    s = np.sin(a * x)[:, None]
    Bcos, B1cos, B2cos = B0 * c, B1 * c, B2 * c
    Bsin, B1sin, B2sin = B0 * s, B1 * s, B2 * s
    wI = np.diff(np.concatenate([[0.0], x]))[:, None]   # Numerical_I (:94)

    def ip(U, V):
        return U.T @ (wI * V)

    def ss(M):
        return M + M.T

    T00, T10, T11 = ip(Bcos, Bcos), ip(B1cos, Bcos), ip(B1cos, B1cos)
    T20, T21, T22 = ip(B2cos, Bcos), ip(B2cos, B1cos), ip(B2cos, B2cos)
    L00, L10, L11 = ip(Bsin, Bsin), ip(B1sin, Bsin), ip(B1sin, B1sin)
    L20, L21, L22 = ip(B2sin, Bsin), ip(B2sin, B1sin), ip(B2sin, B2sin)
    I00, I10, I11 = ip(Bsin, Bcos), ip(B1sin, Bcos), ip(B1sin, B1cos)
    I20, I21, I22 = ip(B2sin, Bcos), ip(B2sin, B1cos), ip(B2sin, B2cos)
    BB, B2B2, BB2 = ip(B0, B0), ip(B2, B2), ip(B0, B2)
    BS, BC = ip(B0, Bsin), ip(B0, Bcos)
    BS1, BC1 = ip(B0, B1sin), ip(B0, B1cos)
    BS2, BC2 = ip(B0, B2sin), ip(B0, B2cos)
    B2S, B2C = ip(B2, Bsin), ip(B2, Bcos)
    B2S1, B2C1 = ip(B2, B1sin), ip(B2, B1cos)
    B2S2, B2C2 = ip(B2, B2sin), ip(B2, B2cos)

    G = np.block([[T00, I00.T, BC.T], [I00, L00, BS.T], [BC, BS, BB]])
    C11 = T22 - 2 * a * ss(I21) - (a ** 2) * ss(T20) + 2 * (a ** 3) * ss(I10) + 4 * (a ** 2) * L11 + (a ** 4) * T00
    C22 = L22 + 2 * a * ss(I21) - (a ** 2) * ss(L20) - 2 * (a ** 3) * ss(I10) + 4 * (a ** 2) * T11 + (a ** 4) * L00
    C12 = (I22 + 2 * a * T21 - (a ** 2) * ss(I20) - 2 * a * L21.T - 4 * (a ** 2) * I11
           + 2 * (a ** 3) * L10 - 2 * (a ** 3) * T10.T + (a ** 4) * I00)
    C13 = B2C2.T - 2 * a * B2S1.T - (a ** 2) * B2C.T
    C23 = B2S2.T + 2 * a * B2C1.T - (a ** 2) * B2S.T
    C33 = B2B2
    C = np.block([[C11, C12, C13], [C12.T, C22, C23], [C13.T, C23.T, C33]])
    M11 = T20.T - (2 * a) * I10.T - (a ** 2) * T00
    M12 = I20.T + (2 * a) * T10.T - (a ** 2) * I00
    M21 = I20.T - (2 * a) * L10.T - (a ** 2) * I00
    M22 = L20.T + (2 * a) * I10.T - (a ** 2) * L00
    M13, M23 = B2C.T, B2S.T
    M31 = BC2 - (2 * a) * BS1 - (a ** 2) * BC
    M32 = BS2 + (2 * a) * BC1 - (a ** 2) * BS
    M33 = BB2
    M = np.block([[M11, M12, M13], [M21, M22, M23], [M31, M32, M33]])
    Q = (a ** 4) * G + C + (a ** 2) * ss(M)
    # Matrix::forceSymmetric keeps the upper triangle
    Q = np.triu(Q) + np.triu(Q, 1).T
    return Q


def compute_P_sGP(a, k, m, region, accuracy=0.01):
    """R/01_utility.R:255-272 (compute_P for sGP; always boundary=TRUE, A.8)."""
    blocks = [compute_Q_sB(i * a, k, region, accuracy) for i in range(1, m + 1)]
    n = sum(b.shape[0] for b in blocks)
    Q = np.zeros((n, n))
    o = 0
    for b in blocks:
        Q[o:o + b.shape[0], o:o + b.shape[0]] = b
        o += b.shape[0]
    return Q
