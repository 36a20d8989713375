"""Knot-interval moment identities of the O-spline design — ORACLE restatement (test infrastructure).

The product's moment path (``bayesgp_b200/csrc/ospline.cu``) rests on one fact about ``get_local_poly``
(``/root/reference/R/01_utility.R:346-364``): to the right of its own knot interval an O-spline column is a polynomial
of degree p - 1.  This module restates, in plain NumPy on ONE side of the reference location, how ``eta = B U``,
``B^T r`` and ``B^T diag(w) B`` (and the cross block with dense columns ``D``) follow from the per-interval sums
``sum r u^m``, ``sum w u^m``, ``sum w D_c u^m`` — the same formulas as the CUDA kernels, written independently of them.
``tests/test_oracle_moments.py`` checks it against the dense design of ``oracle/basis.py``.
"""
from __future__ import annotations

from math import comb, factorial

import numpy as np


def tail_coefficients(P, d, s):
    """Coefficients al[m] (m < P) in v of  sum_{l=1..P} d^l (v + s)^(P-l) / (l! (P-l)!)  — the tail of a column with
    knot spacing d, expanded about a point s >= 0 to the right of its own interval.  All terms are non-negative."""
    al = np.zeros(P)
    for m in range(P):
        Q = P - m
        al[m] = sum(d ** l * s ** (Q - l) / (factorial(l) * factorial(Q - l)) for l in range(1, Q + 1)) / factorial(m)
    return al


def locate(knots, z):
    """Interval J (knots[J] < z <= knots[J+1]; J = K beyond the last knot) and local coordinate u = z - knots[J];
    z <= knots[0]: every column is zero (J = 0, u = 0)."""
    knots = np.asarray(knots, dtype=np.float64)
    K = len(knots) - 1
    j = np.searchsorted(knots, z, side="left")          # first knot >= z
    J = np.clip(j - 1, 0, K)
    u = np.where(j == 0, 0.0, z - knots[J])
    return J, u


def interval_moments(knots, z, P, r, w, D):
    """R[J, m] = sum r u^m (m <= P), V[J, m] = sum w u^m (m <= 2P), X[J, c, m] = sum w D_c u^m (m <= P)."""
    K = len(knots) - 1
    J, u = locate(knots, z)
    R = np.zeros((K + 1, P + 1))
    V = np.zeros((K + 1, 2 * P + 1))
    X = np.zeros((K + 1, D.shape[1], P + 1))
    for m in range(2 * P + 1):
        np.add.at(V[:, m], J, w * u ** m)
    for m in range(P + 1):
        np.add.at(R[:, m], J, r * u ** m)
        for c in range(D.shape[1]):
            np.add.at(X[:, c, m], J, w * D[:, c] * u ** m)
    return R, V, X


def eta_from_coefficients(knots, z, P, U):
    """B U evaluated as a piecewise polynomial: on interval J, sum_{m<P} C[J, m] u^m + U_J u^P / P!."""
    knots = np.asarray(knots, dtype=np.float64)
    K = len(knots) - 1
    C = np.zeros((K + 1, P))
    for J in range(K + 1):
        for i in range(min(J, K)):
            C[J] += U[i] * tail_coefficients(P, knots[i + 1] - knots[i], knots[J] - knots[i + 1])
    J, u = locate(knots, z)
    own = np.where(J < K, np.append(U, 0.0)[J], 0.0)
    eta = own * u ** P / factorial(P)
    for m in range(P):
        eta = eta + C[J, m] * u ** m
    return eta


def apply_transpose(knots, P, M):
    """out[i] = sum over the observations of (moment weight) * column_i — from per-interval moments M[J, 0..P]:
    the own interval contributes M[i, P] / P!, every interval J > i its tail re-expanded about knots[J]."""
    knots = np.asarray(knots, dtype=np.float64)
    K = len(knots) - 1
    out = np.zeros(K)
    for i in range(K):
        acc = M[i, P] / factorial(P)
        d = knots[i + 1] - knots[i]
        for J in range(i + 1, K + 1):
            acc += tail_coefficients(P, d, knots[J] - knots[i + 1]) @ M[J, :P]
        out[i] = acc
    return out


def hessian_block(knots, P, V):
    """B^T diag(w) B from V[J, m] = sum w u^m: suffix moments about knots[k+1], weighted by column k's tail."""
    knots = np.asarray(knots, dtype=np.float64)
    K = len(knots) - 1
    H = np.zeros((K, K))
    for k in range(K):
        dk = knots[k + 1] - knots[k]
        # S[e] = sum over observations right of interval k of w (z - knots[k+1])^e
        S = np.zeros(2 * P - 1)
        for J in range(k + 1, K + 1):
            s = knots[J] - knots[k + 1]
            for e in range(2 * P - 1):
                S[e] += sum(comb(e, m) * s ** (e - m) * V[J, m] for m in range(e + 1))
        beta = np.array([dk ** (P - q) / (factorial(P - q) * factorial(q)) for q in range(P)])
        G = np.array([sum(beta[q2] * S[q + q2] for q2 in range(P)) for q in range(P)])
        H[k, k] = beta @ G + V[k, 2 * P] / factorial(P) ** 2
        for i in range(k):
            di = knots[i + 1] - knots[i]
            far = tail_coefficients(P, di, knots[k + 1] - knots[i + 1])
            near = tail_coefficients(P, di, knots[k] - knots[i + 1])
            H[i, k] = H[k, i] = far @ G + near @ V[k, P:2 * P] / factorial(P)
    return H
