"""CPU prototype (NumPy, float64 + long double tables) of a banded solve for the one-term model: local basis eta = T psi
(differences of order P + 1 of the truncated powers, unit upper-triangular banded T), H~ = T H T^T assembled from the same
knot-interval moments, step and log-determinant through H~.  Result (printed below by running it): the assembly is exact
to 1e-12, but H~ is far worse conditioned than H (Jacobi-scaled condition 2.7e11 against 4.4e6 at K = 100, P = 3: the
polynomial boundary columns are almost in the span of the interior B-splines), log-determinant error 8e-7 relative at
K = 300 — outside the 1e-8 tolerance, so the kernel was not built."""
import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from math import factorial
from oracle import basis

def alpha(P, d, s):
    # coefficients al[m], m<P, of sum_{l=1..P} d^l (v+s)^{P-l}/(l!(P-l)!) in v
    al = np.zeros(P, dtype=np.longdouble)
    for m in range(P):
        Q = P - m
        acc = np.longdouble(0)
        for l in range(1, Q + 1):
            acc += d**l * s**(Q - l) / (factorial(l) * factorial(Q - l))
        al[m] = acc / factorial(m)
    return al

def build_tables(t, P):
    t = np.asarray(t, dtype=np.longdouble)
    K = len(t) - 1
    a = np.zeros((K, P + 1), dtype=np.longdouble)     # eta_q = sum_l a[q,l] psi_{q+l}
    for q in range(K):
        if q <= K - P - 1:
            s = t[q:q + P + 2] - t[q]
            b = np.zeros(P + 2, dtype=np.longdouble)
            for l in range(P + 2):
                prod = np.longdouble(1)
                for mm in range(P + 2):
                    if mm != l:
                        prod *= (s[l] - s[mm])
                b[l] = 1 / prod
            b = b / b[0]
            a[q] = np.cumsum(b)[:P + 1]
        else:
            a[q, 0] = 1
    # E[q][r][m]: coefficients in u = z - t_{q+r} of function q on interval J = q + r
    E = np.zeros((K, P + 1, P + 1), dtype=np.longdouble)
    for q in range(K):
        for r in range(P + 1):
            J = q + r
            if J > K:
                continue
            for l in range(r):          # columns q+l < J: tails
                if a[q, l] == 0 or q + l > K - 1:
                    continue
                col = q + l
                d = t[col + 1] - t[col]
                s = t[J] - t[col + 1]
                E[q, r, :P] += a[q, l] * alpha(P, d, s)
            if J <= K - 1 and a[q, r] != 0 and (q <= K - P - 1 or r == 0):
                E[q, r, P] = a[q, r] / factorial(P)
    return a, E

def run(P, K, n, nonuniform, seed):
    rng = np.random.default_rng(seed)
    if nonuniform:
        t = np.concatenate([[0.0], np.cumsum(rng.uniform(0.2, 1.8, K))]); t /= t[-1] / 1.0
    else:
        t = np.linspace(0, 1, K + 1)
    x = rng.uniform(0, 1.05, n)          # some beyond the last knot
    B = basis.local_poly_helper(t, x, P)     # n x K
    D = np.column_stack([x**i for i in range(1, P)] + [np.ones(n)]) if P > 1 else np.ones((n, 1))
    nD = D.shape[1]
    w = np.exp(rng.normal(0, 1, n))
    theta = -6.0
    dk = np.diff(t)
    A = np.column_stack([D, B])
    H = A.T @ (w[:, None] * A)
    H[nD:, nD:] += np.exp(theta) * np.diag(dk)
    H[:nD, :nD] += 0.01 * np.eye(nD)
    g = rng.normal(0, 1, nD + K)
    x_ref = np.linalg.solve(H, g)
    sign, logdet_ref = np.linalg.slogdet(H)
    # ---- banded route
    a, E = build_tables(t, P)
    T = np.zeros((K, K), dtype=np.longdouble)
    for q in range(K):
        for l in range(P + 1):
            if q + l < K:
                T[q, q + l] = a[q, l]
    # moments per interval
    J = np.searchsorted(t, x, side='left') - 1
    J = np.clip(J, 0, K)
    u = x - t[np.minimum(J, K)]
    u = np.where(x <= t[0], 0.0, u)
    V = np.zeros((K + 1, 2 * P + 1)); X = np.zeros((K + 1, nD, P + 1))
    for m in range(2 * P + 1):
        np.add.at(V[:, m], J, w * u**m)
    for c in range(nD):
        for m in range(P + 1):
            np.add.at(X[:, c, m], J, w * D[:, c] * u**m)
    Ed = E.astype(np.float64)
    Ht = np.zeros((nD + K, nD + K))
    Ht[:nD, :nD] = H[:nD, :nD]
    for q in range(K):
        for q2 in range(max(0, q - P), q + 1):
            acc = 0.0
            for Jv in range(q, min(q2 + P, K) + 1):
                r, r2 = Jv - q, Jv - q2
                for m in range(P + 1):
                    for m2 in range(P + 1):
                        acc += Ed[q, r, m] * Ed[q2, r2, m2] * V[Jv, m + m2]
            Ht[nD + q, nD + q2] = Ht[nD + q2, nD + q] = acc
        for c in range(nD):
            acc = 0.0
            for Jv in range(q, min(q + P, K) + 1):
                r = Jv - q
                for m in range(P + 1):
                    acc += Ed[q, r, m] * X[Jv, c, m]
            Ht[c, nD + q] = Ht[nD + q, c] = acc
    Td = T.astype(np.float64)
    Qt = np.exp(theta) * (Td * dk[None, :]) @ Td.T
    Ht[nD:, nD:] += Qt
    # check against transform of H in extended precision
    TT = np.eye(nD + K, dtype=np.longdouble); TT[nD:, nD:] = T
    Hx = TT @ H.astype(np.longdouble) @ TT.T
    scale = np.sqrt(np.abs(np.diag(Hx)).astype(np.float64))
    err_asm = np.max(np.abs(Ht - Hx.astype(np.float64)) / np.outer(scale, scale))
    gt = (TT @ g.astype(np.longdouble)).astype(np.float64)
    L = np.linalg.cholesky(Ht)
    xt = np.linalg.solve(Ht, gt)
    xb = (TT.T.astype(np.float64)) @ xt
    logdet_b = 2 * np.sum(np.log(np.diag(L)))
    # bandwidth check
    off = np.abs(Ht[nD:, nD:])
    bw = max(abs(i - j) for i in range(K) for j in range(K) if off[i, j] > 0)
    print("P=%d K=%d nonuni=%d: asm err %.2e  step relerr %.2e  logdet rel %.2e (%.6f)  band %d cond(H) %.1e cond(Ht) %.1e" % (
        P, K, nonuniform, err_asm, np.max(np.abs(xb - x_ref)) / np.max(np.abs(x_ref)), abs(logdet_b - logdet_ref) / abs(logdet_ref), logdet_ref, bw,
        np.linalg.cond(H), np.linalg.cond(Ht)))

for P in (1, 2, 3, 4):
    for nonuni in (0, 1):
        run(P, 40, 4000, nonuni, 3 + P)
run(3, 300, 30000, 0, 1)
run(3, 300, 30000, 1, 2)

def scaled_cond(M):
    d = 1/np.sqrt(np.diag(M))
    return np.linalg.cond(M * np.outer(d, d))
