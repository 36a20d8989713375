// Device-side closed forms of the basis functions BayesGP builds in R.
//   O-spline (IWP):  get_local_poly          /root/reference/R/01_utility.R:346-364
//   cubic B-spline:  fda::create.bspline.basis(range, nbasis = k, norder = 4) + eval.basis, the
//                    building block of Compute_B_sB (/root/reference/R/01_utility.R:177-195)
#pragma once

namespace bgp {

__device__ __forceinline__ double ipow(double x, int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= x;
  return r;
}

__device__ __forceinline__ double inv_factorial(int k) {
  // 1/k! for k <= 8
  const double t[9] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0};
  return t[k];
}

// phi_i(x) for the knot interval [k0, k1], smoothness order q
__device__ __forceinline__ double iwp_phi(double x, double k0, double k1, int q) {
  if (x <= k0) return 0.0;
  if (x <= k1) return inv_factorial(q) * ipow(x - k0, q);
  const double dif = k1 - k0, xm = x - k1;
  double s = 0.0;
  for (int l = 1; l <= q; ++l) s += ipow(dif, l) * ipow(xm, q - l) * (inv_factorial(l) * inv_factorial(q - l));
  return s;
}

// All (at most 4) non-zero cubic B-splines at x for equally spaced breaks on [lo, hi] with
// `nbreaks` break points (k = nbreaks + 2 basis functions, 4-fold boundary knots).
// Returns the index of the first non-zero basis function in `first`, values in v[0..3].
__device__ __forceinline__ void bspline4(double x, double lo, double hi, int nbreaks, int& first, double v[4]) {
  const int nint = nbreaks - 1;
  const double h = (hi - lo) / nint;
  int iv = (int)floor((x - lo) / h);
  if (iv < 0) iv = 0;
  if (iv > nint - 1) iv = nint - 1;          // right end point belongs to the last interval
  // knot vector t: 3 extra copies of lo, breaks, 3 extra copies of hi; interval iv = [t[iv+3], t[iv+4])
  auto knot = [&](int idx) {
    int b = idx - 3;
    if (b < 0) b = 0;
    if (b > nint) b = nint;
    return b == nint ? hi : lo + h * b;
  };
  const int mu = iv + 3;
  double N[4] = {1.0, 0.0, 0.0, 0.0};
  // de Boor's triangular scheme (BSPLVB)
  double dl[4], dr[4];
  for (int j = 1; j <= 3; ++j) {
    dr[j] = knot(mu + j) - x;
    dl[j] = x - knot(mu + 1 - j);
    double saved = 0.0;
    for (int r = 0; r < j; ++r) {
      const double term = N[r] / (dr[r + 1] + dl[j - r]);
      N[r] = saved + dr[r + 1] * term;
      saved = dl[j - r] * term;
    }
    N[j] = saved;
  }
  first = iv;
  v[0] = N[0];
  v[1] = N[1];
  v[2] = N[2];
  v[3] = N[3];
}

// The same splines with their first and second derivatives (fda::eval.basis(x, basis, Lfdobj = 1 | 2)):
//   B'_{i,4}  = 3 [ B_{i,3} / (t_{i+3} - t_i) - B_{i+1,3} / (t_{i+4} - t_{i+1}) ]
//   B''_{i,4} = 3 [ B'_{i,3} / (t_{i+3} - t_i) - B'_{i+1,3} / (t_{i+4} - t_{i+1}) ],  B'_{i,3} likewise from order 2.
// Terms whose knot difference is zero belong to splines that vanish identically and are dropped.
__device__ __forceinline__ void bspline4_d(double x, double lo, double hi, int nbreaks, int& first, double v[4],
                                           double d1[4], double d2[4]) {
  const int nint = nbreaks - 1;
  const double h = (hi - lo) / nint;
  int iv = (int)floor((x - lo) / h);
  if (iv < 0) iv = 0;
  if (iv > nint - 1) iv = nint - 1;
  auto knot = [&](int idx) {
    int b = idx - 3;
    if (b < 0) b = 0;
    if (b > nint) b = nint;
    return b == nint ? hi : lo + h * b;
  };
  const int mu = iv + 3;
  double N[4] = {1.0, 0.0, 0.0, 0.0};
  double N2[2] = {0.0, 0.0}, N3[3] = {0.0, 0.0, 0.0};
  double dl[4], dr[4];
  for (int j = 1; j <= 3; ++j) {
    dr[j] = knot(mu + j) - x;
    dl[j] = x - knot(mu + 1 - j);
    double saved = 0.0;
    for (int r = 0; r < j; ++r) {
      const double term = N[r] / (dr[r + 1] + dl[j - r]);
      N[r] = saved + dr[r + 1] * term;
      saved = dl[j - r] * term;
    }
    N[j] = saved;
    if (j == 1) {
      N2[0] = N[0];
      N2[1] = N[1];
    } else if (j == 2) {
      N3[0] = N[0];
      N3[1] = N[1];
      N3[2] = N[2];
    }
  }
  // order-o spline number q (of the o non-zero ones at x) is B_{mu-o+1+q, o}
  auto inv = [&](int a, int b) {
    const double d = knot(b) - knot(a);
    return d > 0.0 ? 1.0 / d : 0.0;
  };
  double dN3[3];     // first derivatives of the three order-3 splines
  for (int q = 0; q < 3; ++q) {
    const int i = mu - 2 + q;
    const double t1 = q >= 1 ? N2[q - 1] * inv(i, i + 2) : 0.0;
    const double t2 = q <= 1 ? N2[q] * inv(i + 1, i + 3) : 0.0;
    dN3[q] = 2.0 * (t1 - t2);
  }
  for (int q = 0; q < 4; ++q) {
    const int i = mu - 3 + q;
    const double w1 = inv(i, i + 3), w2 = inv(i + 1, i + 4);
    d1[q] = 3.0 * ((q >= 1 ? N3[q - 1] * w1 : 0.0) - (q <= 2 ? N3[q] * w2 : 0.0));
    d2[q] = 3.0 * ((q >= 1 ? dN3[q - 1] * w1 : 0.0) - (q <= 2 ? dN3[q] * w2 : 0.0));
    v[q] = N[q];
  }
  first = iv;
}

}  // namespace bgp
