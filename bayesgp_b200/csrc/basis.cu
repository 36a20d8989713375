// basis.cu — device constructors of the IWP design blocks.
//   B = local_poly_helper(knots, x - x0, order)      /root/reference/R/01_utility.R:378-401
//   X = global_poly(x - x0)[, -1]                     /root/reference/R/01_utility.R:291-300,
//                                                     /root/reference/R/02_model_fit.R:460
//   sGP: B = cbind over harmonics of [B cos(i a x), B sin(i a x), B] with the cubic B-spline basis of
//        fda::create.bspline.basis(region, nbasis = k, norder = 4) minus its first two functions, and
//        X = cbind(cos(i a x), sin(i a x))       /root/reference/R/01_utility.R:177-195,224-239,301-312
// One thread per observation; columns are written coalesced (column-major destination) or as
// one row (observation-major destination).
#include "basis_dev.cuh"
#include "bgp_internal.h"

namespace bgp {

struct IwpArgs {
  const double* x;
  int64_t n;
  double x0;
  const double* kneg;   // mirrored negative knots (ascending, first = 0) or NULL
  int nneg;             // number of knots in kneg
  const double* kpos;
  int npos;
  int order;
  double* B;
  int64_t ldB;
  double* X;            // order-1 columns x^1..x^(order-1), may be NULL
  int64_t ldX;
  int col_major;
};

__global__ void iwp_block_kernel(const IwpArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double x = a.x[i] - a.x0;
  const double xn = x < 0.0 ? -x : 0.0;
  const double xp = x > 0.0 ? x : 0.0;
  int col = 0;
  for (int j = 0; j + 1 < a.nneg; ++j, ++col) {
    const double v = iwp_phi(xn, a.kneg[j], a.kneg[j + 1], a.order);
    if (a.col_major) a.B[(size_t)col * a.ldB + i] = v;
    else a.B[(size_t)i * a.ldB + col] = v;
  }
  // all-positive knots: the reference evaluates get_local_poly on x itself (no clamping)
  const double xe = a.nneg > 0 ? xp : x;
  for (int j = 0; j + 1 < a.npos; ++j, ++col) {
    const double v = iwp_phi(xe, a.kpos[j], a.kpos[j + 1], a.order);
    if (a.col_major) a.B[(size_t)col * a.ldB + i] = v;
    else a.B[(size_t)i * a.ldB + col] = v;
  }
  if (a.X) {
    double pw = 1.0;
    for (int c = 0; c < a.order - 1; ++c) {
      pw *= x;
      if (a.col_major) a.X[(size_t)c * a.ldX + i] = pw;
      else a.X[(size_t)i * a.ldX + c] = pw;
    }
  }
}

int launch_iwp_block(bgp_model* m, const double* x_dev, int64_t n, double x0, const double* kneg, int nneg,
                     const double* kpos, int npos, int order, double* dstB, int ldB, double* dstX, int ldX,
                     bool col_major, cudaStream_t st) {
  (void)m;
  IwpArgs a;
  a.x = x_dev;
  a.n = n;
  a.x0 = x0;
  a.kneg = kneg;
  a.nneg = nneg;
  a.kpos = kpos;
  a.npos = npos;
  a.order = order;
  a.B = dstB;
  a.ldB = ldB;
  a.X = dstX;
  a.ldX = ldX;
  a.col_major = col_major ? 1 : 0;
  const int threads = 256;
  iwp_block_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

struct SgpArgs {
  const double* x;
  int64_t n;
  double x0, a, lo, hi;
  int k, m;
  double* B;     // n x 3 (k-2) m, column-major, pre-zeroed
  double* X;     // n x 2 m, column-major
};

__global__ void sgp_block_kernel(const SgpArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double x = a.x[i] - a.x0;
  const int nb = a.k - 2;
  int first = 0;
  double v4[4] = {0.0, 0.0, 0.0, 0.0};
  const bool inside = x >= a.lo && x <= a.hi;
  if (inside) bspline4(x, a.lo, a.hi, a.k - 2, first, v4);
  for (int h = 1; h <= a.m; ++h) {
    double sn, cs;
    sincos(h * a.a * x, &sn, &cs);
    a.X[(size_t)(2 * (h - 1)) * a.n + i] = cs;
    a.X[(size_t)(2 * (h - 1) + 1) * a.n + i] = sn;
    if (!inside) continue;
    double* Bh = a.B + (size_t)(3 * nb * (h - 1)) * a.n;
    for (int q = 0; q < 4; ++q) {
      const int bi = first + q - 2;            // the first two B-splines are dropped (boundary = TRUE at fit time)
      if (bi < 0 || bi >= nb) continue;
      Bh[(size_t)bi * a.n + i] = v4[q] * cs;
      Bh[(size_t)(nb + bi) * a.n + i] = v4[q] * sn;
      Bh[(size_t)(2 * nb + bi) * a.n + i] = v4[q];
    }
  }
}

int launch_sgp_block(const double* x_dev, int64_t n, double x0, double a, int k, int m, double lo, double hi, double* dstB,
                     double* dstX, cudaStream_t st) {
  SgpArgs s;
  s.x = x_dev;
  s.n = n;
  s.x0 = x0;
  s.a = a;
  s.lo = lo;
  s.hi = hi;
  s.k = k;
  s.m = m;
  s.B = dstB;
  s.X = dstX;
  BGP_CUDA(cudaMemsetAsync(dstB, 0, (size_t)n * 3 * (k - 2) * m * sizeof(double), st));
  sgp_block_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

}  // namespace bgp
