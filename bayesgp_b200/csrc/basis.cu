// basis.cu — device constructors of the IWP design blocks.
//   B = local_poly_helper(knots, x - x0, order)      /root/reference/R/01_utility.R:378-401
//   X = global_poly(x - x0)[, -1]                     /root/reference/R/01_utility.R:291-300,
//                                                     /root/reference/R/02_model_fit.R:460
//   sGP: B = cbind over harmonics of [B cos(i a x), B sin(i a x), B] with the cubic B-spline basis of
//        fda::create.bspline.basis(region, nbasis = k, norder = 4) minus its first two functions, and
//        X = cbind(cos(i a x), sin(i a x))       /root/reference/R/01_utility.R:177-195,224-239,301-312
// One thread per observation; columns are written coalesced (column-major destination) or as
// one row (observation-major destination).
#include <cmath>

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

// ---- Compute_Q_sB on the device (/root/reference/R/01_utility.R:67-174) ----------------------------------------
// The sGP precision of one harmonic a is Q = a^4 G + C + a^2 (M + M^T), G = <phi, phi>, C = <D^2 phi, D^2 phi>,
// M = <phi, D^2 phi> for phi = [B cos(a x), B sin(a x), B], every inner product a Riemann sum over the grid
// x = seq(lo, hi, by = accuracy) with weights diff(c(0, x)).  The reference forms 36 Gram matrices of the nine
// families F = {B, B', B''} x {cos, sin, 1}; here one FP64 tensor-pipe GEMM (kgemm.cu) forms the full 9 nb x 9 nb
// Gram matrix GG = F^T diag(w) F and one kernel assembles Q from its blocks with the reference's formulas.
struct QsbBasisArgs {
  double lo, hi, acc, a;
  int k, nb, nx, ld;      // nb = k - 2 functions kept (dropind = c(1, 2)); ld = row pitch of FT / FWT
  double* FT;             // [9 nb][ld]: family u, function i at row u nb + i, grid point j
  double* FWT;            // the same times the quadrature weight of point j
};

__global__ void qsb_basis_kernel(const QsbBasisArgs q) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= q.ld) return;
  const int rows = 9 * q.nb;
  if (j >= q.nx) {                       // padding columns
    for (int r = 0; r < rows; ++r) q.FT[(size_t)r * q.ld + j] = q.FWT[(size_t)r * q.ld + j] = 0.0;
    return;
  }
  const double x = q.lo + q.acc * j;
  const double wI = j == 0 ? x : x - (q.lo + q.acc * (j - 1));      // diff(c(0, x))
  int first;
  double v[4], d1[4], d2[4];
  bspline4_d(x, q.lo, q.hi, q.k - 2, first, v, d1, d2);
  double sn, cs;
  sincos(q.a * x, &sn, &cs);
  for (int r = 0; r < rows; ++r) q.FT[(size_t)r * q.ld + j] = q.FWT[(size_t)r * q.ld + j] = 0.0;
  for (int t = 0; t < 4; ++t) {
    const int bi = first + t - 2;        // the first two B-splines are dropped
    if (bi < 0 || bi >= q.nb) continue;
    const double f[9] = {v[t] * cs, d1[t] * cs, d2[t] * cs, v[t] * sn, d1[t] * sn, d2[t] * sn, v[t], d1[t], d2[t]};
    for (int u = 0; u < 9; ++u) {
      q.FT[(size_t)(u * q.nb + bi) * q.ld + j] = f[u];
      q.FWT[(size_t)(u * q.nb + bi) * q.ld + j] = f[u] * wI;
    }
  }
}

struct QsbAsmArgs {
  const double* GG;       // 9 nb x 9 nb row-major
  int nb;
  double a;
  double* Q;              // 3 nb x 3 nb column-major destination block (leading dimension ldq)
  int ldq;
};

// un-symmetrised Q[I][J] of Compute_Q_sB; family numbers: 0 Bc 1 B'c 2 B''c 3 Bs 4 B's 5 B''s 6 B 7 B' 8 B''
__device__ double qsb_entry(const QsbAsmArgs& q, int I, int J) {
  const int nb = q.nb, n9 = 9 * q.nb;
  const double a = q.a, a2 = a * a, a3 = a2 * a, a4 = a2 * a2;
  auto g = [&](int u, int v, int i, int j) { return q.GG[(size_t)(u * nb + i) * n9 + v * nb + j]; };
  auto ss = [&](int u, int v, int i, int j) { return g(u, v, i, j) + g(u, v, j, i); };
  const int bi = I / nb, bj = J / nb, i = I % nb, j = J % nb;
  // m(bi, bj, i, j) = M[I][J]; Q needs M + M^T
  auto Mblk = [&](int bi_, int bj_, int i_, int j_) -> double {
    switch (bi_ * 3 + bj_) {
      case 0: return g(2, 0, j_, i_) - 2.0 * a * g(4, 0, j_, i_) - a2 * g(0, 0, i_, j_);                 // M11
      case 1: return g(5, 0, j_, i_) + 2.0 * a * g(1, 0, j_, i_) - a2 * g(3, 0, i_, j_);                 // M12
      case 2: return g(8, 0, j_, i_);                                                                   // M13 = t(B2C)
      case 3: return g(5, 0, j_, i_) - 2.0 * a * g(4, 3, j_, i_) - a2 * g(3, 0, i_, j_);                 // M21
      case 4: return g(5, 3, j_, i_) + 2.0 * a * g(4, 0, j_, i_) - a2 * g(3, 3, i_, j_);                 // M22
      case 5: return g(8, 3, j_, i_);                                                                   // M23 = t(B2S)
      case 6: return g(6, 2, i_, j_) - 2.0 * a * g(6, 4, i_, j_) - a2 * g(6, 0, i_, j_);                 // M31
      case 7: return g(6, 5, i_, j_) + 2.0 * a * g(6, 1, i_, j_) - a2 * g(6, 3, i_, j_);                 // M32
      default: return g(6, 8, i_, j_);                                                                  // M33 = BB2
    }
  };
  double G, C;
  switch (bi * 3 + bj) {
    case 0:
      G = g(0, 0, i, j);
      C = g(2, 2, i, j) - 2.0 * a * ss(5, 1, i, j) - a2 * ss(2, 0, i, j) + 2.0 * a3 * ss(4, 0, i, j) + 4.0 * a2 * g(4, 4, i, j) +
          a4 * g(0, 0, i, j);
      break;
    case 4:
      G = g(3, 3, i, j);
      C = g(5, 5, i, j) + 2.0 * a * ss(5, 1, i, j) - a2 * ss(5, 3, i, j) - 2.0 * a3 * ss(4, 0, i, j) + 4.0 * a2 * g(1, 1, i, j) +
          a4 * g(3, 3, i, j);
      break;
    case 1:      // C12
    case 3: {    // t(C12)
      const int ii = bi == 0 ? i : j, jj = bi == 0 ? j : i;
      G = bi == 0 ? g(3, 0, j, i) : g(3, 0, i, j);      // t(I00) | I00
      C = g(5, 2, ii, jj) + 2.0 * a * g(2, 1, ii, jj) - a2 * ss(5, 0, ii, jj) - 2.0 * a * g(5, 4, jj, ii) -
          4.0 * a2 * g(4, 1, ii, jj) + 2.0 * a3 * g(4, 3, ii, jj) - 2.0 * a3 * g(1, 0, jj, ii) + a4 * g(3, 0, ii, jj);
      break;
    }
    case 2:      // C13 = t(B2C2) - 2a t(B2S1) - a^2 t(B2C)
      G = g(6, 0, j, i);                                  // t(BC)
      C = g(8, 2, j, i) - 2.0 * a * g(8, 4, j, i) - a2 * g(8, 0, j, i);
      break;
    case 6:      // t(C13)
      G = g(6, 0, i, j);                                  // BC
      C = g(8, 2, i, j) - 2.0 * a * g(8, 4, i, j) - a2 * g(8, 0, i, j);
      break;
    case 5:      // C23 = t(B2S2) + 2a t(B2C1) - a^2 t(B2S)
      G = g(6, 3, j, i);                                  // t(BS)
      C = g(8, 5, j, i) + 2.0 * a * g(8, 1, j, i) - a2 * g(8, 3, j, i);
      break;
    case 7:      // t(C23)
      G = g(6, 3, i, j);                                  // BS
      C = g(8, 5, i, j) + 2.0 * a * g(8, 1, i, j) - a2 * g(8, 3, i, j);
      break;
    default:     // 8
      G = g(6, 6, i, j);
      C = g(8, 8, i, j);
      break;
  }
  return a4 * G + C + a2 * (Mblk(bi, bj, i, j) + Mblk(bj, bi, j, i));
}

__global__ void qsb_assemble_kernel(const QsbAsmArgs q) {
  const int n3 = 3 * q.nb;
  const int I = blockIdx.x * blockDim.x + threadIdx.x, J = blockIdx.y;
  if (I >= n3) return;
  // Matrix::forceSymmetric(Q): the upper triangle is kept and mirrored
  q.Q[(size_t)J * q.ldq + I] = I <= J ? qsb_entry(q, I, J) : qsb_entry(q, J, I);
}

// P (d x d column-major on the device, d = 3 (k - 2) m, block diagonal over the harmonics, zero elsewhere)
int launch_sgp_precision(double a, int k, int m, double lo, double hi, double accuracy, double* P_dev, cudaStream_t st) {
  const int nb = k - 2, d = 3 * nb * m;
  const int nx = (int)std::floor((hi - lo) / accuracy + 1e-10) + 1;
  const int ld = round_up(nx, 16);
  const int rows = 9 * nb;
  double *FT = nullptr, *FWT = nullptr, *GG = nullptr;
  BGP_CUDA(cudaMemsetAsync(P_dev, 0, (size_t)d * d * sizeof(double), st));
  BGP_CUDA(cudaMalloc(&FT, (size_t)round_up(rows, 128) * ld * sizeof(double)));
  BGP_CUDA(cudaMalloc(&FWT, (size_t)round_up(rows, 128) * ld * sizeof(double)));
  BGP_CUDA(cudaMalloc(&GG, (size_t)rows * rows * sizeof(double)));
  int rc = [&]() -> int {
    for (int hh = 1; hh <= m; ++hh) {
      QsbBasisArgs b;
      b.lo = lo;
      b.hi = hi;
      b.acc = accuracy;
      b.a = hh * a;
      b.k = k;
      b.nb = nb;
      b.nx = nx;
      b.ld = ld;
      b.FT = FT;
      b.FWT = FWT;
      qsb_basis_kernel<<<(ld + 127) / 128, 128, 0, st>>>(b);
      count_launch();
      BGP_TRY(launch_kgemm(FT, rows, ld, FWT, rows, ld, ld, nullptr, GG, rows, false, nullptr, st));
      QsbAsmArgs q;
      q.GG = GG;
      q.nb = nb;
      q.a = hh * a;
      q.Q = P_dev + (size_t)(hh - 1) * 3 * nb * d + (size_t)(hh - 1) * 3 * nb;
      q.ldq = d;
      dim3 grid((3 * nb + 127) / 128, 3 * nb);
      qsb_assemble_kernel<<<grid, 128, 0, st>>>(q);
      count_launch();
      BGP_CUDA(cudaGetLastError());
    }
    BGP_CUDA(cudaStreamSynchronize(st));
    return BGP_OK;
  }();
  cudaFree(FT);
  cudaFree(FWT);
  cudaFree(GG);
  return rc;
}

}  // namespace bgp
