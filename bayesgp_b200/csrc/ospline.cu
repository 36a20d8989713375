// ospline.cu — the O-spline moment path: likelihood pass and Hessian of a model with ONE IWP term in O(n + K^2).
//
// Replaces, for models whose only smoothing term is an IWP, the two passes over the dense design (lik.cu, syrk.cu).
// The reference builds the IWP design from get_local_poly (/root/reference/R/01_utility.R:346-364): with knots
// t_0 < ... < t_K and z the distance from the reference location, column i is
//     0                                   z <= t_i
//     (z - t_i)^P / P!                    t_i < z <= t_{i+1}
//     sum_{l=1..P} d_i^l (z - t_{i+1})^{P-l} / (l! (P-l)!)      z > t_{i+1}      (d_i = t_{i+1} - t_i)
// — to the right of its own knot interval every column is a POLYNOMIAL of degree P-1 in z.  For an observation in
// interval J with local coordinate u = z - t_J in (0, d_J]:
//   * eta = X beta + sum_i B_i U_i  (src/BayesGP.cpp:133-145)  is  D beta + W_J u^P / P! + sum_{m<P} C_{J,m} u^m, where
//     C_J collects the tails of the columns i < J re-expanded about t_J;
//   * g_lik = A^T r and H_lik = A^T diag(w) A  (the AD sweeps of TMB on that objective) are linear in the
//     per-interval moments  sum r u^m, sum w u^m, sum w D_c u^m  and a handful of global sums over the dense columns.
// One streaming pass over (u, y, size, dense columns) — 40-60 bytes per observation instead of the 8 p bytes of a
// design row — produces those moments; the p-sized assembly is O(K^2 P^2) flops in small kernels.  Every re-expansion
// shifts a polynomial in (z - t) to a knot further LEFT, so all binomial terms are positive: no cancellation, the
// results agree with the dense contraction to rounding (tests/test_gpu_ospline.py).
//
// Determinism: observations are sorted by knot interval once (stable radix sort), intervals are cut into pieces of a
// fixed length, one warp per piece, partial moments are combined in a fixed order — bit-reproducible like the dense path.
// The dense design stays resident: the Laplace gradient (grad.cu) and every model with more than one smoothing term
// use it.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdlib>

#include "bgp_internal.h"
#include "finish_dev.cuh"
#include "lik_terms.cuh"

namespace bgp {

constexpr int OSP_MAXP = 4;      // smoothness orders 1..4
constexpr int OSP_MAXD = 8;      // dense (boundary + fixed) columns
constexpr int OSP_STEP = OSP_MAXP * (OSP_MAXP + 1) / 2 + OSP_MAXP;   // per-column constants of the leverage recurrence

struct OspPlan {
  int P = 0, nD = 0, NDC = 0, NG = 0, NC = 0, NM = 0, NACC = 0, np = 0;
  int pass_grid = 0;              // resident CTAs of the pass kernel
  int64_t n = 0;
  // observations in interval order
  double *u = nullptr, *y = nullptr, *size = nullptr, *D = nullptr, *eta = nullptr;
  int64_t* piece_beg = nullptr;   // np + 1
  int* piece_gid = nullptr;       // np
  int* gid_pbeg = nullptr;        // NG + 1
  // per interval (gid): left knot, own column (-1 beyond the last knot), first column of the side, number of tails
  double* g_t0 = nullptr;
  int *g_own = nullptr, *g_c0 = nullptr, *g_nt = nullptr;
  // per column: its knots, own interval, end of the side's intervals, side
  double *c_t0 = nullptr, *c_t1 = nullptr;
  int *c_gid = nullptr, *c_gend = nullptr, *c_side = nullptr;
  // per evaluation
  int* done = nullptr;
  // leverage path (allocated on the first gradient): V G_J per interval and its Gram matrix
  int nside = 0, ldk = 0, maxK = 0;
  int *side_gbase = nullptr, *side_cbase = nullptr, *side_K = nullptr;
  std::vector<double> c_t0_host, c_t1_host;
  double *Yt = nullptr, *Omega = nullptr, *c_step = nullptr;
  double *slots = nullptr, *mom = nullptr, *glob = nullptr, *Hdb = nullptr, *G = nullptr;
};

// accumulator layout of a piece: [R (P+1) | V (2P+1) | X (NDC x (P+1)) | DD (NDC (NDC+1)/2) | gD (NDC) | ll sumsq bad | max]
__host__ __device__ constexpr int osp_offV(int P) { return P + 1; }
__host__ __device__ constexpr int osp_offX(int P) { return 3 * P + 2; }
__host__ __device__ constexpr int osp_NM(int P, int NDC) { return 3 * P + 2 + NDC * (P + 1); }
__host__ __device__ constexpr int osp_offDD(int P, int NDC) { return osp_NM(P, NDC); }
__host__ __device__ constexpr int osp_offgD(int P, int NDC) { return osp_offDD(P, NDC) + NDC * (NDC + 1) / 2; }
__host__ __device__ constexpr int osp_offS(int P, int NDC) { return osp_offgD(P, NDC) + NDC; }
__host__ __device__ constexpr int osp_NACC(int P, int NDC) { return osp_offS(P, NDC) + 4; }

__host__ __device__ constexpr double osp_ifact(int k) {
  double f = 1.0;
  for (int i = 2; i <= k; ++i) f *= (double)i;
  return 1.0 / f;
}

// coefficients of the tail of a column with knot spacing d, expanded about a point s >= 0 to the right of its own
// interval:  sum_{l=1..P} d^l (v + s)^{P-l} / (l! (P-l)!) = sum_{m<P} al[m] v^m,
//            al[m] = (1/m!) sum_{l=1..P-m} d^l s^{P-m-l} / (l! (P-m-l)!)       (all terms non-negative)
template <int P>
__device__ __forceinline__ void osp_alpha(double d, double s, double (&al)[P]) {
  double dp[P + 1], sp[P + 1];
  dp[0] = sp[0] = 1.0;
#pragma unroll
  for (int l = 1; l <= P; ++l) {
    dp[l] = dp[l - 1] * d;
    sp[l] = sp[l - 1] * s;
  }
#pragma unroll
  for (int m = 0; m < P; ++m) {
    const int Q = P - m;
    double acc = 0.0;
#pragma unroll
    for (int l = 1; l <= Q; ++l) acc = fma(dp[l] * sp[Q - l], osp_ifact(l) * osp_ifact(Q - l), acc);
    al[m] = acc * osp_ifact(m);
  }
}

__device__ __forceinline__ double osp_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- one-time layout ---------------------------------------------------------------------------------------------
struct OspLocateArgs {
  const double* x;
  int64_t n;
  double x0;
  const double *kneg, *kpos;
  int nkn, nkp, gbase_pos;
  uint32_t *gid, *idx;
  double* u;
};

__global__ void osp_locate_kernel(const OspLocateArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double xx = a.x[i] - a.x0;
  const double* t;
  int nk, gbase;
  double z;
  if (a.nkn > 0 && (xx < 0.0 || a.nkp == 0)) {       // local_poly_helper: the negative part, mirrored
    z = xx < 0.0 ? -xx : 0.0;
    t = a.kneg;
    nk = a.nkn;
    gbase = 0;
  } else {
    z = a.nkn > 0 ? (xx > 0.0 ? xx : 0.0) : xx;      // all-positive knots: evaluated on x itself
    t = a.kpos;
    nk = a.nkp;
    gbase = a.gbase_pos;
  }
  // j = first knot >= z;  interval J = j - 1  (t_J < z <= t_{J+1});  z <= t_0: every column is zero (J = 0, u = 0)
  int lo = 0, hi = nk;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t[mid] < z) lo = mid + 1; else hi = mid;
  }
  int J = lo - 1;
  double u = 0.0;
  if (J < 0) J = 0; else u = z - t[J];
  a.gid[i] = (uint32_t)(gbase + J);
  a.idx[i] = (uint32_t)i;
  a.u[i] = u;
}

__global__ void osp_gather_kernel(const double* __restrict__ src, const uint32_t* __restrict__ perm, int64_t n,
                                  double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

__global__ void osp_count_kernel(const uint32_t* __restrict__ gid, int64_t n, int* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(cnt + gid[i], 1);
}

// ---- per evaluation ----------------------------------------------------------------------------------------------
struct OspPassArgs {
  const double *u, *y, *size, *D;
  double* eta;
  const int64_t* piece_beg;
  const int* piece_gid;
  const int *g_own, *g_c0, *g_nt;
  const double *g_t0, *c_t0, *c_t1;
  const double* W;
  int nD, np, family;
  int64_t n;
  double tau;
  double* slots;
};

// One warp per piece (a run of observations of one knot interval): eta, the likelihood terms and the moments.
template <int P, int NDC>
__global__ void __launch_bounds__(128, NDC <= 4 ? 3 : 2) osp_pass_kernel(const OspPassArgs a) {
  constexpr int OV = osp_offV(P), OX = osp_offX(P), ODD = osp_offDD(P, NDC), OGD = osp_offgD(P, NDC), OS = osp_offS(P, NDC);
  constexpr int NACC = osp_NACC(P, NDC);
  const int lane = threadIdx.x & 31;
  // resident warps stride over the pieces (similar lengths): no partial last wave
  for (int pc = blockIdx.x * 4 + (threadIdx.x >> 5); pc < a.np; pc += gridDim.x * 4) {
  const int gid = a.piece_gid[pc];
  const int64_t j0 = a.piece_beg[pc], j1 = a.piece_beg[pc + 1];
  // cf[m]: coefficients (in u) of the tails of the columns left of this interval, weighted by W — the same sums in
  // the same order for every piece of the interval
  double cf[P];
#pragma unroll
  for (int m = 0; m < P; ++m) cf[m] = 0.0;
  {
    const int c0 = a.g_c0[gid], J = a.g_nt[gid];
    const double t0 = a.g_t0[gid];
    for (int i = lane; i < J; i += 32) {
      const int col = c0 + i;
      const double t1 = a.c_t1[col];
      double al[P];
      osp_alpha<P>(t1 - a.c_t0[col], t0 - t1, al);
      const double wv = a.W[a.nD + col];
#pragma unroll
      for (int m = 0; m < P; ++m) cf[m] = fma(wv, al[m], cf[m]);
    }
#pragma unroll
    for (int m = 0; m < P; ++m) cf[m] = osp_warp_sum(cf[m]);
  }
  const int own = a.g_own[gid];
  const double wown = own >= 0 ? a.W[a.nD + own] * osp_ifact(P) : 0.0;
  double wd[NDC];
#pragma unroll
  for (int c = 0; c < NDC; ++c) wd[c] = c < a.nD ? a.W[c] : 0.0;
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  double dmax = 0.0;
  // software pipeline: the next observation's inputs are requested before this one's arithmetic (the eta store below
  // may alias them as far as the compiler knows, so it would not hoist the loads itself)
  int64_t jn = j0 + lane;
  double nu = 0.0, ny = 0.0, nsz = 1.0, neta = 0.0, nd[NDC];
#pragma unroll
  for (int c = 0; c < NDC; ++c) nd[c] = 0.0;
  if (jn < j1) {
    nu = a.u[jn];
    ny = a.y[jn];
    if (a.size) nsz = a.size[jn];
    neta = a.eta[jn];
#pragma unroll
    for (int c = 0; c < NDC; ++c)
      if (c < a.nD) nd[c] = a.D[(size_t)c * a.n + jn];
  }
  while (jn < j1) {
    const int64_t j = jn;
    const double u = nu, yv = ny, sz = nsz, eta_old = neta;
    double dv[NDC];
#pragma unroll
    for (int c = 0; c < NDC; ++c) dv[c] = nd[c];
    jn += 32;
    if (jn < j1) {
      nu = a.u[jn];
      ny = a.y[jn];
      if (a.size) nsz = a.size[jn];
      neta = a.eta[jn];
#pragma unroll
      for (int c = 0; c < NDC; ++c)
        if (c < a.nD) nd[c] = a.D[(size_t)c * a.n + jn];
    }
    double up[2 * P + 1];
    up[0] = 1.0;
#pragma unroll
    for (int m = 1; m <= 2 * P; ++m) up[m] = up[m - 1] * u;
    double eta = wown * up[P];
#pragma unroll
    for (int m = P - 1; m >= 0; --m) eta = fma(cf[m], up[m], eta);
#pragma unroll
    for (int c = 0; c < NDC; ++c) eta = fma(dv[c], wd[c], eta);
    double r, w, c3;
    obs_terms(a.family, a.tau, eta, yv, sz, acc[OS], acc[OS + 1], r, w, c3);
    if (!(isfinite(w) && isfinite(r) && isfinite(acc[OS]))) acc[OS + 2] = 1.0;
    const double dd = fabs(eta - eta_old);
    dmax = dd > dmax || !(dd == dd) ? (dd == dd ? dd : INFINITY) : dmax;
    a.eta[j] = eta;
#pragma unroll
    for (int m = 0; m <= P; ++m) acc[m] = fma(r, up[m], acc[m]);
#pragma unroll
    for (int m = 0; m <= 2 * P; ++m) acc[OV + m] = fma(w, up[m], acc[OV + m]);
#pragma unroll
    for (int c = 0; c < NDC; ++c) {
      const double wdc = w * dv[c];
#pragma unroll
      for (int m = 0; m <= P; ++m) acc[OX + c * (P + 1) + m] = fma(wdc, up[m], acc[OX + c * (P + 1) + m]);
#pragma unroll
      for (int c2 = 0; c2 <= c; ++c2) acc[ODD + c * (c + 1) / 2 + c2] = fma(wdc, dv[c2], acc[ODD + c * (c + 1) / 2 + c2]);
      acc[OGD + c] = fma(r, dv[c], acc[OGD + c]);
    }
  }
  // fixed butterfly: every lane ends with the piece's sums; lane k % 32 stores value k
  double* slot = a.slots + (size_t)pc * NACC;
#pragma unroll
  for (int k = 0; k < NACC - 1; ++k) {
    const double v = osp_warp_sum(acc[k]);
    if (lane == (k & 31)) slot[k] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if (lane == 0) slot[NACC - 1] = dmax;
  }
}

// Interval moments (one warp per interval, lanes over its pieces) and the global sums (one CTA per value).
__global__ void __launch_bounds__(128) osp_reduce_kernel(const double* __restrict__ slots, int NACC, int NM, int NG,
                                                         const int* __restrict__ gid_pbeg, int np, double* __restrict__ mom,
                                                         double* __restrict__ glob) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nA = (NG + 3) / 4;
  if ((int)blockIdx.x < nA) {
    const int gid = blockIdx.x * 4 + warp;
    if (gid >= NG) return;
    const int pb = gid_pbeg[gid], pe = gid_pbeg[gid + 1];
    if (pe - pb <= 64) {
      // the usual case: lane m sums moment m over the interval's pieces in order (coalesced rows, no shuffles)
      for (int m = lane; m < NM; m += 32) {
        double v = 0.0;
#pragma unroll 8
        for (int pc = pb; pc < pe; ++pc) v += slots[(size_t)pc * NACC + m];
        mom[(size_t)gid * NM + m] = v;
      }
      return;
    }
    for (int m = 0; m < NM; ++m) {
      double v = 0.0;
      for (int pc = pb + lane; pc < pe; pc += 32) v += slots[(size_t)pc * NACC + m];
      v = osp_warp_sum(v);
      if (lane == 0) mom[(size_t)gid * NM + m] = v;
    }
    return;
  }
  __shared__ double sm[128];
  const int gi = blockIdx.x - nA;             // global value NM + gi
  const bool is_max = NM + gi == NACC - 1;
  // eight independent loads in flight per thread (the pieces are ~24 deep per thread on a large model)
  double v = 0.0;
  const double* src = slots + NM + gi;
  int pc = threadIdx.x;
  for (; pc + 7 * 128 < np; pc += 8 * 128) {
    double t[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) t[q] = src[(size_t)(pc + q * 128) * NACC];
#pragma unroll
    for (int q = 0; q < 8; ++q) v = is_max ? fmax(v, t[q]) : v + t[q];
  }
  for (; pc < np; pc += 128) {
    const double t = src[(size_t)pc * NACC];
    v = is_max ? fmax(v, t) : v + t;
  }
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] = is_max ? fmax(sm[threadIdx.x], sm[threadIdx.x + o]) : sm[threadIdx.x] + sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) glob[gi] = sm[0];
}

struct OspApplyArgs {
  int NC, nD, NM, lda;
  const double *c_t0, *c_t1, *g_t0;
  const int *c_gid, *c_gend;
  const double *mom, *glob;
  double* red;      // [g_lik (lda) | ll | sumsq | bad | max d eta]
  double* Hdb;      // NDC x NC
  int* done;        // CTA counter for the fused prior completion (self-resetting)
  int fuse_prior;
  double* G;        // NC x OSP_MAXP: G[k][q] = sum over the observations right of column k's interval of w c_k(z) (z - t_{k+1})^q
};

// g_lik of the spline columns, the {dense x spline} block of H and the weighted suffix moments G the {spline x spline}
// block is assembled from: one CTA per column, threads over the intervals to its right.  The last CTA moves the global
// sums into the reduction buffer finish.cu reads.
template <int P, int NDC>
__device__ __forceinline__ void osp_apply_body(const OspApplyArgs& a) {
  constexpr int OV = osp_offV(P), OX = osp_offX(P), OGD = osp_offgD(P, NDC) - osp_NM(P, NDC), OS = osp_offS(P, NDC) - osp_NM(P, NDC);
  constexpr int NE = 2 * P - 1;
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == gridDim.x - 1) {
    for (int c = threadIdx.x; c < a.nD; c += 128) a.red[c] = a.glob[OGD + c];
    if (threadIdx.x < 4) a.red[a.lda + threadIdx.x] = a.glob[OS + threadIdx.x];
    return;
  }
  // one CTA per column: its 128 threads stride over the intervals to the right, partial sums are combined warp by
  // warp in a fixed order
  __shared__ double s_part[4][NDC + 1 + NE];
  __shared__ double s_fin[NDC + 1 + NE];
  const int col = blockIdx.x, warp = threadIdx.x >> 5;
  const int own = a.c_gid[col], gend = a.c_gend[col];
  const double t1 = a.c_t1[col], d = t1 - a.c_t0[col];
  double acc[NDC + 1], S[NE];
#pragma unroll
  for (int e = 0; e <= NDC; ++e) acc[e] = 0.0;
#pragma unroll
  for (int e = 0; e < NE; ++e) S[e] = 0.0;
  for (int g = own + 1 + threadIdx.x; g < gend; g += 128) {
    const double s = a.g_t0[g] - t1;
    double al[P];
    osp_alpha<P>(d, s, al);
    const double* mg = a.mom + (size_t)g * a.NM;
#pragma unroll
    for (int m = 0; m < P; ++m) {
      acc[0] = fma(al[m], mg[m], acc[0]);
#pragma unroll
      for (int c = 0; c < NDC; ++c) acc[1 + c] = fma(al[m], mg[OX + c * (P + 1) + m], acc[1 + c]);
    }
    // sum w (u + s)^e = sum_{m<=e} binom(e, m) s^{e-m} V_m
    double sp[NE];
    sp[0] = 1.0;
#pragma unroll
    for (int e = 1; e < NE; ++e) sp[e] = sp[e - 1] * s;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      double binom = 1.0;
#pragma unroll
      for (int m = 0; m <= e; ++m) {
        S[e] = fma(binom * sp[e - m], mg[OV + m], S[e]);
        binom = binom * (double)(e - m) / (double)(m + 1);
      }
    }
  }
#pragma unroll
  for (int e = 0; e <= NDC; ++e) {
    const double v = osp_warp_sum(acc[e]);
    if (lane == 0) s_part[warp][e] = v;
  }
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const double v = osp_warp_sum(S[e]);
    if (lane == 0) s_part[warp][NDC + 1 + e] = v;
  }
  __syncthreads();
  if (threadIdx.x < NDC + 1 + NE)
    s_fin[threadIdx.x] = (s_part[0][threadIdx.x] + s_part[1][threadIdx.x]) + (s_part[2][threadIdx.x] + s_part[3][threadIdx.x]);
  __syncthreads();
  if (threadIdx.x < P) {
    // c_k(z) = sum_{q'} beta_{q'} v^{q'},  beta_{q'} = d^{P-q'} / ((P-q')! q'!)
    double dp[P + 1];
    dp[0] = 1.0;
#pragma unroll
    for (int l = 1; l <= P; ++l) dp[l] = dp[l - 1] * d;
    double gq = 0.0;
#pragma unroll
    for (int q2 = 0; q2 < P; ++q2)
      gq = fma(dp[P - q2] * (osp_ifact(P - q2) * osp_ifact(q2)), s_fin[NDC + 1 + threadIdx.x + q2], gq);
    a.G[col * OSP_MAXP + threadIdx.x] = gq;
  }
  const double* mo = a.mom + (size_t)own * a.NM;
  if (threadIdx.x == 32) a.red[a.nD + col] = s_fin[0] + mo[P] * osp_ifact(P);
  if (threadIdx.x >= 64 && (int)threadIdx.x - 64 < a.nD) {
    const int c = threadIdx.x - 64;
    a.Hdb[(size_t)c * a.NC + col] = s_fin[1 + c] + mo[OX + c * (P + 1) + P] * osp_ifact(P);
  }
}

// fuse_prior (single-device models): the CTA that finishes last completes f, g and max|g| (finish_dev.cuh) — every
// entry of the reduction buffer is in L2 by then (fence + ticket), and one CTA does the whole completion in a fixed
// order, so the result does not depend on which CTA that is.
template <int P, int NDC>
__global__ void __launch_bounds__(128) osp_apply_kernel(const OspApplyArgs a, const PriorArgs pr) {
  osp_apply_body<P, NDC>(a);
  if (!a.fuse_prior) return;
  __shared__ int s_last;
  __shared__ double s_quad[128], s_gmax[128];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(a.done, 1);
    s_last = t == (int)gridDim.x - 1;
    if (s_last) *a.done = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  finish_prior_body<128>(pr, s_quad, s_gmax);
}

struct OspHArgs {
  int p, ldh, nD, NC, NM;
  const double *c_t0, *c_t1;
  const int *c_gid, *c_side;
  const double *mom, *glob, *Hdb, *G;
  double *H, *L;      // L != NULL: a second copy for the Cholesky kernel, which factors in place (saves its device-to-device copy)
  // Q(theta) on the diagonal (an IWP precision is diagonal): qfix on the dense columns, e^theta P_i on the spline
  // columns; NULL when the likelihood part is all-reduced over observation shards first
  const double *qfix, *Pdiag;
  double etheta;
};

// H_lik, both triangles, internal column order (dense columns first): one thread per entry.
template <int P, int NDC>
__global__ void __launch_bounds__(256) osp_hwrite_kernel(const OspHArgs a) {
  constexpr int OV = osp_offV(P);
  const int r = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
  if (r >= a.p) return;
  const int lo = r < c ? r : c, hi = r < c ? c : r;
  double v;
  if (hi < a.nD) {
    v = a.glob[hi * (hi + 1) / 2 + lo];            // DD leads the global block
  } else if (lo < a.nD) {
    v = a.Hdb[(size_t)lo * a.NC + (hi - a.nD)];
  } else {
    const int i = lo - a.nD, k = hi - a.nD;
    if (a.c_side[i] != a.c_side[k]) {
      v = 0.0;                                     // no observation has both sides of the reference location
    } else {
      const double* Gk = a.G + k * OSP_MAXP;
      const double* Vk = a.mom + (size_t)a.c_gid[k] * a.NM + OV;
      const double t0k = a.c_t0[k], t1k = a.c_t1[k];
      if (i == k) {
        const double d = t1k - t0k;
        double dp[P + 1];
        dp[0] = 1.0;
#pragma unroll
        for (int l = 1; l <= P; ++l) dp[l] = dp[l - 1] * d;
        v = Vk[2 * P] * (osp_ifact(P) * osp_ifact(P));
#pragma unroll
        for (int q = 0; q < P; ++q) v = fma(dp[P - q] * (osp_ifact(P - q) * osp_ifact(q)), Gk[q], v);
      } else {
        const double t1i = a.c_t1[i], di = t1i - a.c_t0[i];
        double af[P], an[P];
        osp_alpha<P>(di, t1k - t1i, af);           // about t_{k+1}: the observations right of interval k
        osp_alpha<P>(di, t0k - t1i, an);           // about t_k: the observations inside interval k
        v = 0.0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
          v = fma(af[q], Gk[q], v);
          v = fma(an[q] * osp_ifact(P), Vk[P + q], v);
        }
      }
    }
  }
  if (r == c && a.qfix) v += r < a.nD ? a.qfix[r] : a.etheta * a.Pdiag[r - a.nD];
  a.H[(size_t)c * a.ldh + r] = v;
  if (a.L) a.L[(size_t)c * a.ldh + r] = v;
}


// ---- leverages for the Laplace gradient -----------------------------------------------------------------------------
// q_j = a_j^T H^-1 a_j = || V a_j ||^2 with the upper factor V of grad.cu.  A design row is a_j = G_J v_j with
// v_j = [1, u, .., u^(P-1) | u^P | D_j] and G_J depending on the knot interval only, so q_j = v_j^T Omega_J v_j with
// Omega_J = (V G_J)^T (V G_J), a (P + 1 + nD)-square matrix per interval:
//   1. osp_levY: Y_J = V [tails of the columns left of J, expanded about t_J] by the recurrence
//      Y_J = shift(Y_{J-1}, d_{J-1}) + V[:, J-1] b_{J-1}  (one thread per row of V and side, K sequential steps);
//   2. osp_levOmega: Gram matrix of [Y_J | V[:, own] / P! | V[:, dense]] per interval;
//   3. osp_levpass: per observation q_j, z_j = c3_j q_j (c3 = d w / d eta from the eta of the mode's own pass) and the
//      moments sum z u^m, sum z D_c — A^T z then comes out of the same reduce / apply kernels as g_lik.
template <int P>
__global__ void __launch_bounds__(128) osp_levY_kernel(const double* __restrict__ V, int p, int ldl, int nD, int ldk, int nside,
                                                       const int* __restrict__ side_gbase, const int* __restrict__ side_cbase,
                                                       const int* __restrict__ side_K, const double* __restrict__ c_step,
                                                       double* __restrict__ Yt) {
  extern __shared__ double s_cs[];                 // the side's per-column constants: K x OSP_STEP
  const int k = blockIdx.x * 128 + threadIdx.x, sd = blockIdx.y;
  if (sd >= nside) return;
  const int gbase = side_gbase[sd], cbase = side_cbase[sd], K = side_K[sd];
  for (int t = threadIdx.x; t < K * OSP_STEP; t += 128) s_cs[t] = c_step[(size_t)cbase * OSP_STEP + t];
  __syncthreads();
  if (k >= p) return;
  double y[P];
#pragma unroll
  for (int m = 0; m < P; ++m) {
    y[m] = 0.0;
    Yt[((size_t)gbase * P + m) * ldk + k] = 0.0;
  }
  const double* Vk = V + (size_t)k * ldl + nD + cbase;
  constexpr int CH = 16;                            // this row's entries of V, sixteen loads in flight at a time
  for (int J0 = 0; J0 < K; J0 += CH) {
    double v[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) v[q] = J0 + q < K ? Vk[J0 + q] : 0.0;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int J = J0 + q + 1;
      if (J <= K) {
        // shift matrix binom(r, m) d^(r-m), r >= m, then the tail coefficients d^(P-m) / ((P-m)! m!) of the column
        // that was the previous interval's own one
        const double* cs = s_cs + (J - 1) * OSP_STEP;
        double yn[P];
        int t = 0;
#pragma unroll
        for (int m = 0; m < P; ++m) {
          double acc = v[q] * cs[OSP_STEP - OSP_MAXP + m];
#pragma unroll
          for (int r = m; r < P; ++r, ++t) acc = fma(cs[t], y[r], acc);
          yn[m] = acc;
        }
#pragma unroll
        for (int m = 0; m < P; ++m) {
          y[m] = yn[m];
          Yt[((size_t)(gbase + J) * P + m) * ldk + k] = yn[m];
        }
      }
    }
  }
}

template <int P, int NDC>
__global__ void __launch_bounds__(128) osp_levOmega_kernel(const double* __restrict__ V, int p, int ldl, int nD, int ldk,
                                                           const int* __restrict__ g_own, const double* __restrict__ Yt,
                                                           double* __restrict__ Omega) {
  constexpr int NV = P + 1 + NDC, NT = NV * (NV + 1) / 2;
  __shared__ double sm[4][NT];
  const int gid = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int own = g_own[gid];
  double acc[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t] = 0.0;
  for (int k = threadIdx.x; k < p; k += 128) {
    double v[NV];
#pragma unroll
    for (int m = 0; m < P; ++m) v[m] = Yt[((size_t)gid * P + m) * ldk + k];
    v[P] = own >= 0 ? V[(size_t)k * ldl + nD + own] * osp_ifact(P) : 0.0;
#pragma unroll
    for (int c = 0; c < NDC; ++c) v[P + 1 + c] = c < nD ? V[(size_t)k * ldl + c] : 0.0;
    int t = 0;
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b, ++t) acc[t] = fma(v[a], v[b], acc[t]);
  }
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const double s = osp_warp_sum(acc[t]);
    if (lane == 0) sm[warp][t] = s;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NT; t += 128) {
    const double s = (sm[0][t] + sm[1][t]) + (sm[2][t] + sm[3][t]);
    // t -> (a, b), b <= a; both triangles
    int a = 0;
    while ((a + 1) * (a + 2) / 2 <= t) ++a;
    const int b = t - a * (a + 1) / 2;
    Omega[(size_t)gid * NV * NV + a * NV + b] = s;
    Omega[(size_t)gid * NV * NV + b * NV + a] = s;
  }
}

struct OspLevArgs {
  const double *u, *size, *D, *eta;
  const int64_t* piece_beg;
  const int* piece_gid;
  const double* Omega;
  int nD, np, family;
  int64_t n;
  double* slots;
};

template <int P, int NDC>
__global__ void __launch_bounds__(128) osp_levpass_kernel(const OspLevArgs a) {
  constexpr int NV = P + 1 + NDC, OGD = osp_offgD(P, NDC), NACC = osp_NACC(P, NDC);
  __shared__ double sOm[4][NV * NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int pc = blockIdx.x * 4 + warp; pc < a.np; pc += gridDim.x * 4) {
    const int gid = a.piece_gid[pc];
    const int64_t j0 = a.piece_beg[pc], j1 = a.piece_beg[pc + 1];
    __syncwarp();
    for (int t = lane; t < NV * NV; t += 32) sOm[warp][t] = a.Omega[(size_t)gid * NV * NV + t];
    __syncwarp();
    double zr[P + 1], zd[NDC];
#pragma unroll
    for (int m = 0; m <= P; ++m) zr[m] = 0.0;
#pragma unroll
    for (int c = 0; c < NDC; ++c) zd[c] = 0.0;
    for (int64_t j = j0 + lane; j < j1; j += 32) {
      const double u = a.u[j], eta = a.eta[j];
      const double sz = a.size ? a.size[j] : 1.0;
      double v[NV];
      v[0] = 1.0;
#pragma unroll
      for (int m = 1; m <= P; ++m) v[m] = v[m - 1] * u;
#pragma unroll
      for (int c = 0; c < NDC; ++c) v[P + 1 + c] = c < a.nD ? a.D[(size_t)c * a.n + j] : 0.0;
      double q = 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k) t = fma(sOm[warp][i * NV + k], v[k], t);
        q = fma(v[i], t, q);
      }
      double c3;
      if (a.family == BGP_FAMILY_POISSON) {
        c3 = exp(eta);
      } else {                                   // Binomial: w (1 - 2 pi), as obs_terms
        const double e = exp(-fabs(eta));
        const double inv = 1.0 / (1.0 + e);
        const double pi = eta >= 0.0 ? inv : e * inv, om = eta >= 0.0 ? e * inv : inv;
        c3 = sz * pi * om * (om - pi);
      }
      const double z = c3 * q;
#pragma unroll
      for (int m = 0; m <= P; ++m) zr[m] = fma(z, v[m], zr[m]);
#pragma unroll
      for (int c = 0; c < NDC; ++c) zd[c] = fma(z, v[P + 1 + c], zd[c]);
    }
    double* slot = a.slots + (size_t)pc * NACC;
#pragma unroll
    for (int m = 0; m <= P; ++m) {
      const double s = osp_warp_sum(zr[m]);
      if (lane == 0) slot[m] = s;
    }
#pragma unroll
    for (int c = 0; c < NDC; ++c) {
      const double s = osp_warp_sum(zd[c]);
      if (lane == 0) slot[OGD + c] = s;
    }
  }
}

// ---- host ----------------------------------------------------------------------------------------------------------
// A lane shares the observations, the piece table and the knot tables with its parent and owns what a pass writes.
int osp_plan_clone_for_lane(const bgp_model* parent, bgp_model* lane) {
  const OspPlan* pp = (const OspPlan*)parent->osp_plan;
  lane->osp_plan = nullptr;
  if (!pp) return BGP_OK;
  OspPlan* pl = new OspPlan(*pp);
  pl->eta = pl->slots = pl->mom = pl->glob = pl->Hdb = pl->G = pl->Yt = pl->Omega = pl->c_step = nullptr;
  pl->done = nullptr;
  lane->osp_plan = pl;
  auto zalloc = [&](double** ptr, size_t count) -> int {
    BGP_CUDA(cudaMalloc(ptr, std::max<size_t>(1, count) * sizeof(double)));
    BGP_CUDA(cudaMemsetAsync(*ptr, 0, std::max<size_t>(1, count) * sizeof(double), lane->stream));
    return BGP_OK;
  };
  BGP_TRY(zalloc(&pl->eta, (size_t)pl->n));
  BGP_TRY(zalloc(&pl->slots, (size_t)std::max(1, pl->np) * pl->NACC));
  BGP_TRY(zalloc(&pl->mom, (size_t)pl->NG * pl->NM));
  BGP_TRY(zalloc(&pl->glob, (size_t)(pl->NACC - pl->NM)));
  BGP_TRY(zalloc(&pl->Hdb, (size_t)pl->NDC * pl->NC));
  BGP_TRY(zalloc(&pl->G, (size_t)pl->NC * OSP_MAXP));
  BGP_CUDA(cudaMalloc(&pl->done, sizeof(int)));
  BGP_CUDA(cudaMemsetAsync(pl->done, 0, sizeof(int), lane->stream));
  return BGP_OK;
}

void osp_plan_destroy_lane(bgp_model* lane) {
  OspPlan* pl = (OspPlan*)lane->osp_plan;
  if (!pl) return;
  for (void* ptr : {(void*)pl->eta, (void*)pl->slots, (void*)pl->mom, (void*)pl->glob, (void*)pl->Hdb, (void*)pl->G, (void*)pl->done,
                    (void*)pl->Yt, (void*)pl->Omega, (void*)pl->c_step})
    if (ptr) cudaFree(ptr);
  delete pl;
  lane->osp_plan = nullptr;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& v) {
  BGP_CUDA(cudaMalloc(dst, std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) BGP_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return BGP_OK;
}

template <int P, int NDC>
static void pass_launch(bgp_model* m, OspPlan* pl, const OspPassArgs& pa, const OspApplyArgs& aa, const PriorArgs& pr) {
  osp_pass_kernel<P, NDC><<<std::min((pl->np + 3) / 4, pl->pass_grid), 128, 0, m->stream>>>(pa);
  osp_reduce_kernel<<<(pl->NG + 3) / 4 + (pl->NACC - pl->NM), 128, 0, m->stream>>>(pl->slots, pl->NACC, pl->NM, pl->NG, pl->gid_pbeg,
                                                                                   pl->np, pl->mom, pl->glob);
  osp_apply_kernel<P, NDC><<<pl->NC + 1, 128, 0, m->stream>>>(aa, pr);
}
template <int P, int NDC>
static void hess_launch(bgp_model* m, OspPlan* pl, const OspHArgs& ha) {
  (void)pl;
  dim3 grid((ha.p + 255) / 256, ha.p);
  osp_hwrite_kernel<P, NDC><<<grid, 256, 0, m->stream>>>(ha);
}

#define OSP_DISPATCH(FN, ...)                                                   \
  do {                                                                          \
    const int key_ = pl->P * 16 + pl->NDC;                                      \
    switch (key_) {                                                             \
      case 1 * 16 + 4: FN<1, 4>(__VA_ARGS__); break;                            \
      case 2 * 16 + 4: FN<2, 4>(__VA_ARGS__); break;                            \
      case 3 * 16 + 4: FN<3, 4>(__VA_ARGS__); break;                            \
      case 4 * 16 + 4: FN<4, 4>(__VA_ARGS__); break;                            \
      case 1 * 16 + 8: FN<1, 8>(__VA_ARGS__); break;                            \
      case 2 * 16 + 8: FN<2, 8>(__VA_ARGS__); break;                            \
      case 3 * 16 + 8: FN<3, 8>(__VA_ARGS__); break;                            \
      case 4 * 16 + 8: FN<4, 8>(__VA_ARGS__); break;                            \
      default: set_error("O-spline path: unsupported order / dense width"); return BGP_ERR_ARG; \
    }                                                                           \
  } while (0)

// theta != NULL (single-device models): the prior completion of finish.cu runs inside the last kernel as well
int osp_launch_lik(bgp_model* m, const double* W_dev, double tau, const double* theta) {
  OspPlan* pl = (OspPlan*)m->osp_plan;
  PriorArgs pr;
  memset(&pr, 0, sizeof(pr));
  if (theta) fill_prior_args(m, W_dev, theta, tau, &pr);
  OspPassArgs pa;
  pa.u = pl->u;
  pa.y = pl->y;
  pa.size = pl->size;
  pa.D = pl->D;
  pa.eta = pl->eta;
  pa.piece_beg = pl->piece_beg;
  pa.piece_gid = pl->piece_gid;
  pa.g_own = pl->g_own;
  pa.g_c0 = pl->g_c0;
  pa.g_nt = pl->g_nt;
  pa.g_t0 = pl->g_t0;
  pa.c_t0 = pl->c_t0;
  pa.c_t1 = pl->c_t1;
  pa.W = W_dev;
  pa.nD = pl->nD;
  pa.np = pl->np;
  pa.family = m->family;
  pa.n = pl->n;
  pa.tau = tau;
  pa.slots = pl->slots;
  OspApplyArgs aa;
  aa.NC = pl->NC;
  aa.nD = pl->nD;
  aa.NM = pl->NM;
  aa.lda = m->lda;
  aa.c_t0 = pl->c_t0;
  aa.c_t1 = pl->c_t1;
  aa.g_t0 = pl->g_t0;
  aa.c_gid = pl->c_gid;
  aa.c_gend = pl->c_gend;
  aa.mom = pl->mom;
  aa.glob = pl->glob;
  aa.red = m->red_buf;
  aa.Hdb = pl->Hdb;
  aa.G = pl->G;
  aa.done = pl->done;
  aa.fuse_prior = theta ? 1 : 0;
  OSP_DISPATCH(pass_launch, m, pl, pa, aa, pr);
  count_launch(3);
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

// theta != NULL: Q(theta) is added in the same kernel (single-device models)
int osp_launch_hessian(bgp_model* m, const double* theta) {
  OspPlan* pl = (OspPlan*)m->osp_plan;
  OspHArgs ha;
  ha.qfix = theta ? m->qfix : nullptr;
  ha.Pdiag = m->rnd[0].P_dev;
  ha.etheta = theta ? std::exp(theta[0]) : 0.0;
  ha.p = m->p;
  ha.ldh = m->ldh;
  ha.nD = pl->nD;
  ha.NC = pl->NC;
  ha.NM = pl->NM;
  ha.c_t0 = pl->c_t0;
  ha.c_t1 = pl->c_t1;
  ha.c_gid = pl->c_gid;
  ha.c_side = pl->c_side;
  ha.mom = pl->mom;
  ha.glob = pl->glob;
  ha.Hdb = pl->Hdb;
  ha.G = pl->G;
  ha.H = m->H;
  ha.L = theta ? m->L : nullptr;           // complete (Q included) only on a single device
  OSP_DISPATCH(hess_launch, m, pl, ha);
  m->L_holds_H = theta != nullptr;
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}


template <int P, int NDC>
static void lev_launch(bgp_model* m, OspPlan* pl, const double* V, int ldl, const OspLevArgs& la, const OspApplyArgs& aa) {
  const int p = m->p;
  dim3 gy((p + 127) / 128, pl->nside);
  const size_t ysm = (size_t)pl->maxK * OSP_STEP * sizeof(double);
  if (ysm > 48 * 1024) cudaFuncSetAttribute(osp_levY_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ysm);
  osp_levY_kernel<P><<<gy, 128, ysm, m->stream>>>(V, p, ldl, pl->nD, pl->ldk, pl->nside, pl->side_gbase, pl->side_cbase, pl->side_K,
                                               pl->c_step, pl->Yt);
  osp_levOmega_kernel<P, NDC><<<pl->NG, 128, 0, m->stream>>>(V, p, ldl, pl->nD, pl->ldk, pl->g_own, pl->Yt, pl->Omega);
  osp_levpass_kernel<P, NDC><<<std::min((pl->np + 3) / 4, pl->pass_grid), 128, 0, m->stream>>>(la);
  osp_reduce_kernel<<<(pl->NG + 3) / 4 + (pl->NACC - pl->NM), 128, 0, m->stream>>>(pl->slots, pl->NACC, pl->NM, pl->NG, pl->gid_pbeg,
                                                                                   pl->np, pl->mom, pl->glob);
  PriorArgs pr;
  memset(&pr, 0, sizeof(pr));
  osp_apply_kernel<P, NDC><<<pl->NC + 1, 128, 0, m->stream>>>(aa, pr);
}

// A^T (c3 * q) into red_buf[0 .. lda) from the moments (Poisson / Binomial; the eta of the last moment pass must be the
// mode's).  V: upper factor of H^-1 = V^T V, row-major p x ldl (grad.cu).
int osp_launch_leverage(bgp_model* m, const double* V, int ldl) {
  OspPlan* pl = (OspPlan*)m->osp_plan;
  if (!pl->Yt) {
    pl->ldk = round_up(m->p, 32);
    const int NV = pl->P + 1 + pl->NDC;
    BGP_CUDA(cudaMalloc(&pl->Yt, (size_t)pl->NG * pl->P * pl->ldk * sizeof(double)));
    BGP_CUDA(cudaMalloc(&pl->Omega, (size_t)pl->NG * NV * NV * sizeof(double)));
    // per column: shift matrix entries binom(r, m) d^(r-m) for m <= r < P (row by row), then d^(P-m) / ((P-m)! m!)
    std::vector<double> cs((size_t)pl->NC * OSP_STEP, 0.0);
    for (int col = 0; col < pl->NC; ++col) {
      const double d = pl->c_t1_host[(size_t)col] - pl->c_t0_host[(size_t)col];
      double dp[OSP_MAXP + 1];
      dp[0] = 1.0;
      for (int l = 1; l <= pl->P; ++l) dp[l] = dp[l - 1] * d;
      int t = 0;
      for (int mm = 0; mm < pl->P; ++mm) {
        double binom = 1.0;
        for (int r = mm; r < pl->P; ++r, ++t) {
          cs[(size_t)col * OSP_STEP + t] = binom * dp[r - mm];
          binom = binom * (double)(r + 1) / (double)(r + 1 - mm);
        }
        cs[(size_t)col * OSP_STEP + OSP_STEP - OSP_MAXP + mm] = dp[pl->P - mm] * (osp_ifact(pl->P - mm) * osp_ifact(mm));
      }
    }
    BGP_TRY(upload(&pl->c_step, cs));
  }
  OspLevArgs la;
  la.u = pl->u;
  la.size = pl->size;
  la.D = pl->D;
  la.eta = pl->eta;
  la.piece_beg = pl->piece_beg;
  la.piece_gid = pl->piece_gid;
  la.Omega = pl->Omega;
  la.nD = pl->nD;
  la.np = pl->np;
  la.family = m->family;
  la.n = pl->n;
  la.slots = pl->slots;
  OspApplyArgs aa;
  aa.NC = pl->NC;
  aa.nD = pl->nD;
  aa.NM = pl->NM;
  aa.lda = m->lda;
  aa.c_t0 = pl->c_t0;
  aa.c_t1 = pl->c_t1;
  aa.g_t0 = pl->g_t0;
  aa.c_gid = pl->c_gid;
  aa.c_gend = pl->c_gend;
  aa.mom = pl->mom;
  aa.glob = pl->glob;
  aa.red = m->red_buf;
  aa.Hdb = pl->Hdb;
  aa.G = pl->G;
  aa.done = pl->done;
  aa.fuse_prior = 0;
  OSP_DISPATCH(lev_launch, m, pl, V, ldl, la, aa);
  count_launch(5);
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

void osp_plan_destroy(bgp_model* m) {
  OspPlan* pl = (OspPlan*)m->osp_plan;
  if (!pl) return;
  for (void* ptr : {(void*)pl->u, (void*)pl->y, (void*)pl->size, (void*)pl->D, (void*)pl->eta, (void*)pl->piece_beg,
                    (void*)pl->piece_gid, (void*)pl->gid_pbeg, (void*)pl->g_t0, (void*)pl->g_own, (void*)pl->g_c0, (void*)pl->g_nt,
                    (void*)pl->c_t0, (void*)pl->c_t1, (void*)pl->c_gid, (void*)pl->c_gend, (void*)pl->c_side, (void*)pl->done, (void*)pl->side_gbase, (void*)pl->side_cbase, (void*)pl->side_K, (void*)pl->Yt,
                    (void*)pl->Omega, (void*)pl->c_step,
                    (void*)pl->slots, (void*)pl->mom, (void*)pl->glob, (void*)pl->Hdb, (void*)pl->G})
    if (ptr) cudaFree(ptr);
  delete pl;
  m->osp_plan = nullptr;
  m->osp_on = false;
}

// Called by bgp_model_finalize while the staged (column-major, caller-order) blocks are still alive.
// Eligible: exactly one smoothing term, built by bgp_model_add_iwp, order <= 4, at most 8 dense columns.
int osp_plan_create(bgp_model* m) {
  if (const char* e = getenv("BGP_NO_OSPLINE"))
    if (e[0] == '1') return BGP_OK;
  if (m->rnd.size() != 1 || m->iwp_terms.size() != 1 || m->st_rnd.size() != 1) return BGP_OK;
  const bgp_model::IwpTerm& it = m->iwp_terms[0];
  if (it.order > OSP_MAXP || m->nD > OSP_MAXD || !it.x_dev) return BGP_OK;
  const int64_t n = m->n;
  if (n >= ((int64_t)1 << 31)) return BGP_OK;
  for (const std::vector<double>* t : {&it.kneg, &it.kpos})        // the interval search needs increasing knots
    for (size_t i = 1; i < t->size(); ++i)
      if (!((*t)[i] > (*t)[i - 1])) return BGP_OK;
  OspPlan* pl = new OspPlan;
  m->osp_plan = pl;
  pl->P = it.order;
  pl->nD = m->nD;
  pl->NDC = m->nD <= 4 ? 4 : 8;
  pl->n = n;
  pl->NM = osp_NM(pl->P, pl->NDC);
  pl->NACC = osp_NACC(pl->P, pl->NDC);
  const int nkn = (int)it.kneg.size(), nkp = (int)it.kpos.size();
  const int Kn = nkn > 0 ? nkn - 1 : 0, Kp = nkp > 0 ? nkp - 1 : 0;
  const int gbase_pos = nkn > 0 ? Kn + 1 : 0;
  pl->NC = Kn + Kp;
  pl->NG = gbase_pos + (nkp > 0 ? Kp + 1 : 0);
  std::vector<double> g_t0((size_t)pl->NG), c_t0((size_t)pl->NC), c_t1((size_t)pl->NC);
  std::vector<int> g_own((size_t)pl->NG), g_c0((size_t)pl->NG), g_nt((size_t)pl->NG), c_gid((size_t)pl->NC), c_gend((size_t)pl->NC),
      c_side((size_t)pl->NC);
  auto fill_side = [&](const std::vector<double>& t, int K, int gbase, int cbase, int side) {
    for (int J = 0; J <= K; ++J) {
      g_t0[(size_t)gbase + J] = t[(size_t)J];
      g_own[(size_t)gbase + J] = J < K ? cbase + J : -1;
      g_c0[(size_t)gbase + J] = cbase;
      g_nt[(size_t)gbase + J] = J;
    }
    for (int i = 0; i < K; ++i) {
      c_t0[(size_t)cbase + i] = t[(size_t)i];
      c_t1[(size_t)cbase + i] = t[(size_t)i + 1];
      c_gid[(size_t)cbase + i] = gbase + i;
      c_gend[(size_t)cbase + i] = gbase + K + 1;
      c_side[(size_t)cbase + i] = side;
    }
  };
  std::vector<int> side_gbase, side_cbase, side_K;
  if (nkn > 0) {
    fill_side(it.kneg, Kn, 0, 0, 0);
    side_gbase.push_back(0);
    side_cbase.push_back(0);
    side_K.push_back(Kn);
  }
  if (nkp > 0) {
    fill_side(it.kpos, Kp, gbase_pos, Kn, 1);
    side_gbase.push_back(gbase_pos);
    side_cbase.push_back(Kn);
    side_K.push_back(Kp);
  }
  pl->nside = (int)side_K.size();
  pl->maxK = std::max(Kn, Kp);

  uint32_t *gid = nullptr, *gid2 = nullptr, *idx = nullptr, *idx2 = nullptr;
  double *u0 = nullptr, *kn_dev = nullptr, *kp_dev = nullptr;
  int* cnt_dev = nullptr;
  void* tmp = nullptr;
  int st = [&]() -> int {
    BGP_CUDA(cudaMalloc(&gid, n * sizeof(uint32_t)));
    BGP_CUDA(cudaMalloc(&gid2, n * sizeof(uint32_t)));
    BGP_CUDA(cudaMalloc(&idx, n * sizeof(uint32_t)));
    BGP_CUDA(cudaMalloc(&idx2, n * sizeof(uint32_t)));
    BGP_CUDA(cudaMalloc(&u0, n * sizeof(double)));
    BGP_TRY(upload(&kn_dev, it.kneg));
    BGP_TRY(upload(&kp_dev, it.kpos));
    OspLocateArgs la;
    la.x = it.x_dev;
    la.n = n;
    la.x0 = it.x0;
    la.kneg = kn_dev;
    la.kpos = kp_dev;
    la.nkn = nkn;
    la.nkp = nkp;
    la.gbase_pos = gbase_pos;
    la.gid = gid;
    la.idx = idx;
    la.u = u0;
    const unsigned vb = (unsigned)((n + 255) / 256);
    osp_locate_kernel<<<vb, 256, 0, m->stream>>>(la);
    count_launch();
    BGP_CUDA(cudaGetLastError());
    size_t tb = 0;
    int end_bit = 1;
    while ((1 << end_bit) < pl->NG) ++end_bit;
    BGP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, gid, gid2, idx, idx2, (int64_t)n, 0, end_bit, m->stream));
    BGP_CUDA(cudaMalloc(&tmp, tb));
    BGP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, gid, gid2, idx, idx2, (int64_t)n, 0, end_bit, m->stream));
    count_launch(4);
    BGP_CUDA(cudaMalloc(&cnt_dev, (size_t)pl->NG * sizeof(int)));
    BGP_CUDA(cudaMemsetAsync(cnt_dev, 0, (size_t)pl->NG * sizeof(int), m->stream));
    osp_count_kernel<<<vb, 256, 0, m->stream>>>(gid, n, cnt_dev);
    count_launch();
    // observations into interval order
    BGP_CUDA(cudaMalloc(&pl->u, n * sizeof(double)));
    BGP_CUDA(cudaMalloc(&pl->y, n * sizeof(double)));
    BGP_CUDA(cudaMalloc(&pl->eta, n * sizeof(double)));
    BGP_CUDA(cudaMemsetAsync(pl->eta, 0, n * sizeof(double), m->stream));
    osp_gather_kernel<<<vb, 256, 0, m->stream>>>(u0, idx2, n, pl->u);
    osp_gather_kernel<<<vb, 256, 0, m->stream>>>(m->y, idx2, n, pl->y);
    count_launch(2);
    if (m->size && m->family == BGP_FAMILY_BINOMIAL) {
      BGP_CUDA(cudaMalloc(&pl->size, n * sizeof(double)));
      osp_gather_kernel<<<vb, 256, 0, m->stream>>>(m->size, idx2, n, pl->size);
      count_launch();
    }
    BGP_CUDA(cudaMalloc(&pl->D, std::max<size_t>(1, (size_t)pl->nD) * n * sizeof(double)));
    {
      // dense columns in the internal order: boundary blocks, then fixed blocks (bgp_internal.h)
      int c = 0;
      for (auto* v : {&m->st_bnd, &m->st_fix})
        for (auto& s : *v)
          for (int k = 0; k < s.ncol; ++k, ++c) {
            osp_gather_kernel<<<vb, 256, 0, m->stream>>>(s.dev + (size_t)k * n, idx2, n, pl->D + (size_t)c * n);
            count_launch();
          }
      if (c != pl->nD) {
        set_error("O-spline path: dense column count mismatch (%d vs %d)", c, pl->nD);
        return BGP_ERR_ARG;
      }
    }
    BGP_CUDA(cudaGetLastError());
    std::vector<int> cnt((size_t)pl->NG);
    BGP_CUDA(cudaMemcpyAsync(cnt.data(), cnt_dev, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    BGP_CUDA(cudaStreamSynchronize(m->stream));
    // pieces: every interval is cut into equal runs of at most PL observations; PL is the smallest length for which
    // the pieces fit two rounds of the pass kernel's resident warps (no ragged last wave)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    pl->pass_grid = sms * (pl->NDC <= 4 ? 3 : 2);
    const int64_t target = (int64_t)pl->pass_grid * 4 * 2;
    auto count_pieces = [&](int64_t len) {
      int64_t c = 0;
      for (int g = 0; g < pl->NG; ++g) c += (cnt[(size_t)g] + len - 1) / len;
      return c;
    };
    int64_t PL = std::max<int64_t>(64, n / target);
    while (PL < n && count_pieces(PL) > target) PL += std::max<int64_t>(1, PL / 32);
    if (const char* e = getenv("BGP_OSP_PIECE")) PL = std::max(32, atoi(e));
    std::vector<int64_t> piece_beg;
    std::vector<int> piece_gid, gid_pbeg((size_t)pl->NG + 1);
    int64_t off = 0;
    for (int g = 0; g < pl->NG; ++g) {
      gid_pbeg[(size_t)g] = (int)piece_gid.size();
      const int64_t c = cnt[(size_t)g], k = (c + PL - 1) / PL;
      for (int64_t b = 0; b < k; ++b) {
        piece_beg.push_back(off + c * b / k);
        piece_gid.push_back(g);
      }
      off += c;
    }
    gid_pbeg[(size_t)pl->NG] = (int)piece_gid.size();
    piece_beg.push_back(off);
    // a piece ends where the next begins, except across intervals — consecutive by construction (sorted order)
    pl->np = (int)piece_gid.size();
    if (off != n) {
      set_error("O-spline path: interval counts do not add up");
      return BGP_ERR_ARG;
    }
    BGP_TRY(upload(&pl->piece_beg, piece_beg));
    BGP_TRY(upload(&pl->piece_gid, piece_gid));
    BGP_TRY(upload(&pl->gid_pbeg, gid_pbeg));
    BGP_TRY(upload(&pl->g_t0, g_t0));
    BGP_TRY(upload(&pl->g_own, g_own));
    BGP_TRY(upload(&pl->g_c0, g_c0));
    BGP_TRY(upload(&pl->g_nt, g_nt));
    BGP_TRY(upload(&pl->c_t0, c_t0));
    pl->c_t0_host = c_t0;
    pl->c_t1_host = c_t1;
    BGP_TRY(upload(&pl->c_t1, c_t1));
    BGP_TRY(upload(&pl->c_gid, c_gid));
    BGP_TRY(upload(&pl->c_gend, c_gend));
    BGP_TRY(upload(&pl->c_side, c_side));
    BGP_TRY(upload(&pl->side_gbase, side_gbase));
    BGP_TRY(upload(&pl->side_cbase, side_cbase));
    BGP_TRY(upload(&pl->side_K, side_K));
    auto zalloc = [&](double** ptr, size_t count) -> int {
      BGP_CUDA(cudaMalloc(ptr, std::max<size_t>(1, count) * sizeof(double)));
      BGP_CUDA(cudaMemsetAsync(*ptr, 0, std::max<size_t>(1, count) * sizeof(double), m->stream));
      return BGP_OK;
    };
    BGP_CUDA(cudaMalloc(&pl->done, sizeof(int)));
    BGP_CUDA(cudaMemsetAsync(pl->done, 0, sizeof(int), m->stream));
    BGP_TRY(zalloc(&pl->slots, (size_t)std::max(1, pl->np) * pl->NACC));
    BGP_TRY(zalloc(&pl->mom, (size_t)pl->NG * pl->NM));
    BGP_TRY(zalloc(&pl->glob, (size_t)(pl->NACC - pl->NM)));
    BGP_TRY(zalloc(&pl->Hdb, (size_t)pl->NDC * pl->NC));
    BGP_TRY(zalloc(&pl->G, (size_t)pl->NC * OSP_MAXP));
    return BGP_OK;
  }();
  for (void* p : {(void*)gid, (void*)gid2, (void*)idx, (void*)idx2, (void*)u0, (void*)kn_dev, (void*)kp_dev, (void*)cnt_dev, tmp})
    if (p) cudaFree(p);
  if (st != BGP_OK) {
    osp_plan_destroy(m);
    return st;
  }
  m->osp_on = true;
  if (!getenv("BGP_LANES")) m->n_lanes = 4;     // measured on C3 (scripts/lanes_probe.sh): 1 / 2 / 3 / 4 / 6 / 8 lanes -> 2480 / 3898 / 5053 / 5533 / 5298 / 5079 evals/s
  return BGP_OK;
}

}  // namespace bgp
