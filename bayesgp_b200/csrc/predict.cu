// predict.cu — sample -> function evaluation and its summary.
//
// Replaces compute_post_fun_IWP (/root/reference/R/03_post_fit.R:200-241), compute_post_fun_sGP
// (:261-276), extract_mean_interval_given_samps (:287-296, stats::quantile type 7) and the basis
// evaluators they call (local_poly_helper / global_poly_helper / Compute_B_sB_helper /
// global_poly_helper_sGP, /root/reference/R/01_utility.R:198-208,378-440).
//
//   F (G x M) = [X_deg(x) | B(x)] [global rows; coef]          fitted_samps_deriv, :235 / :272
// is never materialised: rows are processed in strips; per strip the design rows are generated on
// the device, multiplied on the FP64 tensor pipe (kgemm.cu) into an L2-sized scratch strip, and each
// row is reduced to (mean, two type-7 quantiles) by an exact selection (one CTA per row: candidate cuts
// + value buckets, MSB radix select as the assumption-free fallback).
#include <algorithm>

#include "basis_dev.cuh"
#include <cstdlib>
#include <mutex>

#include "bgp_internal.h"

namespace bgp {

// ---- design rows -----------------------------------------------------------------------------------
struct IwpDesignArgs {
  const double* x;      // refined_x (already shifted by initial_location)
  int64_t g0, rows;
  const double* kneg;
  int nneg;
  const double* kpos;
  int npos;
  int order, degree;
  double* D;            // rows x ld, row-major
  int ld;
};

__global__ void iwp_design_kernel(const IwpDesignArgs a) {
  const int64_t r = blockIdx.x;
  if (r >= a.rows) return;
  const double x = a.x[a.g0 + r];
  const int q = a.order - a.degree;            // order of the basis after `degree` derivatives
  const int nX = q;                            // global polynomial columns kept (R/03_post_fit.R:230-234)
  const int nB = (a.nneg > 0 ? a.nneg - 1 : 0) + (a.npos > 0 ? a.npos - 1 : 0);
  const double xn = x < 0.0 ? -x : 0.0, xp = x > 0.0 ? x : 0.0;
  double* row = a.D + (size_t)r * a.ld;
  for (int c = threadIdx.x; c < a.ld; c += blockDim.x) {
    double v = 0.0;
    if (c < nX) {
      // column i = c + 1: x^(i-1) * (i + degree - 1)! / (i - 1)!
      double f = 1.0;
      for (int t = c + 1; t <= c + a.degree; ++t) f *= (double)t;
      v = f * ipow(x, c);
    } else if (c < nX + nB) {
      int j = c - nX;
      const int n1 = a.nneg > 0 ? a.nneg - 1 : 0;
      if (j < n1) {
        v = iwp_phi(xn, a.kneg[j], a.kneg[j + 1], q);
      } else {
        j -= n1;
        v = iwp_phi(a.nneg > 0 ? xp : x, a.kpos[j], a.kpos[j + 1], q);
      }
    }
    row[c] = v;
  }
}

struct SgpDesignArgs {
  const double* x;      // refined_x
  int64_t g0, rows;
  double x0;            // min(refined_x): Compute_B_sB_helper(initial_location = NULL) quirk
  double a;
  int k, m, boundary;
  double lo, hi;
  double* D;
  int ld;
};

__global__ void sgp_design_kernel(const SgpDesignArgs a) {
  const int64_t r = blockIdx.x;
  if (r >= a.rows) return;
  const double x = a.x[a.g0 + r] - a.x0;
  const int nb = a.boundary ? a.k - 2 : a.k;
  const int drop = a.boundary ? 2 : 0;
  double* row = a.D + (size_t)r * a.ld;
  int first = 0;
  double v4[4] = {0.0, 0.0, 0.0, 0.0};
  const bool inside = x >= a.lo && x <= a.hi;
  if (inside) bspline4(x, a.lo, a.hi, a.k - 2, first, v4);
  const int nX = 1 + 2 * a.m;
  for (int c = threadIdx.x; c < a.ld; c += blockDim.x) {
    double v = 0.0;
    if (c == 0) {
      v = 1.0;
    } else if (c < nX) {
      const int i = (c - 1) / 2 + 1;
      v = ((c - 1) & 1) ? sin(i * a.a * x) : cos(i * a.a * x);
    } else if (c < nX + 3 * nb * a.m) {
      const int e = c - nX;
      const int harm = e / (3 * nb) + 1;
      const int part = (e % (3 * nb)) / nb;          // 0: B cos, 1: B sin, 2: B
      const int bi = (e % nb) + drop;                  // index in the full k-function basis
      double bv = 0.0;
      if (inside && bi >= first && bi < first + 4) bv = v4[bi - first];
      if (part == 0) v = bv * cos(harm * a.a * x);
      else if (part == 1) v = bv * sin(harm * a.a * x);
      else v = bv;
    }
    row[c] = v;
  }
}

// ---- coefficient matrix C[s][c] (M x ld, K-major) from R-layout sample blocks --------------------------
struct CoefArgs {
  const double* icpt;     // M or NULL
  const double* glob;     // nglob x M column-major or NULL
  int nglob;
  const double* coef;     // ncoef x M column-major
  int ncoef;
  int skip;               // rows of rbind(icpt, glob) dropped from the top (= degree for IWP, 0 for sGP)
  int nX;                 // rows of rbind(icpt, glob) kept
  int64_t M;
  double* C;
  int ld;
};

__global__ void build_coef_kernel(const CoefArgs a) {
  const int64_t s = blockIdx.x;
  double* row = a.C + (size_t)s * a.ld;
  for (int c = threadIdx.x; c < a.ld; c += blockDim.x) {
    double v = 0.0;
    if (c < a.nX) {
      const int r = c + a.skip;          // row of rbind(intercept_samps, global_samps)
      if (r == 0) v = a.icpt ? a.icpt[s] : 0.0;
      else v = (a.glob && r - 1 < a.nglob) ? a.glob[(size_t)s * a.nglob + (r - 1)] : 0.0;
    } else if (c < a.nX + a.ncoef) {
      v = a.coef[(size_t)s * a.ncoef + (c - a.nX)];
    }
    row[c] = v;
  }
}

// ---- per-row mean and type-7 quantiles ------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dkey(double x) {
  const long long b = __double_as_longlong(x);
  return (unsigned long long)b ^ ((unsigned long long)(b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
  return __longlong_as_double((long long)b);
}

constexpr int RS_THREADS = 256;
constexpr int RS_LIST = 128;      // keys of one target bucket, ranked by a single warp
constexpr int RS_OVF = 256;       // shared overflow list behind the per-thread candidate slots
constexpr int RS_PC_MIN = 8;      // the slot area doubles as the per-warp histograms of the radix path (16 KB)

struct SelectArgs {
  const double* F;
  int64_t ldF, M;
  int64_t g0;
  int64_t r1, r2;     // 0-based ranks floor(index) - 1 of the two probabilities
  double h1, h2;      // interpolation weights (0 => no second order statistic needed)
  double z_lo, z_hi;  // candidate cuts in standard deviations from the row mean (fast path)
  int pc;             // candidate slots per thread (0 => radix path only)
  int nb;             // value buckets per tail (power of two, multiple of 128)
  double* mean;
  double* lo;
  double* hi;
};

// One CTA per row, four steps, each a few instructions per value:
//   A  stream the row once (128-bit loads, eight in flight per thread): sum, sum of squares, copy to shared memory;
//   B  every thread re-reads its own values and keeps those beyond mean +- z sigma (about 1.8 x the wanted tail) in
//      private slots — no ballots or atomics; the rare thread with more hits than slots spills to a shared list;
//   C  the kept values are counted into nb buckets per tail by distance from the cut (a monotone map, so bucket
//      order is value order); a scan from the extreme end finds the bucket holding each wanted order statistic and
//      the exact number of values beyond it;
//   D  the handful of values of that bucket are collected and ranked by one warp.
// Every count is exact, so the result is the exact order statistic.  Anything unexpected — cuts that miss the rank
// (heavy tails, constant rows), NaN / Inf, a crowded bucket, a full overflow list — drops the whole row to the MSB
// radix select below, which needs no assumption about the values.
template <bool IN_SMEM>
__global__ void __launch_bounds__(RS_THREADS, 2) row_select_kernel(const SelectArgs a) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  constexpr int NW = RS_THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t M = a.M;
  const int64_t Mpad = (M + 1) & ~(int64_t)1;
  const int pc = a.pc, nb = a.nb;
  double* vals = reinterpret_cast<double*>(rs_smem);
  double* priv = vals + (IN_SMEM ? Mpad : 0);                                  // [slot][thread]
  double* ovf = priv + (size_t)(pc > RS_PC_MIN ? pc : RS_PC_MIN) * RS_THREADS;
  unsigned int* hist = reinterpret_cast<unsigned int*>(ovf + RS_OVF);          // [2][nb]
  double* lists = reinterpret_cast<double*>(hist + 2 * nb);                    // [2 tails][2 ranks][RS_LIST]
  __shared__ double s_red[2 * NW];
  __shared__ unsigned long long s_kmin[NW], s_kmax[NW];
  __shared__ unsigned int s_novf, s_wtot[NW], s_n[2], s_lcnt[2][2];
  __shared__ int s_bsel[2][2];
  __shared__ unsigned int s_before[2][2];
  __shared__ double s_val[2][2];
  __shared__ unsigned long long s_prefix[2];
  __shared__ long long s_below[2];
  __shared__ unsigned int s_eq[2];
  const double* row = a.F + (size_t)blockIdx.x * a.ldF;
  const double2* row2 = reinterpret_cast<const double2*>(row);                 // rows are 16-byte aligned (ldF even)
  const int64_t npair = M >> 1;

  // ---- A: mean (fixed-order tree) and the row into shared memory ----------------------------------------------
  // (squares are taken about the first value of the row: the variance only places the cuts, but a row whose
  // spread is tiny against its level must not lose it to cancellation)
  const double shift = row[0];
  double sum = 0.0, sumsq = 0.0;
  for (int64_t p0 = tid; p0 < npair; p0 += 8 * RS_THREADS) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t p = p0 + (int64_t)u * RS_THREADS;
      v[u] = p < npair ? row2[p] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t p = p0 + (int64_t)u * RS_THREADS;
      if (p < npair) {
        const double dx = v[u].x - shift, dy = v[u].y - shift;
        sum += v[u].x;
        sumsq = fma(dx, dx, sumsq);
        sum += v[u].y;
        sumsq = fma(dy, dy, sumsq);
        if (IN_SMEM) reinterpret_cast<double2*>(vals)[p] = v[u];
      }
    }
  }
  if ((M & 1) && tid == 0) {
    const double v = row[M - 1];
    sum += v;
    sumsq = fma(v - shift, v - shift, sumsq);
    if (IN_SMEM) vals[M - 1] = v;
  }
  for (int t = tid; t < 2 * nb; t += RS_THREADS) hist[t] = 0;
  if (tid < 4) {
    s_lcnt[tid >> 1][tid & 1] = 0;
    s_bsel[tid >> 1][tid & 1] = -1;
  }
  if (tid == 0) s_novf = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
  }
  if (lane == 0) {
    s_red[warp] = sum;
    s_red[NW + warp] = sumsq;
  }
  __syncthreads();
  sum = 0.0;
  sumsq = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    sum += s_red[w];
    sumsq += s_red[NW + w];
  }
  const double mean = sum / (double)M;
  double q_lo = 0.0, q_hi = 0.0;
  bool done = false;

  const double sd = sqrt(fmax(sumsq / (double)M - (mean - shift) * (mean - shift), 0.0));
  if (pc > 0 && sd > 0.0 && isfinite(sd) && isfinite(mean)) {                  // CTA-uniform
    const double cut_lo = mean + a.z_lo * sd, cut_hi = mean + a.z_hi * sd;
    const double scale = (double)nb / (3.0 * sd);                              // buckets span three sigma beyond a cut
    bool ok = true;
    // ---- B: candidates into private slots -------------------------------------------------------------------
    int c = 0;
    auto visit = [&](double v) {
      if (v < cut_lo || v > cut_hi) {
        if (c < pc) {
          priv[c * RS_THREADS + tid] = v;
        } else {
          const unsigned pos = atomicAdd(&s_novf, 1u);
          if (pos < (unsigned)RS_OVF) ovf[pos] = v;
        }
        ++c;
      }
    };
    if (IN_SMEM) {
      for (int64_t p = tid; p < npair; p += RS_THREADS) {
        const double2 v = reinterpret_cast<const double2*>(vals)[p];          // written by this thread in step A
        visit(v.x);
        visit(v.y);
      }
      if ((M & 1) && tid == 0) visit(vals[M - 1]);
    } else {
      for (int64_t p0 = tid; p0 < npair; p0 += 4 * RS_THREADS) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t p = p0 + (int64_t)u * RS_THREADS;
          v[u] = p < npair ? row2[p] : make_double2(mean, mean);               // the mean is never a candidate
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          visit(v[u].x);
          visit(v[u].y);
        }
      }
      if ((M & 1) && tid == 0) visit(row[M - 1]);
    }
    // ---- C: bucket counts, extreme end first ----------------------------------------------------------------
    // tail 0 = below cut_lo, tail 1 = above cut_hi; bucket nb - 1 is the far end.  v1 <= v2 implies
    // bucket(v1) >= bucket(v2) in tail 0 (and the mirror image in tail 1): subtraction, the scaling and the
    // floor are all monotone, so the buckets partition the candidates in value order.
    auto bucket_of = [&](double v, int& q) {
      q = v > cut_hi ? 1 : 0;
      const double d = q ? v - cut_hi : cut_lo - v;
      const int b = __double2int_rd(d * scale);                                // saturating; d > 0
      return b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
    };
    const int cp = c < pc ? c : pc;
    for (int j = 0; j < cp; ++j) {
      int q;
      const int b = bucket_of(priv[j * RS_THREADS + tid], q);
      atomicAdd(&hist[q * nb + b], 1u);
    }
    __syncthreads();
    const unsigned novf = s_novf;
    if (novf > (unsigned)RS_OVF) ok = false;
    const unsigned novf_c = novf > (unsigned)RS_OVF ? 0u : novf;
    for (unsigned i = tid; i < novf_c; i += RS_THREADS) {
      int q;
      const int b = bucket_of(ovf[i], q);
      atomicAdd(&hist[q * nb + b], 1u);
    }
    __syncthreads();
    // 128 threads per tail, nb / 128 consecutive buckets each, walking inwards from the far end
    const int q = tid >> 7, u = tid & 127, w = nb >> 7;
    const unsigned int* hq = hist + q * nb;
    const int btop = nb - 1 - u * w;
    unsigned tot = 0;
    for (int j = 0; j < w; ++j) tot += hq[btop - j];
    unsigned inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wtot[warp] = inc;
    __syncthreads();
    unsigned base = 0;
    for (int ww = q * 4; ww < warp; ++ww) base += s_wtot[ww];
    const long long cum0 = (long long)base + (long long)(inc - tot);         // candidates beyond this thread's buckets
    if (u == 127) s_n[q] = base + inc;
    // wanted positions counted from the extreme end: P0 = rank r (x_lo of the interpolation), P1 = rank r + 1
    const long long top2 = M - 1 - a.r2;
    const bool need_lo = a.h1 != 0.0, need_hi = a.h2 != 0.0;
    {
      const long long P0 = q == 0 ? a.r1 : top2, P1 = q == 0 ? a.r1 + 1 : top2 - 1;
      const bool need1 = q == 0 ? need_lo : need_hi;
#pragma unroll
      for (int wch = 0; wch < 2; ++wch) {
        if (wch == 1 && !need1) continue;
        const long long P = wch == 0 ? P0 : P1;
        if (P >= cum0 && P < cum0 + (long long)tot) {
          long long cum = cum0;
          for (int j = 0; j < w; ++j) {
            const unsigned cnt = hq[btop - j];
            if (P < cum + (long long)cnt) {
              s_bsel[q][wch] = btop - j;
              s_before[q][wch] = (unsigned)cum;
              break;
            }
            cum += cnt;
          }
        }
      }
    }
    __syncthreads();
    // the cuts must not have missed a wanted rank
    if ((need_lo ? a.r1 + 1 : a.r1) >= (long long)s_n[0]) ok = false;
    if (top2 >= (long long)s_n[1] || (need_hi && top2 < 1)) ok = false;
    // ---- D: collect the target buckets, rank inside them -------------------------------------------------------
    const int b00 = s_bsel[0][0], b01 = s_bsel[0][1], b10 = s_bsel[1][0], b11 = s_bsel[1][1];
    if (ok) {
      auto collect = [&](double v) {
        int qq;
        const int b = bucket_of(v, qq);
        const int t0 = qq ? b10 : b00, t1 = qq ? b11 : b01;
        if (b == t0) {
          const unsigned pos = atomicAdd(&s_lcnt[qq][0], 1u);
          if (pos < (unsigned)RS_LIST) lists[(qq * 2) * RS_LIST + pos] = v;
        } else if (b == t1) {
          const unsigned pos = atomicAdd(&s_lcnt[qq][1], 1u);
          if (pos < (unsigned)RS_LIST) lists[(qq * 2 + 1) * RS_LIST + pos] = v;
        }
      };
      for (int j = 0; j < cp; ++j) collect(priv[j * RS_THREADS + tid]);
      for (unsigned i = tid; i < novf_c; i += RS_THREADS) collect(ovf[i]);
    }
    __syncthreads();
    if (ok) {
      if (s_lcnt[0][0] > (unsigned)RS_LIST || s_lcnt[0][1] > (unsigned)RS_LIST || s_lcnt[1][0] > (unsigned)RS_LIST ||
          s_lcnt[1][1] > (unsigned)RS_LIST)
        ok = false;                                                            // crowded bucket
    }
    if (ok && warp < 4) {
      const int qq = warp >> 1, wch = warp & 1;
      if (wch == 0 || (qq == 0 ? need_lo : need_hi)) {
        const int bsel0 = qq ? b10 : b00, bsel1 = qq ? b11 : b01;
        const int src = (wch == 1 && bsel1 == bsel0) ? 0 : wch;
        const int n = (int)s_lcnt[qq][src];
        const double* L = lists + (qq * 2 + src) * RS_LIST;
        const long long P = qq == 0 ? (wch == 0 ? a.r1 : a.r1 + 1) : (wch == 0 ? top2 : top2 - 1);
        const long long t = P - (long long)s_before[qq][wch];                  // position inside the bucket
        for (int e = lane; e < n; e += 32) {
          const double x = L[e];
          int r = 0;
          for (int j = 0; j < n; ++j) {
            const double y = L[j];
            r += ((qq == 0 ? y < x : y > x) || (y == x && j < e)) ? 1 : 0;
          }
          if ((long long)r == t) s_val[qq][wch] = x;
        }
      }
    }
    __syncthreads();
    if (ok) {
      // R: qs <- x[lo]; where (index > lo & x[hi] != qs): qs <- (1 - h) * qs + h * x[hi]
      {
        const double x_lo = s_val[0][0];
        const double x_hi = need_lo ? s_val[0][1] : x_lo;
        q_lo = (need_lo && x_hi != x_lo) ? (1.0 - a.h1) * x_lo + a.h1 * x_hi : x_lo;
      }
      {
        const double x_lo = s_val[1][0];
        const double x_hi = need_hi ? s_val[1][1] : x_lo;
        q_hi = (need_hi && x_hi != x_lo) ? (1.0 - a.h2) * x_lo + a.h2 * x_hi : x_lo;
      }
      done = true;
    }
  }
  if (done) {
    if (tid == 0) {
      const int64_t g = a.g0 + blockIdx.x;
      if (a.mean) a.mean[g] = mean;
      if (a.lo) a.lo[g] = q_lo;
      if (a.hi) a.hi[g] = q_hi;
    }
    return;
  }

  // ---- fallback: exact MSB radix select (8-bit digits) of both order statistics in the same sweeps ---------------
  // per-warp private histograms fed by warp-aggregated increments (the keys of a row share their leading bytes, a
  // single shared histogram serialises on one bin), a warp-parallel bin scan, and a start digit chosen from the
  // highest bit in which the row's keys differ.  The histograms live in the candidate-slot area.
  __syncthreads();
  const int64_t rk[2] = {a.r1, a.r2};
  double qv[2] = {0.0, 0.0};
  unsigned int (*whist)[NW][256] = reinterpret_cast<unsigned int (*)[NW][256]>(priv);
  unsigned int (*rhist)[256] = reinterpret_cast<unsigned int (*)[256]>(hist);  // [2][256] (nb >= 256)
  auto key_at = [&](int64_t i) -> unsigned long long { return dkey(IN_SMEM ? vals[i] : row[i]); };
  unsigned long long kmin = ~0ull, kmax = 0ull;
  for (int64_t i = tid; i < M; i += RS_THREADS) {
    const unsigned long long k = key_at(i);
    kmin = k < kmin ? k : kmin;
    kmax = k > kmax ? k : kmax;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long mn = __shfl_xor_sync(0xffffffffu, kmin, o), mx = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmin = mn < kmin ? mn : kmin;
    kmax = mx > kmax ? mx : kmax;
  }
  if (lane == 0) {
    s_kmin[warp] = kmin;
    s_kmax[warp] = kmax;
  }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    kmin = s_kmin[w] < kmin ? s_kmin[w] : kmin;
    kmax = s_kmax[w] > kmax ? s_kmax[w] : kmax;
  }
  __syncthreads();
  // all keys agree above byte `top`: start there
  const unsigned long long diff = kmin ^ kmax;
  const int top = diff ? (63 - __clzll((long long)diff)) / 8 : 0;
  const unsigned long long hi_mask = top == 7 ? 0ull : (~0ull << (8 * (top + 1)));
  unsigned long long prefix[2] = {kmin & hi_mask, kmin & hi_mask}, mask = hi_mask;
  long long below[2] = {0, 0};
  unsigned int eq[2] = {(unsigned)M, (unsigned)M};
  for (int pass = top; pass >= 0; --pass) {
    const int shift = pass * 8;
    const bool same = prefix[0] == prefix[1];          // both quantiles still in the same bucket: one histogram
    for (int t = tid; t < 2 * NW * 256; t += RS_THREADS) (&whist[0][0][0])[t] = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < M; i0 += RS_THREADS) {
      const int64_t i = i0 + tid;
      const bool in = i < M;
      const unsigned long long k = in ? key_at(i) : 0ull;
      const unsigned digit = (unsigned)(k >> shift) & 255u;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q == 1 && same) break;
        const bool hit = in && (k & mask) == prefix[q];
        const unsigned act = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const unsigned peers = __match_any_sync(act, digit);
          if (lane == __ffs(peers) - 1) whist[q][warp][digit] += __popc(peers);
        }
      }
    }
    __syncthreads();
    for (int t = tid; t < 2 * 256; t += RS_THREADS) {
      const int q = t >> 8, bin = t & 255;
      unsigned s = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += whist[same ? 0 : q][w][bin];
      rhist[q][bin] = s;
    }
    __syncthreads();
    if (warp < 2) {
      // warp q finds the bin of rank rk[q]: 8 bins per lane, exclusive scan across lanes
      const int q = warp;
      unsigned loc[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        loc[j] = rhist[q][lane * 8 + j];
        tot += loc[j];
      }
      unsigned inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      long long cum = below[q] + (long long)(inc - tot);
      const bool mine = cum <= rk[q] && rk[q] < cum + (long long)tot;
      if (mine) {
        int j = 0;
        for (; j < 7; ++j) {
          if (cum + (long long)loc[j] > rk[q]) break;
          cum += loc[j];
        }
        s_below[q] = cum;
        s_prefix[q] = prefix[q] | ((unsigned long long)(lane * 8 + j) << shift);
        s_eq[q] = loc[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      below[q] = s_below[q];
      prefix[q] = s_prefix[q];
      eq[q] = s_eq[q];
    }
    mask |= 0xFFull << shift;
    __syncthreads();
  }
  for (int which = 0; which < 2; ++which) {
    const int64_t r = rk[which];
    const double h = which == 0 ? a.h1 : a.h2;
    const double x_lo = dunkey(prefix[which]);
    double x_hi = x_lo;
    if (h != 0.0 && below[which] + (long long)eq[which] <= r + 1) {
      // next order statistic = smallest key strictly above
      unsigned long long mn = ~0ull;
      for (int64_t i = tid; i < M; i += RS_THREADS) {
        const unsigned long long k = key_at(i);
        if (k > prefix[which] && k < mn) mn = k;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, mn, o);
        mn = t < mn ? t : mn;
      }
      if (lane == 0) s_kmin[warp] = mn;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < NW; ++w) mn = s_kmin[w] < mn ? s_kmin[w] : mn;
      x_hi = dunkey(mn);
      __syncthreads();
    }
    // R: qs <- x[lo]; where (index > lo & x[hi] != qs): qs <- (1 - h) * qs + h * x[hi]
    qv[which] = (h != 0.0 && x_hi != x_lo) ? (1.0 - h) * x_lo + h * x_hi : x_lo;
  }
  if (tid == 0) {
    const int64_t g = a.g0 + blockIdx.x;
    if (a.mean) a.mean[g] = mean;
    if (a.lo) a.lo[g] = qv[0];
    if (a.hi) a.hi[g] = qv[1];
  }
}

// ---- host orchestration ---------------------------------------------------------------------------------
// Scratch buffers of the predict / basis entry points.  Every entry point is blocking (its stream is drained
// before it returns), so a released block can be handed to the next call as is: a small per-device free list
// replaces cudaMalloc / cudaFree (several milliseconds per call for the 100 MB strips).
struct ScratchPool {
  struct Block { size_t bytes; void* p; int dev; };
  std::vector<Block> free_blocks;
  size_t cached = 0;
  std::mutex mu;
  ~ScratchPool() {}   // blocks are returned to the driver at process exit
};
static ScratchPool g_scratch;

// device time of the last predict call on this thread (CUDA events on its stream): GEMM, select, whole call
static thread_local double g_pred_ms[3] = {0.0, 0.0, 0.0};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int dev = 0;
  ~DevBuf() {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_scratch.mu);
    if (g_scratch.free_blocks.size() < 32 && g_scratch.cached + bytes <= ((size_t)4 << 30)) {
      g_scratch.free_blocks.push_back({bytes, p, dev});
      g_scratch.cached += bytes;
    } else {
      cudaFree(p);
    }
  }
  int alloc(size_t want) {
    want = want ? want : 8;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lk(g_scratch.mu);
      int best = -1;
      for (int i = 0; i < (int)g_scratch.free_blocks.size(); ++i) {
        const auto& b = g_scratch.free_blocks[(size_t)i];
        if (b.dev == dev && b.bytes >= want && b.bytes <= 2 * want + 4096 &&
            (best < 0 || b.bytes < g_scratch.free_blocks[(size_t)best].bytes))
          best = i;
      }
      if (best >= 0) {
        p = g_scratch.free_blocks[(size_t)best].p;
        bytes = g_scratch.free_blocks[(size_t)best].bytes;
        g_scratch.cached -= bytes;
        g_scratch.free_blocks.erase(g_scratch.free_blocks.begin() + best);
        return BGP_OK;
      }
    }
    bytes = want;
    BGP_CUDA(cudaMalloc(&p, bytes));
    return BGP_OK;
  }
  template <class T>
  T* as() { return (T*)p; }
};

struct DesignSpec {
  bool iwp;
  // IWP
  std::vector<double> kneg, kpos;
  int order = 0, degree = 0;
  // sGP
  double a = 0, lo = 0, hi = 0, x0 = 0;
  int k = 0, m = 0, boundary = 1;
  int ncols = 0;
};

static int predict_core(const DesignSpec& ds, const double* Cmat_dev, int ldk, int64_t M, const double* x_host,
                        int64_t G, double level, cudaStream_t st, double* mean, double* plower, double* pupper,
                        double* samples) {
  DevBuf xb, knb, kpb, Db, Fb, ob, Sb;
  BGP_TRY(xb.alloc((size_t)G * sizeof(double)));
  BGP_CUDA(cudaMemcpyAsync(xb.p, x_host, (size_t)G * sizeof(double), cudaMemcpyHostToDevice, st));
  if (ds.iwp) {
    BGP_TRY(knb.alloc(ds.kneg.size() * sizeof(double)));
    BGP_TRY(kpb.alloc(ds.kpos.size() * sizeof(double)));
    if (!ds.kneg.empty())
      BGP_CUDA(cudaMemcpyAsync(knb.p, ds.kneg.data(), ds.kneg.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!ds.kpos.empty())
      BGP_CUDA(cudaMemcpyAsync(kpb.p, ds.kpos.data(), ds.kpos.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  // strip height: keep the F strip around 96 MB (L2-sized), multiple of 128 rows
  int64_t strip = (int64_t)(96.0e6 / (8.0 * (double)M));
  strip = std::max<int64_t>(128, strip / 128 * 128);
  strip = std::min<int64_t>(strip, round_up64(G, 128));
  const int64_t ldF = round_up64(M, 2);
  BGP_TRY(Db.alloc((size_t)strip * ldk * sizeof(double)));
  BGP_TRY(Fb.alloc((size_t)strip * ldF * sizeof(double)));
  BGP_TRY(ob.alloc((size_t)3 * G * sizeof(double)));
  double* o_mean = ob.as<double>();
  double* o_lo = o_mean + G;
  double* o_hi = o_lo + G;
  if (samples) BGP_TRY(Sb.alloc((size_t)G * M * sizeof(double)));
  const double alpha = 1.0 - level;
  const double q1 = alpha / 2.0, q2 = level + alpha / 2.0;
  auto rank_of = [&](double q, int64_t* r, double* h) {
    const double index = 1.0 + (double)(M - 1) * q;
    const double lo = std::floor(index);
    *r = (int64_t)lo - 1;
    *h = index - lo;
    if (*r < 0) *r = 0;
    if (*r > M - 1) *r = M - 1;
  };
  SelectArgs sa;
  rank_of(q1, &sa.r1, &sa.h1);
  rank_of(q2, &sa.r2, &sa.h2);
  sa.F = Fb.as<double>();
  sa.ldF = ldF;
  sa.M = M;
  sa.mean = o_mean;
  sa.lo = o_lo;
  sa.hi = o_hi;
  // candidate cuts of the fast path: expected tail fraction 1.5 q + 5 sqrt(q / M) (Gaussian-ish rows hold the
  // wanted ranks with a wide margin).  Per-thread slots for the expected hits plus 1.5 sigma (the shared overflow
  // list takes the rest), value buckets of about 16 / M of a tail each.
  auto norm_inv = [](double pr) {                 // Acklam's rational approximation, |error| < 1.2e-9
    static const double a_[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b_[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                 6.680131188771972e+01, -1.328068155288572e+01};
    static const double c_[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                 -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d_[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                 3.754408661907416e+00};
    if (pr < 0.02425) {
      const double q = std::sqrt(-2.0 * std::log(pr));
      return (((((c_[0] * q + c_[1]) * q + c_[2]) * q + c_[3]) * q + c_[4]) * q + c_[5]) /
             ((((d_[0] * q + d_[1]) * q + d_[2]) * q + d_[3]) * q + 1.0);
    }
    const double q = pr - 0.5, r = q * q;
    return (((((a_[0] * r + a_[1]) * r + a_[2]) * r + a_[3]) * r + a_[4]) * r + a_[5]) * q /
           (((((b_[0] * r + b_[1]) * r + b_[2]) * r + b_[3]) * r + b_[4]) * r + 1.0);
  };
  const double qt = std::max(q1, 1.0 - q2);
  const double frac = 1.5 * qt + 5.0 * std::sqrt(qt / (double)M);
  sa.pc = 0;
  sa.nb = 256;
  sa.z_lo = sa.z_hi = 0.0;
  while (sa.nb < 2048 && (int64_t)sa.nb * 16 < M) sa.nb <<= 1;
  const size_t key_bytes = (size_t)round_up64(M, 2) * sizeof(double);
  const size_t fixed_bytes = (size_t)RS_OVF * sizeof(double) + (size_t)2 * sa.nb * sizeof(unsigned int) +
                             (size_t)4 * RS_LIST * sizeof(double);
  const size_t slot_bytes = (size_t)RS_THREADS * sizeof(double);               // one candidate slot of every thread
  const size_t two_per_sm = 112 * 1024, one_per_sm = 200 * 1024;              // dynamic bytes for 2 / 1 CTAs per SM
  bool in_smem = key_bytes + fixed_bytes + RS_PC_MIN * slot_bytes <= one_per_sm;
  if (std::min(q1, 1.0 - q2) > 0.0 && frac < 0.45 && M >= 64 && !getenv("BGP_SELECT_RADIX")) {   // env: diagnostics
    const double e = (double)M / RS_THREADS * 2.0 * frac;                     // expected candidates per thread
    int want = std::max(RS_PC_MIN, (int)std::ceil(e + 1.5 * std::sqrt(e) + 2.0));
    // room for the slots: beside the row when that still leaves two CTAs per SM, else within one CTA per SM, else
    // with the row left in global memory / L2
    const size_t base = fixed_bytes + (in_smem ? key_bytes : 0);
    size_t budget = base + (size_t)RS_PC_MIN * slot_bytes <= two_per_sm ? two_per_sm : one_per_sm;
    if (base + (size_t)want * slot_bytes > budget && (size_t)want * slot_bytes > (budget - base) * 2) {
      // the slots would mostly overflow: keep the row out of shared memory instead
      in_smem = false;
      budget = fixed_bytes + (size_t)want * slot_bytes <= two_per_sm ? two_per_sm : one_per_sm;
    }
    const size_t base2 = fixed_bytes + (in_smem ? key_bytes : 0);
    const int fit = (int)((budget - base2) / slot_bytes);
    if (fit >= RS_PC_MIN && (double)fit >= e + 2.0) {
      sa.pc = std::min(want, fit);
      sa.z_lo = norm_inv(frac);
      sa.z_hi = -sa.z_lo;
    }
  }
  const size_t dyn_bytes = (in_smem ? key_bytes : 0) + (size_t)std::max(sa.pc, RS_PC_MIN) * slot_bytes + fixed_bytes;
  if (in_smem)
    BGP_CUDA(cudaFuncSetAttribute(row_select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_bytes));
  else
    BGP_CUDA(cudaFuncSetAttribute(row_select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_bytes));
  std::vector<cudaEvent_t> evs;
  auto mark = [&]() {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    evs.push_back(e);
  };
  mark();                                             // evs[0]: start
  for (int64_t g0 = 0; g0 < G; g0 += strip) {
    const int64_t rows = std::min(strip, G - g0);
    if (ds.iwp) {
      IwpDesignArgs a;
      a.x = xb.as<double>();
      a.g0 = g0;
      a.rows = rows;
      a.kneg = knb.as<double>();
      a.nneg = (int)ds.kneg.size();
      a.kpos = kpb.as<double>();
      a.npos = (int)ds.kpos.size();
      a.order = ds.order;
      a.degree = ds.degree;
      a.D = Db.as<double>();
      a.ld = ldk;
      iwp_design_kernel<<<(unsigned)rows, 128, 0, st>>>(a);
    } else {
      SgpDesignArgs a;
      a.x = xb.as<double>();
      a.g0 = g0;
      a.rows = rows;
      a.x0 = ds.x0;
      a.a = ds.a;
      a.k = ds.k;
      a.m = ds.m;
      a.boundary = ds.boundary;
      a.lo = ds.lo;
      a.hi = ds.hi;
      a.D = Db.as<double>();
      a.ld = ldk;
      sgp_design_kernel<<<(unsigned)rows, 128, 0, st>>>(a);
    }
    count_launch();
    BGP_CUDA(cudaGetLastError());
    mark();                                           // per strip: [gemm start, gemm end, select end]
    BGP_TRY(launch_kgemm(Db.as<double>(), rows, ldk, Cmat_dev, M, ldk, ds.ncols, nullptr, Fb.as<double>(), ldF, false,
                         nullptr, st));
    mark();
    if (samples)   // G x M column-major copy for only.samples = TRUE
      BGP_TRY(launch_kgemm(Db.as<double>(), rows, ldk, Cmat_dev, M, ldk, ds.ncols, nullptr, Sb.as<double>() + g0, G,
                           true, nullptr, st));
    sa.g0 = g0;
    if (in_smem) row_select_kernel<true><<<(unsigned)rows, RS_THREADS, dyn_bytes, st>>>(sa);
    else row_select_kernel<false><<<(unsigned)rows, RS_THREADS, dyn_bytes, st>>>(sa);
    count_launch();
    BGP_CUDA(cudaGetLastError());
    mark();
  }
  if (mean) BGP_CUDA(cudaMemcpyAsync(mean, o_mean, (size_t)G * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (plower) BGP_CUDA(cudaMemcpyAsync(plower, o_lo, (size_t)G * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (pupper) BGP_CUDA(cudaMemcpyAsync(pupper, o_hi, (size_t)G * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (samples)
    BGP_CUDA(cudaMemcpyAsync(samples, Sb.p, (size_t)G * M * sizeof(double), cudaMemcpyDeviceToHost, st));
  BGP_CUDA(cudaStreamSynchronize(st));
  g_pred_ms[0] = g_pred_ms[1] = g_pred_ms[2] = 0.0;
  for (size_t i = 1; i + 2 < evs.size(); i += 3) {
    float a_ms = 0, b_ms = 0;
    cudaEventElapsedTime(&a_ms, evs[i], evs[i + 1]);
    cudaEventElapsedTime(&b_ms, evs[i + 1], evs[i + 2]);
    g_pred_ms[0] += a_ms;
    g_pred_ms[1] += b_ms;
  }
  if (evs.size() >= 2) {
    float t_ms = 0;
    cudaEventElapsedTime(&t_ms, evs.front(), evs.back());
    g_pred_ms[2] = t_ms;
  }
  for (cudaEvent_t e : evs) cudaEventDestroy(e);
  return BGP_OK;
}

static void split_knots(const double* knots, int nknots, std::vector<double>& kneg, std::vector<double>& kpos) {
  double kmin = knots[0], kmax = knots[0];
  for (int i = 1; i < nknots; ++i) {
    kmin = std::min(kmin, knots[i]);
    kmax = std::max(kmax, knots[i]);
  }
  auto uniq = [](std::vector<double>& v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  };
  if (kmin >= 0) {
    kpos.assign(knots, knots + nknots);
  } else {
    for (int i = 0; i < nknots; ++i) kneg.push_back(knots[i] < 0 ? -knots[i] : 0.0);
    uniq(kneg);
    if (kmax > 0) {
      for (int i = 0; i < nknots; ++i) kpos.push_back(knots[i] > 0 ? knots[i] : 0.0);
      uniq(kpos);
    }
  }
}

// upload R-layout sample blocks and assemble the K-major coefficient matrix
static int build_coef(const double* coef, int ncoef, const double* glob, int nglob, const double* icpt, int64_t M, int skip,
                      int nX, int ldk, cudaStream_t st, DevBuf& Cb, bool inputs_on_device) {
  DevBuf cb, gb, ib;
  const double *cd = coef, *gd = glob, *id = icpt;
  if (!inputs_on_device) {
    BGP_TRY(cb.alloc((size_t)ncoef * M * sizeof(double)));
    BGP_CUDA(cudaMemcpyAsync(cb.p, coef, (size_t)ncoef * M * sizeof(double), cudaMemcpyHostToDevice, st));
    cd = cb.as<double>();
    if (glob && nglob > 0) {
      BGP_TRY(gb.alloc((size_t)nglob * M * sizeof(double)));
      BGP_CUDA(cudaMemcpyAsync(gb.p, glob, (size_t)nglob * M * sizeof(double), cudaMemcpyHostToDevice, st));
      gd = gb.as<double>();
    } else {
      gd = nullptr;
    }
    if (icpt) {
      BGP_TRY(ib.alloc((size_t)M * sizeof(double)));
      BGP_CUDA(cudaMemcpyAsync(ib.p, icpt, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, st));
      id = ib.as<double>();
    }
  }
  BGP_TRY(Cb.alloc((size_t)M * ldk * sizeof(double)));
  CoefArgs a;
  a.icpt = id;
  a.glob = gd;
  a.nglob = nglob;
  a.coef = cd;
  a.ncoef = ncoef;
  a.skip = skip;
  a.nX = nX;
  a.M = M;
  a.C = Cb.as<double>();
  a.ld = ldk;
  build_coef_kernel<<<(unsigned)M, 128, 0, st>>>(a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  BGP_CUDA(cudaStreamSynchronize(st));   // staging buffers go out of scope
  return BGP_OK;
}

}  // namespace bgp

using namespace bgp;

extern "C" {

int bgp_predict_iwp(const double* coef, const double* global, const double* icpt, int64_t M, const double* knots,
                    int nknots, int order, int degree, const double* x, int64_t G, double level, int device,
                    double* mean, double* plower, double* pupper, double* samples) {
  if (!coef || !knots || !x || M <= 0 || G <= 0 || nknots < 2 || order < 1 || order > 8 || degree < 0) {
    set_error("bgp_predict_iwp: bad arguments");
    return BGP_ERR_ARG;
  }
  if (order <= degree) {   // R/03_post_fit.R:201-203
    set_error("Error: The degree of derivative to compute is not defined. Should consider higher order smoothing "
              "model or lower order of the derivative degree.");
    return BGP_ERR_ARG;
  }
  if (!(level > 0.0 && level < 1.0)) {
    set_error("bgp_predict_iwp: level must be in (0, 1)");
    return BGP_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("no CUDA device %d available: libbgp has no CPU fallback", device);
    return BGP_ERR_CUDA;
  }
  BGP_CUDA(cudaSetDevice(device));
  DesignSpec ds;
  ds.iwp = true;
  ds.order = order;
  ds.degree = degree;
  split_knots(knots, nknots, ds.kneg, ds.kpos);
  const int nB = (ds.kneg.empty() ? 0 : (int)ds.kneg.size() - 1) + (ds.kpos.empty() ? 0 : (int)ds.kpos.size() - 1);
  const int nX = order - degree;
  ds.ncols = nX + nB;
  const int ldk = round_up(ds.ncols, 16);
  cudaStream_t st;
  BGP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  DevBuf Cb;
  int rc = build_coef(coef, nB, global, order - 1, icpt, M, degree, nX, ldk, st, Cb, false);
  if (rc == BGP_OK) rc = predict_core(ds, Cb.as<double>(), ldk, M, x, G, level, st, mean, plower, pupper, samples);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return rc;
}

int bgp_predict_sgp(const double* coef, const double* global, const double* icpt, int64_t M, double a, int k, int m,
                    const double* region, int boundary, const double* x, int64_t G, double level, int device,
                    double* mean, double* plower, double* pupper, double* samples) {
  if (!coef || !region || !x || M <= 0 || G <= 0 || k < 5 || m < 1) {
    set_error("bgp_predict_sgp: bad arguments (k must be >= 5)");
    return BGP_ERR_ARG;
  }
  if (!(level > 0.0 && level < 1.0)) {
    set_error("bgp_predict_sgp: level must be in (0, 1)");
    return BGP_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("no CUDA device %d available: libbgp has no CPU fallback", device);
    return BGP_ERR_CUDA;
  }
  BGP_CUDA(cudaSetDevice(device));
  DesignSpec ds;
  ds.iwp = false;
  ds.a = a;
  ds.k = k;
  ds.m = m;
  ds.boundary = boundary ? 1 : 0;
  ds.lo = std::min(region[0], region[1]);
  ds.hi = std::max(region[0], region[1]);
  ds.x0 = x[0];
  for (int64_t i = 1; i < G; ++i) ds.x0 = std::min(ds.x0, x[i]);   // initial_location = NULL => min(refined_x)
  const int nb = boundary ? k - 2 : k;
  const int nX = 1 + 2 * m;
  ds.ncols = nX + 3 * nb * m;
  const int ldk = round_up(ds.ncols, 16);
  cudaStream_t st;
  BGP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  DevBuf Cb;
  int rc = build_coef(coef, 3 * nb * m, global, 2 * m, icpt, M, 0, nX, ldk, st, Cb, false);
  if (rc == BGP_OK) rc = predict_core(ds, Cb.as<double>(), ldk, M, x, G, level, st, mean, plower, pupper, samples);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return rc;
}

int bgp_predict_last_timing(double* gemm_ms, double* select_ms, double* total_ms) {
  if (gemm_ms) *gemm_ms = g_pred_ms[0];
  if (select_ms) *select_ms = g_pred_ms[1];
  if (total_ms) *total_ms = g_pred_ms[2];
  return BGP_OK;
}

int bgp_basis_iwp(const double* knots, int nknots, int order, const double* x, int64_t G, int device, double* out) {
  if (!knots || !x || !out || nknots < 2 || order < 1 || order > 8 || G <= 0) {
    set_error("bgp_basis_iwp: bad arguments");
    return BGP_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("no CUDA device %d available: libbgp has no CPU fallback", device);
    return BGP_ERR_CUDA;
  }
  BGP_CUDA(cudaSetDevice(device));
  std::vector<double> kneg, kpos;
  split_knots(knots, nknots, kneg, kpos);
  const int nB = (kneg.empty() ? 0 : (int)kneg.size() - 1) + (kpos.empty() ? 0 : (int)kpos.size() - 1);
  DevBuf xb, knb, kpb, Bb;
  BGP_TRY(xb.alloc((size_t)G * sizeof(double)));
  BGP_TRY(knb.alloc(kneg.size() * sizeof(double)));
  BGP_TRY(kpb.alloc(kpos.size() * sizeof(double)));
  BGP_TRY(Bb.alloc((size_t)G * nB * sizeof(double)));
  BGP_CUDA(cudaMemcpy(xb.p, x, (size_t)G * sizeof(double), cudaMemcpyHostToDevice));
  if (!kneg.empty()) BGP_CUDA(cudaMemcpy(knb.p, kneg.data(), kneg.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (!kpos.empty()) BGP_CUDA(cudaMemcpy(kpb.p, kpos.data(), kpos.size() * sizeof(double), cudaMemcpyHostToDevice));
  BGP_TRY(launch_iwp_block(nullptr, xb.as<double>(), G, 0.0, kneg.empty() ? nullptr : knb.as<double>(), (int)kneg.size(),
                           kpos.empty() ? nullptr : kpb.as<double>(), (int)kpos.size(), order, Bb.as<double>(), (int)G,
                           nullptr, 0, true, 0));
  BGP_CUDA(cudaDeviceSynchronize());
  BGP_CUDA(cudaMemcpy(out, Bb.p, (size_t)G * nB * sizeof(double), cudaMemcpyDeviceToHost));
  return BGP_OK;
}

}  // extern "C"
