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
#include "ptx.cuh"

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
  const double* icpt;     // M (stride icpt_ld) or NULL
  const double* glob;     // nglob x M column-major (column pitch glob_ld) or NULL
  int nglob;
  const double* coef;     // ncoef x M column-major (column pitch coef_ld)
  int ncoef;
  int64_t icpt_ld, glob_ld, coef_ld;
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
      if (r == 0) v = a.icpt ? a.icpt[(size_t)s * a.icpt_ld] : 0.0;
      else v = (a.glob && r - 1 < a.nglob) ? a.glob[(size_t)s * a.glob_ld + (r - 1)] : 0.0;
    } else if (c < a.nX + a.ncoef) {
      v = a.coef[(size_t)s * a.coef_ld + (c - a.nX)];
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

// R: qs <- x[lo]; where (index > lo & x[hi] != qs): qs <- (1 - h) * qs + h * x[hi]   (stats::quantile type 7).
// Spelled with explicit roundings: a contracted FMA would differ from R by one ulp, and differently per call site.
__device__ __forceinline__ double type7_interp(double h, double x_lo, double x_hi) {
  return (h != 0.0 && x_hi != x_lo) ? __dadd_rn(__dmul_rn(1.0 - h, x_lo), __dmul_rn(h, x_hi)) : x_lo;
}

constexpr int RS_LIST = 128;      // keys of one target bucket, ranked by a single warp
constexpr int RS_OVF = 256;       // shared overflow list behind the per-thread candidate slots
constexpr int RS_BATCH = 10;      // 128-bit loads in flight per thread
constexpr int RS_RH = 8;          // radix path: histogram copies (warps w and w + 8 share one), 16 KB
constexpr int RS_SLOT_AREA_MIN = 2 * RS_RH * 256 * 4;   // the slot area doubles as those histograms
constexpr int RS_DBG_POINTS = 8;
constexpr int RS_PCREG = 16;      // 256-thread CTAs: bucket codes of up to this many private candidates stay in registers

struct SelectArgs {
  const double* F;
  int64_t ldF, M;
  int64_t g0;
  int64_t r1, r2;     // 0-based ranks floor(index) - 1 of the two probabilities
  double h1, h2;      // interpolation weights (0 => no second order statistic needed)
  double z_lo, z_hi;  // candidate cuts in standard deviations from the row mean (fast path)
  int pc;             // candidate slots per thread (0 => radix path only)
  int nb;             // value buckets per tail (power of two, multiple of 256)
  double* mean;
  double* lo;
  double* hi;
  long long* dbg;     // BGP_SEL_DEBUG: clock64 at RS_DBG_POINTS points of every row of the first strip
  int64_t rows;       // rows of this strip
  unsigned int* counter;   // next row to hand out (zeroed before the launch)
};

// Persistent CTAs (two per SM), rows handed out by an atomic counter.  Per row, four steps, each a few
// instructions per value:
//   A  the row arrives in shared memory by one bulk copy (cp.async.bulk + mbarrier) that was started while the
//      previous row was still being ranked; one pass over it gives the sum and the sum of squares;
//   B  every thread goes over its own values again and keeps those further than z sigma from the mean (about 1.8 x
//      the wanted tail) in private slots — one FP64 subtraction and an integer compare per value, no ballots or
//      atomics; the rare thread with more hits than slots makes a second pass and spills to a shared list;
//   C  the kept values are counted into nb buckets per tail by distance from the cut (a monotone map, so bucket
//      order is value order); a scan from the extreme end finds the bucket holding each wanted order statistic and
//      the exact number of values beyond it;
//   D  the handful of values of that bucket are collected and ranked by one warp.
// Every count is exact, so the result is the exact order statistic.  Anything unexpected — cuts that miss the rank
// (heavy tails, constant rows), NaN / Inf, a crowded bucket, a full overflow list — drops the whole row to the MSB
// radix select at the end of the loop body, which needs no assumption about the values.
template <bool IN_SMEM, int RS_THREADS>
__global__ void __launch_bounds__(RS_THREADS, 2) row_select_kernel(const SelectArgs a) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  constexpr int NW = RS_THREADS / 32, RS_HALF = RS_THREADS / 2;
  constexpr int RS_PC_MIN = RS_SLOT_AREA_MIN / (RS_THREADS * 8);
  constexpr int PCREG = RS_THREADS == 256 ? RS_PCREG : 1;                      // candidates cached in registers
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t M = a.M;
  const int64_t Mpad = (M + 1) & ~(int64_t)1;
  const int pc = a.pc, nb = a.nb;
  double* vals = reinterpret_cast<double*>(rs_smem);
  double* priv = vals + (IN_SMEM ? Mpad : 0);                                  // [slot][thread]
  double* ovf = priv + (size_t)(pc > RS_PC_MIN ? pc : RS_PC_MIN) * RS_THREADS;
  unsigned int* hist = reinterpret_cast<unsigned int*>(ovf + RS_OVF);          // [2][nb]
  double* lists = reinterpret_cast<double*>(hist);     // [2 tails][2 ranks][RS_LIST]: used once the scan is done with hist
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ long long s_rowix[2];
  __shared__ double s_red[2 * NW];
  __shared__ unsigned long long s_kmin[NW], s_kmax[NW];
  __shared__ unsigned int s_novf, s_wtot[NW], s_n[2], s_lcnt[2][2];
  __shared__ int s_bsel[2][2];
  __shared__ unsigned int s_before[2][2];
  __shared__ double s_val[2][2];
  __shared__ unsigned long long s_prefix[2];
  __shared__ long long s_below[2];
  __shared__ unsigned int s_eq[2];
  const int64_t npair = M >> 1;
  const uint32_t bar = ptx::smem_u32(&s_bar), vals_addr = ptx::smem_u32(vals);
  const uint32_t row_bytes = (uint32_t)(Mpad * sizeof(double));

  // thread 0: claim the next row and, when rows are staged in shared memory, start its copy
  auto fetch = [&](int slot, unsigned claimed) {
    const long long r = (long long)claimed;
    s_rowix[slot] = r;
    if (IN_SMEM && r < a.rows) {
      ptx::mbar_expect_tx(bar, row_bytes);
      ptx::bulk_load_1d(vals_addr, a.F + (size_t)r * a.ldF, row_bytes, bar);   // 16-byte aligned: ldF is even
    }
  };
  if (tid == 0) {
    if (IN_SMEM) {
      ptx::mbar_init(bar, 1);
      ptx::mbar_fence_init();
    }
    fetch(0, atomicAdd(a.counter, 1u));
  }
  __syncthreads();
  uint32_t parity = 0;
  for (int slot = 0;; slot ^= 1) {
    const long long rix = s_rowix[slot];
    if (rix >= a.rows) break;
    const double* row = a.F + (size_t)rix * a.ldF;
    const double2* row2 = reinterpret_cast<const double2*>(row);
    bool fetched = false;                          // next row's copy started: `vals` no longer holds this row
    // the next row is claimed now and used at the hand-over: the atomic's round trip hides behind steps A and B
    unsigned claimed = 0;
    if (tid == 0) claimed = atomicAdd(a.counter, 1u);
    auto stamp = [&](int k) {
      if (a.dbg && tid == 0) a.dbg[(size_t)rix * RS_DBG_POINTS + k] = clock64();
    };
    stamp(0);

    // ---- A: mean (fixed-order tree) and spread -------------------------------------------------------------
    // (squares are taken about the first value of the row: the variance only places the cuts, but a row whose
    // spread is tiny against its level must not lose it to cancellation)
    double sum = 0.0, sumsq = 0.0;
    double shift;
    if (IN_SMEM) {
      ptx::mbar_wait(bar, parity);
      parity ^= 1u;
      shift = vals[0];
      const double2* vals2 = reinterpret_cast<const double2*>(vals);
      double sum1 = 0.0, sumsq1 = 0.0;                   // second chain: the FP64 adds are latency bound otherwise
      auto acc = [&](const double2 v) {
        const double dx = v.x - shift, dy = v.y - shift;
        sum += v.x;
        sumsq = fma(dx, dx, sumsq);
        sum1 += v.y;
        sumsq1 = fma(dy, dy, sumsq1);
      };
      int64_t p = tid;
      for (; p + 3 * RS_THREADS < npair; p += 4 * RS_THREADS) {
        const double2 v0 = vals2[p], v1 = vals2[p + RS_THREADS], v2 = vals2[p + 2 * RS_THREADS],
                      v3 = vals2[p + 3 * RS_THREADS];
        acc(v0);
        acc(v1);
        acc(v2);
        acc(v3);
      }
      for (; p < npair; p += RS_THREADS) acc(vals2[p]);
      sum += sum1;
      sumsq += sumsq1;
      if ((M & 1) && tid == 0) {
        const double v = vals[M - 1];
        sum += v;
        sumsq = fma(v - shift, v - shift, sumsq);
      }
    } else {
      shift = row[0];
      for (int64_t p0 = tid; p0 < npair; p0 += RS_BATCH * RS_THREADS) {
        double2 v[RS_BATCH];
#pragma unroll
        for (int u = 0; u < RS_BATCH; ++u) {
          const int64_t p = p0 + (int64_t)u * RS_THREADS;
          v[u] = p < npair ? row2[p] : make_double2(shift, shift);
        }
#pragma unroll
        for (int u = 0; u < RS_BATCH; ++u) {
          const int64_t p = p0 + (int64_t)u * RS_THREADS;
          if (p < npair) {
            const double dx = v[u].x - shift, dy = v[u].y - shift;
            sum += v[u].x;
            sumsq = fma(dx, dx, sumsq);
            sum += v[u].y;
            sumsq = fma(dy, dy, sumsq);
          }
        }
      }
      if ((M & 1) && tid == 0) {
        const double v = row[M - 1];
        sum += v;
        sumsq = fma(v - shift, v - shift, sumsq);
      }
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
    stamp(1);

    const double sd = sqrt(fmax(sumsq / (double)M - (mean - shift) * (mean - shift), 0.0));
    if (pc > 0 && sd > 0.0 && isfinite(sd) && isfinite(mean)) {                // CTA-uniform
      // a value is a candidate when the high word of |v - mean| exceeds that of z sd: v -> fl(v - mean) is monotone,
      // so each tail's candidates are exactly the values beyond some threshold, whatever the rounding
      const double rad = a.z_hi * sd;
      const int rad_hi = __double2hiint(rad);
      const double scale = (double)nb / (3.0 * sd);                            // buckets span three sigma beyond the cut
      bool ok = true;
      // ---- B: candidates into private slots ---------------------------------------------------------------
      auto is_hit = [&](double v) { return (__double2hiint(v - mean) & 0x7fffffff) > rad_hi; };
      auto for_own = [&](auto&& f) {
        if (IN_SMEM) {
          const double2* vals2 = reinterpret_cast<const double2*>(vals);
          int64_t p = tid;
          for (; p + 3 * RS_THREADS < npair; p += 4 * RS_THREADS) {            // loads first: four in flight
            const double2 v0 = vals2[p], v1 = vals2[p + RS_THREADS], v2 = vals2[p + 2 * RS_THREADS],
                          v3 = vals2[p + 3 * RS_THREADS];
            f(v0.x);
            f(v0.y);
            f(v1.x);
            f(v1.y);
            f(v2.x);
            f(v2.y);
            f(v3.x);
            f(v3.y);
          }
          for (; p < npair; p += RS_THREADS) {
            const double2 v = vals2[p];
            f(v.x);
            f(v.y);
          }
          if ((M & 1) && tid == 0) f(vals[M - 1]);
        } else {
          for (int64_t p0 = tid; p0 < npair; p0 += 4 * RS_THREADS) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int64_t p = p0 + (int64_t)u * RS_THREADS;
              v[u] = p < npair ? row2[p] : make_double2(mean, mean);           // the mean is never a candidate
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              f(v[u].x);
              f(v[u].y);
            }
          }
          if ((M & 1) && tid == 0) f(row[M - 1]);
        }
      };
      // (a predicated store, spelled out: the compiler otherwise branches around every candidate)
      const uint32_t priv_addr = ptx::smem_u32(priv) + (uint32_t)tid * 8u;
      int c = 0;
      for_own([&](double v) {
        const int hit = is_hit(v) ? 1 : 0;
        const int keep = (hit != 0 && c < pc) ? 1 : 0;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %2, 0;\n"
            "@p st.shared.f64 [%0], %1;\n"
            "}\n" ::"r"(priv_addr + (uint32_t)c * (uint32_t)(RS_THREADS * 8)),
            "d"(v), "r"(keep)
            : "memory");
        c += hit;
      });
      if (c > pc) {
        int k = 0;
        for_own([&](double v) {
          k += is_hit(v) ? 1 : 0;
          if (k > pc && is_hit(v)) {                       // rare: only the excess takes the branch
            const unsigned pos = atomicAdd(&s_novf, 1u);
            if (pos < (unsigned)RS_OVF) ovf[pos] = v;
          }
        });
      }
      stamp(2);
      // ---- C: bucket counts, extreme end first -------------------------------------------------------------
      // tail 0 = below the mean, tail 1 = above; bucket nb - 1 is the far end.  |v1 - mean| <= |v2 - mean| implies
      // bucket(v1) <= bucket(v2) within a tail: subtraction, scaling and floor are all monotone, so the buckets
      // partition the candidates of a tail in value order.  code = tail * nb + bucket.
      auto code_of = [&](double v) {
        const double d = v - mean;
        const int q = __double2hiint(d) < 0 ? 0 : 1;
        const int b = __double2int_rd((fabs(d) - rad) * scale);                // saturating; |d| > rad for a candidate
        return q * nb + (b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b));
      };
      const int cp = c < pc ? c : pc;
      const bool in_regs = pc <= PCREG;                                        // CTA-uniform
      int code[PCREG];
      if (in_regs) {
#pragma unroll
        for (int j = 0; j < PCREG; ++j) {
          const bool act = j < cp;
          const double v = priv[(act ? j : 0) * RS_THREADS + tid];             // slot 0 of an idle lane: stale, unused
          const int cd = code_of(v);
          code[j] = act ? cd : -1;
          if (act) atomicAdd(&hist[cd], 1u);
        }
      } else {
        for (int j = 0; j < cp; ++j) atomicAdd(&hist[code_of(priv[j * RS_THREADS + tid])], 1u);
      }
      __syncthreads();
      // every thread is past step B: the row buffer is free for the next row
      if (tid == 0) fetch(slot ^ 1, claimed);
      fetched = true;
      const unsigned novf = s_novf;
      if (novf > (unsigned)RS_OVF) ok = false;
      const unsigned novf_c = novf > (unsigned)RS_OVF ? 0u : novf;
      for (unsigned i = tid; i < novf_c; i += RS_THREADS) atomicAdd(&hist[code_of(ovf[i])], 1u);
      __syncthreads();
      stamp(3);
      // half of the threads per tail, nb / RS_HALF consecutive buckets each, walking inwards from the far end
      const int q = tid / RS_HALF, u = tid % RS_HALF, w = nb / RS_HALF;
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
      for (int ww = q * (NW / 2); ww < warp; ++ww) base += s_wtot[ww];
      const long long cum0 = (long long)base + (long long)(inc - tot);       // candidates beyond this thread's buckets
      if (u == RS_HALF - 1) s_n[q] = base + inc;
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
      stamp(4);
      // the cuts must not have missed a wanted rank
      if ((need_lo ? a.r1 + 1 : a.r1) >= (long long)s_n[0]) ok = false;
      if (top2 >= (long long)s_n[1] || (need_hi && top2 < 1)) ok = false;
      // ---- D: collect the target buckets, rank inside them -----------------------------------------------------
      const int b00 = s_bsel[0][0], b01 = s_bsel[0][1], b10 = s_bsel[1][0], b11 = s_bsel[1][1];
      if (ok) {
        // bucket codes of the four wanted order statistics (-1: not wanted, -2: same bucket as the first of the tail)
        const int c00 = b00, c01 = b01 < 0 ? -1 : (b01 == b00 ? -2 : b01);
        const int c10 = nb + b10, c11 = b11 < 0 ? -1 : (b11 == b10 ? -2 : nb + b11);
        auto collect = [&](int cd, double v) {
          int l = -1;
          if (cd == c00) l = 0;
          else if (cd == c01) l = 1;
          else if (cd == c10) l = 2;
          else if (cd == c11) l = 3;
          if (l >= 0) {
            const unsigned pos = atomicAdd(&s_lcnt[l >> 1][l & 1], 1u);
            if (pos < (unsigned)RS_LIST) lists[l * RS_LIST + pos] = v;
          }
        };
        if (in_regs) {
#pragma unroll
          for (int j = 0; j < PCREG; ++j)
            if (code[j] >= 0 && (code[j] == c00 || code[j] == c01 || code[j] == c10 || code[j] == c11))
              collect(code[j], priv[j * RS_THREADS + tid]);
        } else {
          for (int j = 0; j < cp; ++j) {
            const double v = priv[j * RS_THREADS + tid];
            collect(code_of(v), v);
          }
        }
        for (unsigned i = tid; i < novf_c; i += RS_THREADS) collect(code_of(ovf[i]), ovf[i]);
      }
      __syncthreads();
      stamp(5);
      if (ok) {
        if (s_lcnt[0][0] > (unsigned)RS_LIST || s_lcnt[0][1] > (unsigned)RS_LIST || s_lcnt[1][0] > (unsigned)RS_LIST ||
            s_lcnt[1][1] > (unsigned)RS_LIST)
          ok = false;                                                          // crowded bucket
      }
      if (ok && warp < 4) {
        const int qq = warp >> 1, wch = warp & 1;
        if (wch == 0 || (qq == 0 ? need_lo : need_hi)) {
          const int bsel0 = qq ? b10 : b00, bsel1 = qq ? b11 : b01;
          const int src = (wch == 1 && bsel1 == bsel0) ? 0 : wch;
          const int n = (int)s_lcnt[qq][src];
          const double* L = lists + (qq * 2 + src) * RS_LIST;
          const long long P = qq == 0 ? (wch == 0 ? a.r1 : a.r1 + 1) : (wch == 0 ? top2 : top2 - 1);
          const long long t = P - (long long)s_before[qq][wch];                // position inside the bucket
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
      stamp(6);
      if (ok) {
        q_lo = type7_interp(a.h1, s_val[0][0], need_lo ? s_val[0][1] : s_val[0][0]);
        q_hi = type7_interp(a.h2, s_val[1][0], need_hi ? s_val[1][1] : s_val[1][0]);
        done = true;
      }
    }
    if (!done) {
      // ---- fallback: exact MSB radix select (8-bit digits) of both order statistics in the same sweeps -------------
      // per-warp histograms fed by warp-aggregated increments (the keys of a row share their leading bytes, a
      // single shared histogram serialises on one bin), a warp-parallel bin scan, and a start digit chosen from the
      // highest bit in which the row's keys differ.  The histograms live in the candidate-slot area; the keys come
      // from shared memory while it still holds this row, else from global memory / L2.
      __syncthreads();
      const bool from_smem = IN_SMEM && !fetched;
      const int64_t rk[2] = {a.r1, a.r2};
      double qv[2] = {0.0, 0.0};
      unsigned int (*whist)[RS_RH][256] = reinterpret_cast<unsigned int (*)[RS_RH][256]>(priv);
      unsigned int (*rhist)[256] = reinterpret_cast<unsigned int (*)[256]>(hist);  // [2][256] (nb >= 256)
      auto key_at = [&](int64_t i) -> unsigned long long { return dkey(from_smem ? vals[i] : row[i]); };
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
        const int shift_bits = pass * 8;
        const bool same = prefix[0] == prefix[1];        // both quantiles still in the same bucket: one histogram
        for (int t = tid; t < 2 * RS_RH * 256; t += RS_THREADS) (&whist[0][0][0])[t] = 0;
        __syncthreads();
        for (int64_t i0 = 0; i0 < M; i0 += RS_THREADS) {
          const int64_t i = i0 + tid;
          const bool in = i < M;
          const unsigned long long k = in ? key_at(i) : 0ull;
          const unsigned digit = (unsigned)(k >> shift_bits) & 255u;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q == 1 && same) break;
            const bool hit = in && (k & mask) == prefix[q];
            const unsigned act = __ballot_sync(0xffffffffu, hit);
            if (hit) {
              const unsigned peers = __match_any_sync(act, digit);
              if (lane == __ffs(peers) - 1) atomicAdd(&whist[q][warp % RS_RH][digit], (unsigned)__popc(peers));
            }
          }
        }
        __syncthreads();
        for (int t = tid; t < 2 * 256; t += RS_THREADS) {
          const int q = t >> 8, bin = t & 255;
          unsigned s = 0;
#pragma unroll
          for (int w = 0; w < RS_RH; ++w) s += whist[same ? 0 : q][w][bin];
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
            s_prefix[q] = prefix[q] | ((unsigned long long)(lane * 8 + j) << shift_bits);
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
        mask |= 0xFFull << shift_bits;
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
        qv[which] = type7_interp(h, x_lo, x_hi);
      }
      q_lo = qv[0];
      q_hi = qv[1];
    }
    if (tid == 0) {
      const int64_t g = a.g0 + rix;
      if (a.mean) a.mean[g] = mean;
      if (a.lo) a.lo[g] = q_lo;
      if (a.hi) a.hi[g] = q_hi;
    }
    if (done) stamp(7);
    // hand-over to the next row: every read of this row's shared state is behind this barrier
    __syncthreads();
    if (!fetched && tid == 0) fetch(slot ^ 1, claimed);
    __syncthreads();
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
static thread_local double g_pred_slices[2] = {0.0, 0.0};     // executed / total {128-row x 16-column} design slices

// cnt[0] += popcount of the tile words, cnt[1] += tiles * slices
__global__ void occ_count_kernel(const unsigned long long* __restrict__ occ, int ntiles, int nslices, double* __restrict__ cnt) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int t = 0; t < ntiles; ++t) a += (double)__popcll(occ[t]);
    atomicAdd(&cnt[0], a);                      // two strips may be in flight on two streams
    atomicAdd(&cnt[1], (double)ntiles * nslices);
  }
}

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
  DevBuf xb, knb, kpb, Db, Fb, Db2, Fb2, ob, Sb;
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
  // Two strips are in flight on two streams (the GEMM of one overlaps the quantile selection of the other: the
  // selection is latency-bound, the GEMM tensor-pipe bound); strip height: both F strips together around 96 MB
  // (L2-sized), multiple of 128 rows.  BGP_PREDICT_SERIAL=1: one strip at a time on one stream (per-kernel timing).
  const bool overlap = getenv("BGP_PREDICT_SERIAL") == nullptr && getenv("BGP_SEL_DEBUG") == nullptr;   // env: diagnostics
  int64_t strip = (int64_t)((overlap ? 48.0e6 : 96.0e6) / (8.0 * (double)M));
  strip = std::max<int64_t>(128, strip / 128 * 128);
  strip = std::min<int64_t>(strip, round_up64(G, 128));
  const int64_t ldF = round_up64(M, 2);
  BGP_TRY(Db.alloc((size_t)strip * ldk * sizeof(double)));
  BGP_TRY(Fb.alloc((size_t)strip * ldF * sizeof(double)));
  if (overlap) {
    BGP_TRY(Db2.alloc((size_t)strip * ldk * sizeof(double)));
    BGP_TRY(Fb2.alloc((size_t)strip * ldF * sizeof(double)));
  }
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
  sa.dbg = nullptr;
  while (sa.nb < 2048 && (int64_t)sa.nb * 32 < M) sa.nb <<= 1;
  // CTA width: per-row fixed work (scans, barriers) grows with the thread count, the latency hidden too
  int threads = M > 16384 ? 512 : 256;
  if (const char* e = getenv("BGP_SEL_THREADS")) threads = atoi(e) == 512 ? 512 : 256;       // env: diagnostics
  const size_t key_bytes = (size_t)round_up64(M, 2) * sizeof(double);
  const size_t fixed_bytes = (size_t)RS_OVF * sizeof(double) +          // overflow list, then bucket counts / rank lists
                             std::max((size_t)2 * sa.nb * sizeof(unsigned int), (size_t)4 * RS_LIST * sizeof(double));
  const size_t slot_bytes = (size_t)threads * sizeof(double);                  // one candidate slot of every thread
  const int pc_min = (int)(RS_SLOT_AREA_MIN / slot_bytes);
  const size_t two_per_sm = 112 * 1024, one_per_sm = 200 * 1024;              // dynamic bytes for 2 / 1 CTAs per SM
  bool in_smem = key_bytes + fixed_bytes + RS_SLOT_AREA_MIN <= one_per_sm;
  if (std::min(q1, 1.0 - q2) > 0.0 && frac < 0.45 && M >= 64 && !getenv("BGP_SELECT_RADIX")) {   // env: diagnostics
    const double e = (double)M / threads * 2.0 * frac;                        // expected candidates per thread
    const int want = std::max(pc_min, (int)std::ceil(e + 3.5 * std::sqrt(e) + 1.0));
    const int least = std::max(pc_min, (int)std::ceil(e + 2.0));
    // slots that fit beside (row in shared memory?) under a budget
    auto fit = [&](bool row_in, size_t budget) {
      const size_t base = fixed_bytes + (row_in ? key_bytes : 0);
      return base >= budget ? 0 : (int)((budget - base) / slot_bytes);
    };
    // preference: row in shared memory at two CTAs per SM, row re-read from L2 at two CTAs per SM, then one CTA per SM
    const bool row_opts[4] = {true, false, true, false};
    const size_t budgets[4] = {two_per_sm, two_per_sm, one_per_sm, one_per_sm};
    for (int o = 0; o < 4; ++o) {
      if (row_opts[o] && !in_smem) continue;
      const int f = fit(row_opts[o], budgets[o]);
      if (f >= least) {
        in_smem = row_opts[o];
        sa.pc = std::min(want, f);
        sa.z_lo = norm_inv(frac);
        sa.z_hi = -sa.z_lo;
        break;
      }
    }
  }
  const size_t dyn_bytes = (in_smem ? key_bytes : 0) + std::max((size_t)sa.pc * slot_bytes, (size_t)RS_SLOT_AREA_MIN) +
                           fixed_bytes;
  DevBuf cntb;
  BGP_TRY(cntb.alloc(2 * sizeof(unsigned int)));
  sa.counter = cntb.as<unsigned int>();
  int n_sm = 0, cur_dev = 0;
  BGP_CUDA(cudaGetDevice(&cur_dev));
  BGP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cur_dev));
  // persistent CTAs: as many as are resident at once, rows handed out by the counter
  auto launch_select = [&](unsigned rows, cudaStream_t st) -> int {
    sa.rows = rows;
    BGP_CUDA(cudaMemsetAsync(sa.counter, 0, sizeof(unsigned int), st));
#define BGP_RS_LAUNCH(IS, T)                                                                                        \
  do {                                                                                                              \
    BGP_CUDA(cudaFuncSetAttribute(row_select_kernel<IS, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                  (int)dyn_bytes));                                                                 \
    int per_sm = 0;                                                                                                 \
    BGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, row_select_kernel<IS, T>, T, dyn_bytes));       \
    const unsigned grid = std::min<unsigned>(rows, (unsigned)std::max(1, per_sm) * (unsigned)n_sm);                 \
    row_select_kernel<IS, T><<<grid, T, dyn_bytes, st>>>(sa);                                                       \
  } while (0)
    if (in_smem) {
      if (threads == 512) BGP_RS_LAUNCH(true, 512);
      else BGP_RS_LAUNCH(true, 256);
    } else {
      if (threads == 512) BGP_RS_LAUNCH(false, 512);
      else BGP_RS_LAUNCH(false, 256);
    }
#undef BGP_RS_LAUNCH
    count_launch();
    BGP_CUDA(cudaGetLastError());
    return BGP_OK;
  };
  DevBuf dbgb;
  const bool sel_debug = getenv("BGP_SEL_DEBUG") != nullptr;
  std::vector<cudaEvent_t> evs;
  auto mark_on = [&](cudaStream_t sx) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, sx);
    evs.push_back(e);
  };
  auto mark = [&]() { mark_on(st); };
  // second stream: waits for the uploads on the first, joins it again after the last strip
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_up = nullptr, ev_join = nullptr;
  if (overlap) {
    BGP_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
    BGP_CUDA(cudaEventCreateWithFlags(&ev_up, cudaEventDisableTiming));
    BGP_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  struct StreamGuard {
    cudaStream_t& s;
    cudaEvent_t &a, &b;
    ~StreamGuard() {
      if (s) {
        cudaStreamSynchronize(s);
        cudaStreamDestroy(s);
      }
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } stream_guard{st2, ev_up, ev_join};
  static const bool occ_env = getenv("BGP_PREDICT_DENSE") == nullptr;     // env: diagnostics (multiply the zeros too)
  const bool use_occ = occ_env && ds.ncols <= 1024;
  DevBuf occb, occ_cnt;
  const size_t occ_words = (size_t)((strip + 127) / 128);
  if (use_occ) {
    BGP_TRY(occb.alloc(2 * occ_words * sizeof(unsigned long long)));
    BGP_TRY(occ_cnt.alloc(2 * sizeof(double)));
    BGP_CUDA(cudaMemsetAsync(occ_cnt.p, 0, 2 * sizeof(double), st));
  }
  g_pred_slices[0] = g_pred_slices[1] = 0.0;
  mark();                                             // evs[0]: start
  if (overlap) {
    BGP_CUDA(cudaEventRecord(ev_up, st));
    BGP_CUDA(cudaStreamWaitEvent(st2, ev_up, 0));
  }
  cudaStream_t st_main = st;
  int strip_no = 0;
  for (int64_t g0 = 0; g0 < G; g0 += strip, ++strip_no) {
    const int64_t rows = std::min(strip, G - g0);
    const int set = overlap ? (strip_no & 1) : 0;
    cudaStream_t st = set ? st2 : st_main;            // everything of this strip goes to its own stream
    DevBuf& Dset = set ? Db2 : Db;
    DevBuf& Fset = set ? Fb2 : Fb;
    sa.F = Fset.as<double>();
    sa.counter = cntb.as<unsigned int>() + set;
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
      a.D = Dset.as<double>();
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
      a.D = Dset.as<double>();
      a.ld = ldk;
      sgp_design_kernel<<<(unsigned)rows, 128, 0, st>>>(a);
    }
    count_launch();
    BGP_CUDA(cudaGetLastError());
    // which 16-column slices of each 128-row tile of the design carry anything: x_new is sorted, so an O-spline
    // block is a staircase and a cubic-B-spline block a band — about half of the slices are structurally empty
    const unsigned long long* occ = nullptr;
    if (use_occ) {
      unsigned long long* occ_set = occb.as<unsigned long long>() + (size_t)set * occ_words;
      BGP_TRY(launch_kgemm_occ(Dset.as<double>(), rows, ldk, ds.ncols, occ_set, st));
      occ = occ_set;
      occ_count_kernel<<<1, 32, 0, st>>>(occ, (int)((rows + 127) / 128), (ds.ncols + 15) / 16, occ_cnt.as<double>());
      count_launch();
    }
    if (!overlap) mark();                             // per strip: [gemm start, gemm end, select end]
    BGP_TRY(launch_kgemm(Dset.as<double>(), rows, ldk, Cmat_dev, M, ldk, ds.ncols, nullptr, Fset.as<double>(), ldF, false,
                         nullptr, st, occ));
    if (!overlap) mark();
    if (samples)   // G x M column-major copy for only.samples = TRUE
      BGP_TRY(launch_kgemm(Dset.as<double>(), rows, ldk, Cmat_dev, M, ldk, ds.ncols, nullptr, Sb.as<double>() + g0, G,
                           true, nullptr, st, occ));
    sa.g0 = g0;
    sa.dbg = nullptr;
    if (sel_debug && g0 == 0) {
      BGP_TRY(dbgb.alloc((size_t)rows * RS_DBG_POINTS * sizeof(long long)));
      BGP_CUDA(cudaMemsetAsync(dbgb.p, 0, (size_t)rows * RS_DBG_POINTS * sizeof(long long), st));
      sa.dbg = dbgb.as<long long>();
    }
    BGP_TRY(launch_select((unsigned)rows, st));
    if (sa.dbg) {
      std::vector<long long> h((size_t)rows * RS_DBG_POINTS);
      BGP_CUDA(cudaMemcpyAsync(h.data(), dbgb.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
      BGP_CUDA(cudaStreamSynchronize(st));
      double acc[RS_DBG_POINTS] = {0};
      int64_t nfast = 0;
      for (int64_t r = 0; r < rows; ++r) {
        const long long* t = h.data() + (size_t)r * RS_DBG_POINTS;
        if (!t[7]) continue;                                     // row went to the radix path
        ++nfast;
        for (int k = 1; k < RS_DBG_POINTS; ++k) acc[k] += (double)(t[k] - t[k - 1]);
      }
      fprintf(stderr, "[select] M %lld threads %d in_smem %d pc %d nb %d dyn %zu B; %lld of %lld rows on the fast path; "
              "mean clocks: load+stats %.0f, cuts %.0f, candidates %.0f, buckets %.0f, scan %.0f, collect %.0f, rank %.0f\n",
              (long long)M, threads, (int)in_smem, sa.pc, sa.nb, dyn_bytes, (long long)nfast, (long long)rows,
              acc[1] / std::max<int64_t>(nfast, 1), 0.0, acc[2] / std::max<int64_t>(nfast, 1),
              acc[3] / std::max<int64_t>(nfast, 1), acc[4] / std::max<int64_t>(nfast, 1),
              acc[5] / std::max<int64_t>(nfast, 1), (acc[6] + acc[7]) / std::max<int64_t>(nfast, 1));
    }
    if (!overlap) mark();
  }
  if (overlap) {
    BGP_CUDA(cudaEventRecord(ev_join, st2));
    BGP_CUDA(cudaStreamWaitEvent(st_main, ev_join, 0));
    mark();                                           // evs[1]: end of the strips
  }
  if (use_occ) BGP_CUDA(cudaMemcpyAsync(g_pred_slices, occ_cnt.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
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
// inputs_on_device: the three blocks are rows of a device-resident p x M sample matrix (column pitch resident_ld)
static int build_coef(const double* coef, int ncoef, const double* glob, int nglob, const double* icpt, int64_t M, int skip,
                      int nX, int ldk, cudaStream_t st, DevBuf& Cb, bool inputs_on_device, int64_t resident_ld = 0) {
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
  a.glob = (gd && nglob > 0) ? gd : nullptr;
  a.nglob = nglob;
  a.coef = cd;
  a.ncoef = ncoef;
  a.icpt_ld = inputs_on_device ? resident_ld : 1;
  a.glob_ld = inputs_on_device ? resident_ld : nglob;
  a.coef_ld = inputs_on_device ? resident_ld : ncoef;
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
  if (!coef || !region || !x || M <= 0 || G <= 0 || k < 4 || m < 1) {
    set_error("bgp_predict_sgp: bad arguments (k must be >= 4: cubic B-splines)");
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

// predict from the samples bgp_sample* left on the device; grid rows split over the node group:
// every rank summarises its block of x (all M sample columns are resident on every rank), the three G-vectors are
// assembled by the group's SUM all-reduce (one owner per row, zeros elsewhere)
static int fit_predict(bgp_fit* f, const DesignSpec& ds, int coef_row0, int ncoef, int glob_row0, int nglob, int icpt_row,
                       int skip, int nX, const double* x, int64_t G, double level, double* mean, double* plower,
                       double* pupper) {
  bgp_model* m = f->model;
  const int p = f->p;
  const int64_t M = f->samps_M;
  if (!f->samps_dev || M <= 0) {
    set_error("no resident samples: call bgp_sample / bgp_sample_draw on this fit first");
    return BGP_ERR_STATE;
  }
  if (coef_row0 < 0 || coef_row0 + ncoef > p || (nglob > 0 && (glob_row0 < 0 || glob_row0 + nglob > p)) || icpt_row >= p) {
    set_error("bgp_fit_predict: coefficient rows outside the latent vector (p = %d)", p);
    return BGP_ERR_ARG;
  }
  BGP_CUDA(cudaSetDevice(m->device));
  const int ldk = round_up(ds.ncols, 16);
  DevBuf Cb;
  BGP_TRY(build_coef(f->samps_dev + coef_row0, ncoef, nglob > 0 ? f->samps_dev + glob_row0 : nullptr, nglob,
                     icpt_row >= 0 ? f->samps_dev + icpt_row : nullptr, M, skip, nX, ldk, m->stream, Cb, true, p));
  int64_t lo = 0, hi = G;
  piece_bounds(G, m->node_rank, m->node_world, &lo, &hi);
  const int64_t Gl = hi - lo;
  std::vector<double> out((size_t)3 * G, 0.0);
  if (Gl > 0)
    BGP_TRY(predict_core(ds, Cb.as<double>(), ldk, M, x + lo, Gl, level, m->stream, out.data() + lo, out.data() + G + lo,
                         out.data() + 2 * G + lo, nullptr));
  if (m->node_world > 1) {
    DevBuf ob;
    BGP_TRY(ob.alloc((size_t)3 * G * sizeof(double)));
    BGP_CUDA(cudaMemcpyAsync(ob.p, out.data(), (size_t)3 * G * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    BGP_TRY(node_allreduce_sum(m, ob.as<double>(), (size_t)3 * G));
    BGP_CUDA(cudaMemcpyAsync(out.data(), ob.p, (size_t)3 * G * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    BGP_CUDA(cudaStreamSynchronize(m->stream));
  }
  if (mean) std::copy(out.begin(), out.begin() + G, mean);
  if (plower) std::copy(out.begin() + G, out.begin() + 2 * G, plower);
  if (pupper) std::copy(out.begin() + 2 * G, out.end(), pupper);
  return BGP_OK;
}

int bgp_fit_predict_iwp(bgp_fit* f, int coef_row0, int global_row0, int icpt_row, const double* knots, int nknots, int order,
                        int degree, const double* x, int64_t G, double level, double* mean, double* plower,
                        double* pupper) {
  if (!f || !f->model || !knots || !x || G <= 0 || nknots < 2 || order < 1 || order > 8 || degree < 0) {
    set_error("bgp_fit_predict_iwp: bad arguments");
    return BGP_ERR_ARG;
  }
  if (order <= degree) {   // R/03_post_fit.R:201-203
    set_error("Error: The degree of derivative to compute is not defined. Should consider higher order smoothing "
              "model or lower order of the derivative degree.");
    return BGP_ERR_ARG;
  }
  if (!(level > 0.0 && level < 1.0)) {
    set_error("bgp_fit_predict_iwp: level must be in (0, 1)");
    return BGP_ERR_ARG;
  }
  DesignSpec ds;
  ds.iwp = true;
  ds.order = order;
  ds.degree = degree;
  split_knots(knots, nknots, ds.kneg, ds.kpos);
  const int nB = (ds.kneg.empty() ? 0 : (int)ds.kneg.size() - 1) + (ds.kpos.empty() ? 0 : (int)ds.kpos.size() - 1);
  const int nX = order - degree;
  ds.ncols = nX + nB;
  return fit_predict(f, ds, coef_row0, nB, global_row0, global_row0 >= 0 ? order - 1 : 0, icpt_row, degree, nX, x, G, level,
                     mean, plower, pupper);
}

int bgp_fit_predict_sgp(bgp_fit* f, int coef_row0, int global_row0, int icpt_row, double a, int k, int m,
                        const double* region, int boundary, const double* x, int64_t G, double level, double* mean,
                        double* plower, double* pupper) {
  if (!f || !f->model || !region || !x || G <= 0 || k < 4 || m < 1) {
    set_error("bgp_fit_predict_sgp: bad arguments (k must be >= 4: cubic B-splines)");
    return BGP_ERR_ARG;
  }
  if (!(level > 0.0 && level < 1.0)) {
    set_error("bgp_fit_predict_sgp: level must be in (0, 1)");
    return BGP_ERR_ARG;
  }
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
  return fit_predict(f, ds, coef_row0, 3 * nb * m, global_row0, global_row0 >= 0 ? 2 * m : 0, icpt_row, 0, nX, x, G, level,
                     mean, plower, pupper);
}

int bgp_predict_last_occupancy(double* executed_slices, double* total_slices) {
  if (executed_slices) *executed_slices = g_pred_slices[0];
  if (total_slices) *total_slices = g_pred_slices[1];
  return BGP_OK;
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
