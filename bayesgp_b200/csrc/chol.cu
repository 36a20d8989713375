// chol.cu — dense p x p Cholesky, log-determinant and Newton solve for one Laplace problem.
//
// Replaces the CHOLMOD factorisation / solve TMB's inner `newton()` performs on the random-effect
// Hessian and the 1/2 logdet H term of the Laplace approximation (TMB::MakeADFun(random = "W"),
// call site /root/reference/R/02_model_fit.R:276-284).
//
// One thread-block cluster (8 CTAs x 256 threads) owns one problem; H stays in global memory
// (L2-resident: <= 8 MB).  Right-looking blocked algorithm, panel width NB, two cluster barriers per panel:
//   1. every CTA factors the NB x NB diagonal block itself (one warp, rows in registers, shuffles —
//      no CTA barrier inside the 32 dependent steps);
//   2. the panel below it is solved one row per thread, 32-row blocks dealt round-robin to the CTAs;
//   3. the trailing update C -= P P^T runs on the FP64 tensor pipe (mma.sync m8n8k4 / DMMA), 32x32
//      tiles dealt round-robin to the (CTA, warp) pairs, fragments read conflict-free from a padded
//      shared-memory copy of the panel.
// Rank 0 then does the log-determinant and the blocked forward/backward substitution for step = -H^-1 g.
// Every CTA performs the same operations on the same data in the same order: results do not depend on
// the cluster size.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "bgp_internal.h"

namespace bgp {

struct CholArgs {
  double* L;         // p x ldh column-major; on entry a copy of H (lower triangle is used)
  double* dinv;      // out: 1 / L_jj (p) — the substitutions multiply instead of divide
  int p, ldh;
  const double* g;   // gradient (p)
  double* step;      // out: -H^-1 g (p)
  EvalScalars* sc;
  int solve;
  const double* W;   // with Wtrial: the full-step trial point W + step is written as well (lda entries, zero padded)
  double* Wtrial;
  int lda;
  unsigned long long* dbg;   // BGP_CHOL_DEBUG: per-phase nanoseconds of rank 0 (9 slots)
};

constexpr int CH_THREADS = 256;   // up to 255 registers / thread: row blocks and 32x32 DMMA accumulators stay in registers
constexpr int CH_MAXP = 2048;
constexpr int CH_CS = 8;          // CTAs per cluster (portable maximum)
namespace cg = cooperative_groups;

__device__ __forceinline__ void dmma884c(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}


// x <- (L L^T)^-1 x for the vector held in shared memory (blocked forward / backward substitution,
// 32-wide diagonal blocks solved by warp 0 from a shared copy, off-diagonal updates by all threads)
__device__ void tri_forward_inplace(const double* L, const double* dinv, int p, int ldh, double* sv, double* sS) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- forward substitution, blocks of 32
  for (int b0 = 0; b0 < p; b0 += 32) {
    const int bn = (p - b0) < 32 ? (p - b0) : 32;
    for (int t = tid; t < 1024; t += CH_THREADS) {
      const int i = t & 31, c = t >> 5;
      sS[i * 33 + c] = (i < bn && c < bn && i >= c) ? __ldcg(L + (size_t)(b0 + c) * ldh + b0 + i) : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double yv = lane < bn ? sv[b0 + lane] : 0.0;
      const double di = lane < bn ? __ldcg(dinv + b0 + lane) : 0.0;
      for (int j = 0; j < bn; ++j) {
        double yj = __shfl_sync(0xffffffffu, yv, j) * __shfl_sync(0xffffffffu, di, j);
        if (lane == j) yv = yj;
        else if (lane > j) yv = fma(-yj, sS[lane * 33 + j], yv);
      }
      if (lane < bn) sv[b0 + lane] = yv;
    }
    __syncthreads();
    for (int i = b0 + bn + tid; i < p; i += CH_THREADS) {
      double s = sv[i];
#pragma unroll 8
      for (int c = 0; c < bn; ++c) s = fma(-__ldcg(L + (size_t)(b0 + c) * ldh + i), sv[b0 + c], s);
      sv[i] = s;
    }
    __syncthreads();
  }
}

// ---- backward substitution L^T x = y (y in sv) -------------------------------------------------------
__device__ void tri_backward_inplace(const double* L, const double* dinv, int p, int ldh, double* sv, double* sS,
                                     double* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (p + 31) / 32;
  for (int bi = nblk - 1; bi >= 0; --bi) {
    const int b0 = bi * 32;
    const int bn = (p - b0) < 32 ? (p - b0) : 32;
    // y_block[c] -= sum_{i >= b0+bn} L[i][b0+c] x[i]   (warp w owns columns w, w + 8, w + 16, w + 24; the four
    // dot products advance together so their loads and shuffle trees overlap, each in its own order)
    {
      constexpr int NWARP = CH_THREADS / 32;
      static_assert(4 * NWARP >= 32, "four columns per warp cover a 32-wide block");
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = b0 + bn + lane; i < p; i += 32) {
        const double xv = sv[i];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = warp + NWARP * u;
          if (c < bn) s[u] = fma(__ldcg(L + (size_t)(b0 + c) * ldh + i), xv, s[u]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (warp + NWARP * u < bn) s_red[warp + NWARP * u] = s[u];
      }
    }
    for (int t = tid; t < 1024; t += CH_THREADS) {
      const int i = t & 31, c = t >> 5;
      sS[i * 33 + c] = (i < bn && c < bn && i >= c) ? __ldcg(L + (size_t)(b0 + c) * ldh + b0 + i) : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double xv = lane < bn ? sv[b0 + lane] - s_red[lane] : 0.0;
      const double di = lane < bn ? __ldcg(dinv + b0 + lane) : 0.0;
      for (int j = bn - 1; j >= 0; --j) {
        double xj = __shfl_sync(0xffffffffu, xv, j) * __shfl_sync(0xffffffffu, di, j);
        if (lane == j) xv = xj;
        else if (lane < j) xv = fma(-xj, sS[j * 33 + lane], xv);
      }
      if (lane < bn) sv[b0 + lane] = xv;
    }
    __syncthreads();
  }
}

__device__ void tri_solve_inplace(const double* L, const double* dinv, int p, int ldh, double* sv, double* sS,
                                  double* s_red) {
  tri_forward_inplace(L, dinv, p, ldh, sv, sS);
  tri_backward_inplace(L, dinv, p, ldh, sv, sS, s_red);
}

// Cholesky of the NB x NB block in shared memory (working copy sD, row i at sD + i * DP, rows >= nb padded with
// the identity) by the whole CTA: thread (i, q) owns row i and the columns c = j + 1 + q, j + 1 + q + Q, ...
// of the rank-1 update, one CTA barrier per column.  The scaled column goes to a separate array sL, so the
// unscaled working column can still be read by the other threads of the same step (no second barrier).
// Columns are scaled with rsqrt of the pivot: L_jj = d * rsqrt(d) (within 2 ulp of sqrt(d)); 1 / L_jj =
// rsqrt(d) goes to sInv for the substitutions.  Returns 0 or j + 1 for a non-positive pivot in column j.
template <int NB>
__device__ __forceinline__ int block_chol(double* sD, double* sL, double* sInv, int tid) {
  // Two columns per barrier: every thread forms the 2 x 2 pivot block (L_jj, L_j+1,j, L_j+1,j+1) itself, scales its
  // row's two entries and applies both rank-1 updates to its trailing entries in the order the one-column
  // algorithm would (same roundings), so 16 barriers instead of 32 sit on the critical path of a panel.
  static_assert(NB % 2 == 0, "columns are eliminated in pairs");
  constexpr int DP = NB + 1;
  constexpr int Q = CH_THREADS / 32;
  const int i = tid & 31, q = tid >> 5;
  int info = 0;
  for (int j = 0; j < NB; j += 2) {
    const double d00 = sD[j * DP + j], d10 = sD[(j + 1) * DP + j], d11 = sD[(j + 1) * DP + j + 1];
    if (info == 0 && (!(d00 > 0.0) || !isfinite(d00))) info = j + 1;
    const double rs0 = rsqrt(d00);
    const double l10 = d10 * rs0;
    const double e11 = fma(-l10, l10, d11);
    if (info == 0 && (!(e11 > 0.0) || !isfinite(e11))) info = j + 2;
    const double rs1 = rsqrt(e11);
    if (i < NB && i >= j) {
      const double li0 = sD[i * DP + j] * rs0;
      const double li1 = i > j ? fma(-li0, l10, sD[i * DP + j + 1]) * rs1 : 0.0;
      if (q == 0) {
        sL[i * DP + j] = li0;
        if (i > j) sL[i * DP + j + 1] = li1;
        if (i == j) sInv[j] = rs0;
        if (i == j + 1) sInv[j + 1] = rs1;
      }
      for (int c = j + 2 + q; c <= i; c += Q) {
        const double lc0 = sD[c * DP + j] * rs0;
        const double lc1 = fma(-lc0, l10, sD[c * DP + j + 1]) * rs1;
        sD[i * DP + c] = fma(-li1, lc1, fma(-li0, lc0, sD[i * DP + c]));
      }
    }
    __syncthreads();
  }
  return info;
}

// ---- tangent predictor: T_k = d w_hat / d theta_k = -H^-1 c_k, c_k = d2 f / dW dtheta_k -------------------
// (same closed form as the implicit term of the Laplace gradient).  Used only to
// warm-start the next inner Newton solve; results do not depend on it.
struct TanBlock {
  int off, d, diag;
  const double* P;
  double etheta;
};
struct TangentArgs {
  const double* L;
  const double* dinv;
  int p, ldh, lda;
  const double* W;      // the mode
  const double* mu0;
  const double* qfix;
  int nrnd, S;
  TanBlock rnd[16];
  double* T;            // S x lda
};

// right-hand side c_k = d2 f / dW dtheta_k of tangent k into sv (length p)
__device__ __forceinline__ void tangent_rhs(const TangentArgs& a, int k, double* sv) {
  for (int i = threadIdx.x; i < a.p; i += CH_THREADS) {
    double v = 0.0;
    if (k < a.nrnd) {
      const TanBlock& rb = a.rnd[k];
      if (i >= rb.off && i < rb.off + rb.d) {
        const int r = i - rb.off;
        if (rb.diag) {
          v = rb.etheta * rb.P[r] * a.W[i];
        } else {
          double s = 0.0;
          for (int c = 0; c < rb.d; ++c) s = fma(rb.P[(size_t)c * rb.d + r], a.W[rb.off + c], s);
          v = rb.etheta * s;
        }
      }
    } else {
      // Gaussian noise theta: c = -A^T r = -Q (w_hat - mu0) at the mode
      double q = a.qfix[i] * (a.W[i] - a.mu0[i]);
      for (int b = 0; b < a.nrnd; ++b) {
        const TanBlock& rb = a.rnd[b];
        if (i >= rb.off && i < rb.off + rb.d) {
          const int r = i - rb.off;
          if (rb.diag) {
            q = rb.etheta * rb.P[r] * a.W[i];
          } else {
            double s = 0.0;
            for (int c = 0; c < rb.d; ++c) s = fma(rb.P[(size_t)c * rb.d + r], a.W[rb.off + c], s);
            q = rb.etheta * s;
          }
        }
      }
      v = -q;
    }
    sv[i] = v;
  }
}

#define CH_MARK(slot)                                                    \
  do {                                                                   \
    if (a.dbg && rank == 0 && tid == 0) {                                \
      unsigned long long _t;                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));             \
      a.dbg[slot] += _t - t_last;                                        \
      t_last = _t;                                                       \
    }                                                                    \
  } while (0)

template <int NB>
__global__ void __cluster_dims__(CH_CS, 1, 1) __launch_bounds__(CH_THREADS, 1)
    chol_kernel(const CholArgs a, const TangentArgs ta, const int ntan) {
  extern __shared__ double sm[];
  constexpr int DP = NB + 1;        // pitch of the diagonal block
  constexpr int PP = NB + 4;        // pitch of the panel (conflict-free DMMA fragment loads)
  double* sD = sm;                  // NB x DP
  double* sS = sD + 32 * 33;        // 32 x 33 scratch for the substitution phase
  double* sv = sS + 32 * 33;        // CH_MAXP solution vector
  double* sP = sv + CH_MAXP;        // panel rows x PP
  __shared__ int s_info;
  __shared__ double s_red[32];
  __shared__ double s_inv[32];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = a.p, ldh = a.ldh;
  double* L = a.L;
  if (tid == 0) s_info = 0;
  __syncthreads();
  unsigned long long t_last = 0;
  if (a.dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_last));

  // The forward substitution rides along with the factorisation: this CTA's right-hand side (rank 0: -g for the
  // Newton step; rank r: tangent r - 1) sits in shared memory, block k is solved against L_kk by an otherwise idle
  // warp as soon as the diagonal block is factored, and the rest of the vector is updated from the panel while it
  // is in shared memory for the trailing update.  Same operations in the same order as tri_forward_inplace.
  const bool has_rhs = rank == 0 ? a.solve != 0 : (rank - 1 < ntan);
  if (has_rhs) {
    if (rank == 0) {
      for (int i = tid; i < p; i += CH_THREADS) sv[i] = -a.g[i];
    } else {
      tangent_rhs(ta, rank - 1, sv);
    }
  }
  __syncthreads();

  for (int k0 = 0; k0 < p; k0 += NB) {
    const int nb = (p - k0) < NB ? (p - k0) : NB;
    const int m = p - k0 - nb;
    // ---- 1. diagonal block: every CTA factors its own copy (identical arithmetic) -------------------
    for (int t = tid; t < NB * NB; t += CH_THREADS) {
      const int i = t % NB, c = t / NB;
      sD[i * DP + c] = (i < nb && c < nb && i >= c) ? __ldcg(L + (size_t)(k0 + c) * ldh + k0 + i) : (i == c ? 1.0 : 0.0);
    }
    __syncthreads();
    CH_MARK(0);
    {
      const int info = block_chol<NB>(sD, sS, s_inv, tid);       // every thread sees the same pivots
      if (info != 0 && tid == 0) s_info = k0 + info;
    }
    __syncthreads();
    for (int t = tid; t < NB * NB; t += CH_THREADS) {             // factor back into sD for the panel solve
      const int i = t % NB, c = t / NB;
      if (c <= i) sD[i * DP + c] = sS[i * DP + c];
    }
    __syncthreads();
    if (rank == 0 && s_info == 0) {
      for (int t = tid; t < NB * NB; t += CH_THREADS) {
        const int i = t % NB, c = t / NB;
        if (i < nb && c <= i) __stcg(L + (size_t)(k0 + c) * ldh + k0 + i, sD[i * DP + c]);
      }
      if (tid < nb) __stcg(a.dinv + k0 + tid, s_inv[tid]);
    }
    CH_MARK(1);
    if (s_info != 0) break;        // same decision in every CTA of the cluster
    if (has_rhs && warp == CH_THREADS / 32 - 1) {
      // y_k = L_kk^-1 b_k (the last warp has no share in the panel solve or the trailing tiles at these sizes)
      double yv = lane < nb ? sv[k0 + lane] : 0.0;
      const double di = lane < nb ? s_inv[lane] : 0.0;
      for (int j = 0; j < nb; ++j) {
        const double yj = __shfl_sync(0xffffffffu, yv, j) * __shfl_sync(0xffffffffu, di, j);
        if (lane == j) yv = yj;
        else if (lane > j) yv = fma(-yj, sD[lane * DP + j], yv);
      }
      if (lane < nb) sv[k0 + lane] = yv;
    }
    if (m <= 0) break;
    // ---- 2. panel solve: row r of L[k0+nb.., k0..k0+NB) <- row * L_kk^-T ; 32-row blocks round-robin ---
    const int mpad = (m + 31) & ~31;
    for (int blk = rank + CH_CS * warp; blk * 32 < m; blk += CH_CS * (CH_THREADS / 32)) {
      const int t = blk * 32 + lane;
      if (t < m) {
        double x[NB];
        const int r = k0 + nb + t;
#pragma unroll
        for (int c = 0; c < NB; ++c) x[c] = c < nb ? __ldcg(L + (size_t)(k0 + c) * ldh + r) : 0.0;
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          x[c] *= s_inv[c];
#pragma unroll
          for (int q = c + 1; q < NB; ++q) x[q] = fma(-x[c], sD[q * DP + c], x[q]);
        }
#pragma unroll
        for (int c = 0; c < NB; ++c)
          if (c < nb) __stcg(L + (size_t)(k0 + c) * ldh + r, x[c]);
      }
    }
    CH_MARK(2);
    __threadfence();
    cluster.sync();
    CH_MARK(3);
    // ---- 3. the whole panel into shared memory (zero padded) -----------------------------------------
    for (int t = tid; t < mpad * NB; t += CH_THREADS) {
      const int r = t % mpad, c = t / mpad;
      sP[(size_t)r * PP + c] = (r < m && c < nb) ? __ldcg(L + (size_t)(k0 + c) * ldh + k0 + nb + r) : 0.0;
    }
    __syncthreads();
    if (has_rhs) {
      // b_trail -= L_trail,k y_k from the panel in shared memory (y_k was written before the barriers above)
      for (int i = tid; i < m; i += CH_THREADS) {
        double sacc = sv[k0 + nb + i];
#pragma unroll 8
        for (int c = 0; c < nb; ++c) sacc = fma(-sP[(size_t)i * PP + c], sv[k0 + c], sacc);
        sv[k0 + nb + i] = sacc;
      }
    }
    CH_MARK(4);
    // ---- 4. trailing update on the FP64 tensor pipe, tiles round-robin over (CTA, warp) ---------------
    const int ntd = mpad / 32;
    const int ntile = ntd * (ntd + 1) / 2;
    const int fj = lane >> 2, fk = lane & 3;
    for (int idx = rank + CH_CS * warp; idx < ntile; idx += CH_CS * (CH_THREADS / 32)) {
      // idx -> (ti >= tj)
      int ti = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
      while ((ti + 1) * (ti + 2) / 2 <= idx) ++ti;
      while (ti * (ti + 1) / 2 > idx) --ti;
      const int tj = idx - ti * (ti + 1) / 2;
      double acc[4][4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      const double* pa = sP + (size_t)(ti * 32 + fj) * PP + fk;
      const double* pb = sP + (size_t)(tj * 32 + fj) * PP + fk;
#pragma unroll
      for (int kk = 0; kk < NB / 4; ++kk) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          af[i] = pa[(size_t)(i * 8) * PP + kk * 4];
          bf[i] = pb[(size_t)(i * 8) * PP + kk * 4];
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma884c(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
      }
      const int base = k0 + nb;
      // read-modify-write of the 32 x 32 tile in two phases (all loads in flight, then all stores): the
      // compiler must otherwise order every store before the next load (possible aliasing)
      double cur[4][4][2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int row = ti * 32 + mi * 8 + fj;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = tj * 32 + ni * 8 + 2 * fk + e;
            cur[mi][ni][e] = (row < m && col <= row) ? __ldcg(L + (size_t)(base + col) * ldh + base + row) : 0.0;
          }
        }
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int row = ti * 32 + mi * 8 + fj;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = tj * 32 + ni * 8 + 2 * fk + e;
            if (row < m && col <= row) __stcg(L + (size_t)(base + col) * ldh + base + row, cur[mi][ni][e] - acc[mi][ni][e]);
          }
        }
      }
    }
    CH_MARK(5);
    __threadfence();
    cluster.sync();
    CH_MARK(6);
  }
  // no cluster barrier from here on.  Rank 0: log-determinant and Newton step; ranks 1 .. ntan: one tangent
  // d w_hat / d theta_k each (warm-start predictor), solved at the same time on their own SMs
  if (rank != 0) {
    if (rank - 1 < ntan && s_info == 0) {
      __syncthreads();
      tri_backward_inplace(L, a.dinv, p, ldh, sv, sS, s_red);
      for (int i = tid; i < ta.lda; i += CH_THREADS) ta.T[(size_t)(rank - 1) * ta.lda + i] = i < p ? -sv[i] : 0.0;
    }
    return;
  }
  __syncthreads();
  if (s_info != 0) {
    if (tid == 0) {
      a.sc->chol_info = s_info;
      a.sc->logdet = NAN;
      a.sc->smax = NAN;
    }
    return;
  }
  // ---- log det H = 2 sum log L_jj (fixed-order tree) ------------------------------------------
  {
    double s = 0.0;
    for (int j = tid; j < p; j += CH_THREADS) s += log(__ldcg(L + (size_t)j * ldh + j));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < CH_THREADS / 32; ++w) t += s_red[w];
      a.sc->logdet = 2.0 * t;
      a.sc->chol_info = 0;
    }
    __syncthreads();
  }
  CH_MARK(7);
  if (!a.solve) return;
  __syncthreads();
  tri_backward_inplace(L, a.dinv, p, ldh, sv, sS, s_red);
  double mx = 0.0;
  for (int i = tid; i < p; i += CH_THREADS) {
    const double x = sv[i];
    a.step[i] = x;
    if (a.Wtrial) a.Wtrial[i] = fma(1.0, x, a.W[i]);          // axpy_trial_kernel's t = 1 (newton.cu)
    mx = fmax(mx, isfinite(x) ? fabs(x) : INFINITY);
  }
  if (a.Wtrial)
    for (int i = p + tid; i < a.lda; i += CH_THREADS) a.Wtrial[i] = 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < CH_THREADS / 32; ++w) t = fmax(t, s_red[w]);
    a.sc->smax = t;
  }
  CH_MARK(8);
}

template <int NB>
static int launch_chol_t(bgp_model* m, const CholArgs& a, const TangentArgs& ta, int ntan) {
  const int mpad = round_up(std::max(0, a.p - NB), 32);
  const size_t smem = (size_t)(2 * 32 * 33 + CH_MAXP + (size_t)mpad * (NB + 4)) * sizeof(double);
  if (smem > 227 * 1024) {
    set_error("Cholesky panel does not fit shared memory (p = %d)", a.p);
    return BGP_ERR_ARG;
  }
  // per device and cheap: set on every launch (a process may drive models on several devices)
  BGP_CUDA(cudaFuncSetAttribute(chol_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  chol_kernel<NB><<<CH_CS, CH_THREADS, smem, m->stream>>>(a, ta, ntan);   // one cluster (compile-time __cluster_dims__)
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

static void fill_tangent_args(bgp_model* m, const double* theta, const double* W, TangentArgs& a);

// L <- chol(H) (H is left untouched), logdet, optionally step = -H^-1 g
int launch_chol_solve(bgp_model* m, bool solve, const double* theta_tan, const double* W_tan, bool write_trial) {
  m->L_is_reversed = false;
  if (m->L_holds_H) m->L_holds_H = false;      // the Hessian kernel of the moment path wrote L alongside H
  else BGP_CUDA(cudaMemcpyAsync(m->L, m->H, (size_t)m->ldh * m->p * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
  CholArgs a;
  a.W = write_trial ? m->W : nullptr;
  a.Wtrial = write_trial ? m->Wtrial : nullptr;
  a.lda = m->lda;
  a.L = m->L;
  a.dinv = m->Ldinv;
  a.p = m->p;
  a.ldh = m->ldh;
  a.g = m->g;
  a.step = m->step;
  a.sc = m->sc_dev;
  a.solve = solve ? 1 : 0;
  static unsigned long long* dbg_dev = nullptr;
  static int dbg_calls = 0;
  if (getenv("BGP_CHOL_DEBUG")) {
    if (!dbg_dev) {
      cudaMalloc(&dbg_dev, 16 * sizeof(unsigned long long));
      cudaMemset(dbg_dev, 0, 16 * sizeof(unsigned long long));
    }
    if (++dbg_calls % 40 == 0) {
      unsigned long long h[16];
      cudaStreamSynchronize(m->stream);
      cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[chol] %d calls, us/call: diag-load %.1f factor %.1f panel-solve %.1f sync1 %.1f panel-load %.1f update %.1f sync2 %.1f logdet %.1f trisolve %.1f\n",
              dbg_calls - 1, h[0] * 1e-3 / (dbg_calls - 1), h[1] * 1e-3 / (dbg_calls - 1), h[2] * 1e-3 / (dbg_calls - 1),
              h[3] * 1e-3 / (dbg_calls - 1), h[4] * 1e-3 / (dbg_calls - 1), h[5] * 1e-3 / (dbg_calls - 1),
              h[6] * 1e-3 / (dbg_calls - 1), h[7] * 1e-3 / (dbg_calls - 1), h[8] * 1e-3 / (dbg_calls - 1));
    }
  }
  a.dbg = dbg_dev;
  if (m->p > CH_MAXP) {
    set_error("p = %d exceeds the Cholesky kernel limit %d", m->p, CH_MAXP);
    return BGP_ERR_ARG;
  }
  // tangents ride along on the idle ranks of the cluster when there is one rank per theta
  TangentArgs ta;
  memset(&ta, 0, sizeof(ta));
  int ntan = 0;
  if (theta_tan && W_tan && m->S <= CH_CS - 1) {
    fill_tangent_args(m, theta_tan, W_tan, ta);
    ntan = m->S;
  }
  if (m->p <= 512) return launch_chol_t<32>(m, a, ta, ntan);
  if (m->p <= 1200) return launch_chol_t<16>(m, a, ta, ntan);
  return launch_chol_t<8>(m, a, ta, ntan);
}

__global__ void __launch_bounds__(CH_THREADS, 1) tangent_kernel(const TangentArgs a) {
  __shared__ double sS[32 * 33];
  __shared__ double sv[CH_MAXP];
  __shared__ double s_red[32];
  const int k = blockIdx.x, tid = threadIdx.x;
  tangent_rhs(a, k, sv);
  __syncthreads();
  tri_solve_inplace(a.L, a.dinv, a.p, a.ldh, sv, sS, s_red);
  for (int i = tid; i < a.lda; i += CH_THREADS) a.T[(size_t)k * a.lda + i] = i < a.p ? -sv[i] : 0.0;
}

static void fill_tangent_args(bgp_model* m, const double* theta, const double* W, TangentArgs& a) {
  a.L = m->L;
  a.dinv = m->Ldinv;
  a.p = m->p;
  a.ldh = m->ldh;
  a.lda = m->lda;
  a.W = W;
  a.mu0 = m->mu0;
  a.qfix = m->qfix;
  a.nrnd = m->J;
  a.S = m->S;
  for (int j = 0; j < m->J; ++j) {
    a.rnd[j].off = m->rnd[j].off;
    a.rnd[j].d = m->rnd[j].d;
    a.rnd[j].diag = m->rnd[j].diag ? 1 : 0;
    a.rnd[j].P = m->rnd[j].P_dev;
    a.rnd[j].etheta = std::exp(theta[j]);
  }
  a.T = m->Tan;
}

int launch_tangent(bgp_model* m, const double* theta) {
  TangentArgs a;
  fill_tangent_args(m, theta, m->Wmode, a);
  tangent_kernel<<<m->S, CH_THREADS, 0, m->stream>>>(a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

}  // namespace bgp
