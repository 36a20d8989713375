// lik.cu — the per-observation pass of the latent-Gaussian objective (HBM-streaming).
//
// Replaces, for one (W, theta):
//   eta = sum Xf*beta_f + sum X*beta + sum B*U          /root/reference/src/BayesGP.cpp:133-145
//   ll  = Gaussian / Poisson / Binomial log-likelihood   /root/reference/src/BayesGP.cpp:155-168
// and the first-order AD sweep TMB runs on them: r = d ll/d eta, w = -d2 ll/d eta2,
// c3 = d w/d eta, g_lik = A^T r.
//
// Design (sm_100a): one persistent CTA per SM, 1 producer warp + 8 or 16 consumer warps.
//   * A is observation-major (n x lda).  The producer streams it through a shared-memory ring with TMA:
//     a stage is 8 or 16 observations, copied as {64 columns x rows} boxes (512-byte lines, no swizzle) plus
//     the responses; up to ~160 KB per SM are in flight, which is what it takes to keep HBM3e busy.
//   * Column groups whose {64-observation chunk x 16-column box} cells are all structurally zero
//     (occupancy map of rowsort.cu) are neither copied nor multiplied: after the zero-pattern sort an
//     O-spline design moves ~60 % of its bytes.
//   * Designs up to 512 columns: a consumer warp owns two observations of a stage — lanes read the rows as
//     16-byte words (conflict-free), the two dot products are reduced with one transposing butterfly, the
//     likelihood terms are evaluated once per pair (even lanes: first observation, odd lanes: second), and the
//     warp accumulates its share of A^T r in registers; the 16 warps form two teams on alternate stages.
//     Wider designs: one observation per warp.  A is read from HBM exactly once per evaluation.
//   * Per-CTA partials are written out and reduced in a fixed order by finish.cu (bit-reproducible).
#include <cuda.h>

#include <cstdlib>

#include "bgp_internal.h"
#include "ptx.cuh"
#include "lik_terms.cuh"

namespace bgp {

using namespace ptx;

struct LikArgs {
  int lda;
  int64_t n;
  int64_t nchunks;      // ceil(n / 64)
  const double* W;
  const double* y;
  const double* size;
  int family;
  double tau;
  double* eta;
  double* wobs;
  double* c3;
  double* part_g;   // [gridDim.x][lda]
  double* part_s;   // [gridDim.x][4] : ll, sumsq, nonfinite, max |eta - previous eta|
  const double* rvec;   // if set: skip the likelihood and accumulate A^T rvec only (leverage term)
  const unsigned long long* occ;
};

struct LikPlan {
  CUtensorMap tmA;
};

constexpr int LK_CONSUMERS = 8;
// consumer groups: a stage holds 8 * NG observations, one per consumer warp.  Two groups for narrow designs:
// twice the bytes per TMA operation (the single producer thread issues ~1 operation per 100 clk) and twice
// the rows in flight (the per-row chain dot -> shuffle tree -> exp -> A^T r is ~600 clk).
__host__ __device__ constexpr int lk_groups(int NJ) { return NJ <= 6 ? 2 : 1; }
__host__ __device__ constexpr int lk_kb(int NJ) { return 8 * lk_groups(NJ); }
// R = observations per consumer warp and stage.  R = 2 (narrow designs): the two dot products are reduced with a
// transposing butterfly (5 adds instead of 10) and the likelihood terms — the exp / log1p chains — are evaluated
// once per pair, even lanes for the first observation, odd lanes for the second.
__host__ __device__ constexpr int lk_rows_per_warp(int NJ) { return NJ <= 8 ? 2 : 1; }
// With R = 2 the consumer warps form two teams that take alternate stages, so the warp count (and with it the
// latency the SM can hide) stays what it is with one observation per warp.
__host__ __device__ constexpr int lk_threads(int NJ, int R) { return 32 * (lk_kb(NJ) + 1) + 0 * R; }
constexpr int LK_SMEM_BUDGET = 200 * 1024;
__host__ __device__ constexpr int lk_group_bytes(int NJ) { return lk_kb(NJ) * 64 * 8; }   // one {64 columns x KB rows} box
__host__ __device__ constexpr int lk_stage_bytes(int NJ) { return NJ * lk_group_bytes(NJ) + 4 * lk_kb(NJ) * 8; }   // boxes | y | size | previous eta | pad
__host__ __device__ constexpr int lk_stages(int NJ) {
  return (LK_SMEM_BUDGET - 4096 - NJ * 512 * 2) / lk_stage_bytes(NJ) > 10 ? 10
                                                                         : (LK_SMEM_BUDGET - 4096 - NJ * 512 * 2) / lk_stage_bytes(NJ);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void lk_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One observation, one warp.  MASKED = false: the occupied column groups are the G leading ones (the usual
// case after the zero-pattern sort) — straight-line code over exactly those groups; MASKED = true: arbitrary
// group mask gm over all NJ groups.
struct RowCtx {
  uint32_t ra, w_base, aux;     // shared addresses: this lane's 16 bytes of the row / of W; y | size | previous eta
  int wrow, lane;
  int64_t row;
};
template <int NJ, int G, bool MASKED>
__device__ __forceinline__ void lik_row(const LikArgs& a, const RowCtx& c, uint32_t gm, int group_bytes, int kb, double2 (&ga)[NJ],
                                        double& ll, double& sumsq, double& dmax, int& bad) {
  double2 av[G];
  double rr;
  if (a.rvec) {
    rr = __ldg(a.rvec + c.row);
#pragma unroll
    for (int j = 0; j < G; ++j)
      av[j] = (!MASKED || ((gm >> j) & 1u)) ? lds128(c.ra + j * group_bytes) : make_double2(0.0, 0.0);
  } else {
    double s = 0.0, s2 = 0.0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      if (!MASKED || ((gm >> j) & 1u)) {
        av[j] = lds128(c.ra + j * group_bytes);
        const double2 wv = lds128(c.w_base + j * 512 + c.lane * 16);
        s = fma(av[j].x, wv.x, s);
        s2 = fma(av[j].y, wv.y, s2);
      } else {
        av[j] = make_double2(0.0, 0.0);
      }
    }
    s = warp_sum(s + s2);
    const double yv = lds64(c.aux + c.wrow * 8);
    const double sz = a.size ? lds64(c.aux + kb * 8 + c.wrow * 8) : 1.0;
    double ww, cc;
    obs_terms(a.family, a.tau, s, yv, sz, ll, sumsq, rr, ww, cc);
    if (!(isfinite(ww) && isfinite(rr) && isfinite(ll))) bad = 1;
    const double eta_old = lds64(c.aux + 2 * kb * 8 + c.wrow * 8);
    if (c.lane == 0) {
      const double d = fabs(s - eta_old);           // change of the linear predictor since the last pass
      dmax = d > dmax || !(d == d) ? (d == d ? d : INFINITY) : dmax;
      a.eta[c.row] = s;
      a.wobs[c.row] = ww;
      if (a.c3) a.c3[c.row] = cc;
    }
  }
#pragma unroll
  for (int j = 0; j < G; ++j) {
    if (!MASKED || ((gm >> j) & 1u)) {
      ga[j].x = fma(rr, av[j].x, ga[j].x);
      ga[j].y = fma(rr, av[j].y, ga[j].y);
    }
  }
}


// Two observations (stage rows wrow, wrow + 1), one warp: see lk_rows_per_warp.  ll / sumsq / dmax / bad are
// per-lane accumulators here: lane 0 carries the first observation's terms, lane 1 the second's.
template <int NJ, int G, bool MASKED>
__device__ __forceinline__ void lik_row2(const LikArgs& a, const RowCtx& c, uint32_t gm, int group_bytes, int kb,
                                         double2 (&ga)[NJ], double& ll, double& sumsq, double& dmax, int& bad) {
  double2 av0[G], av1[G];
  double r0, r1;
  const int odd = c.lane & 1;
  if (a.rvec) {
    r0 = __ldg(a.rvec + c.row);
    r1 = c.row + 1 < a.n ? __ldg(a.rvec + c.row + 1) : 0.0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const bool on = !MASKED || ((gm >> j) & 1u);
      av0[j] = on ? lds128(c.ra + j * group_bytes) : make_double2(0.0, 0.0);
      av1[j] = on ? lds128(c.ra + 512 + j * group_bytes) : make_double2(0.0, 0.0);
    }
  } else {
    double s0 = 0.0, t0 = 0.0, s1 = 0.0, t1 = 0.0;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      if (!MASKED || ((gm >> j) & 1u)) {
        av0[j] = lds128(c.ra + j * group_bytes);
        av1[j] = lds128(c.ra + 512 + j * group_bytes);
        const double2 wv = lds128(c.w_base + j * 512 + c.lane * 16);
        s0 = fma(av0[j].x, wv.x, s0);
        t0 = fma(av0[j].y, wv.y, t0);
        s1 = fma(av1[j].x, wv.x, s1);
        t1 = fma(av1[j].y, wv.y, t1);
      } else {
        av0[j] = make_double2(0.0, 0.0);
        av1[j] = make_double2(0.0, 0.0);
      }
    }
    s0 += t0;
    s1 += t1;
    // transposing butterfly: after the first exchange even lanes hold the first observation, odd lanes the second
    double s = (odd ? s1 : s0) + __shfl_xor_sync(0xffffffffu, odd ? s0 : s1, 1);
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int mine = c.wrow + odd;
    const bool valid = c.row + odd < a.n;
    const double yv = lds64(c.aux + mine * 8);
    const double sz = a.size ? lds64(c.aux + kb * 8 + mine * 8) : 1.0;
    double ll_t = 0.0, sq_t = 0.0, rr, ww, cc;
    obs_terms(a.family, a.tau, s, yv, sz, ll_t, sq_t, rr, ww, cc);
    if (valid) {
      ll += ll_t;
      sumsq += sq_t;
      if (!(isfinite(ww) && isfinite(rr) && isfinite(ll_t))) bad = 1;
    } else {
      rr = 0.0;
    }
    const double eta_old = lds64(c.aux + 2 * kb * 8 + mine * 8);
    if (c.lane < 2 && valid) {
      const double d = fabs(s - eta_old);           // change of the linear predictor since the last pass
      dmax = d > dmax || !(d == d) ? (d == d ? d : INFINITY) : dmax;
      a.eta[c.row + odd] = s;
      a.wobs[c.row + odd] = ww;
      if (a.c3) a.c3[c.row + odd] = cc;
    }
    r0 = __shfl_sync(0xffffffffu, rr, 0);
    r1 = __shfl_sync(0xffffffffu, rr, 1);
  }
#pragma unroll
  for (int j = 0; j < G; ++j) {
    if (!MASKED || ((gm >> j) & 1u)) {
      ga[j].x = fma(r0, av0[j].x, ga[j].x);
      ga[j].y = fma(r0, av0[j].y, ga[j].y);
      ga[j].x = fma(r1, av1[j].x, ga[j].x);
      ga[j].y = fma(r1, av1[j].y, ga[j].y);
    }
  }
}

// dispatch on the number of leading occupied groups (1 .. NJ); anything else takes the masked body
template <int NJ, int G, int R>
__device__ __forceinline__ void lik_row_dispatch(int g, const LikArgs& a, const RowCtx& c, int group_bytes, int kb,
                                                 double2 (&ga)[NJ], double& ll, double& sumsq, double& dmax, int& bad) {
  if constexpr (G >= 1) {
    if (g == G) {
      if constexpr (R == 2) lik_row2<NJ, G, false>(a, c, 0u, group_bytes, kb, ga, ll, sumsq, dmax, bad);
      else lik_row<NJ, G, false>(a, c, 0u, group_bytes, kb, ga, ll, sumsq, dmax, bad);
    } else {
      lik_row_dispatch<NJ, G - 1, R>(g, a, c, group_bytes, kb, ga, ll, sumsq, dmax, bad);
    }
  }
}

// smem: [stages][ NJ boxes | y[8] | size[8] ] | W (NJ * 64) | meta[stages] | full[stages] | empty[stages]
template <int NJ, int R>
__global__ void __launch_bounds__(lk_threads(NJ, R), 1) lik_kernel(const __grid_constant__ CUtensorMap tmA, const LikArgs a) {
  constexpr int STAGES = lk_stages(NJ);
  constexpr int STAGE_BYTES = lk_stage_bytes(NJ);
  constexpr int LK_KB = lk_kb(NJ);
  constexpr int TEAM = LK_KB / R;              // warps that share a stage
  constexpr int NCW = TEAM * R;                // consumer warps: R teams on alternate stages
  constexpr int LK_GROUP_BYTES = lk_group_bytes(NJ);
  constexpr int SPC = 64 / LK_KB;              // stages per 64-observation chunk
  constexpr int LK_THREADS = lk_threads(NJ, R);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t w_base = base + STAGES * STAGE_BYTES;
  const uint32_t meta_base = w_base + NJ * 512;
  const uint32_t full_base = meta_base + 8 * STAGES;
  const uint32_t empty_base = full_base + 8 * STAGES;
  double* sm_gen = reinterpret_cast<double*>(smem_raw + (base - smem_u32(smem_raw)));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lda = a.lda;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_base + 8 * s, 1);
      mbar_init(empty_base + 8 * s, TEAM);
    }
    mbar_fence_init();
  }
  // W into shared memory (zero padded to NJ * 64)
  for (int c = tid; c < NJ * 64; c += LK_THREADS) sm_gen[(w_base - base) / 8 + c] = (c < lda && a.W) ? a.W[c] : 0.0;
  __syncthreads();

  // 64-observation chunks (SPC stages) are dealt round-robin to the CTAs (cheap early chunks and dense late
  // chunks mix); local stage `it` of a CTA is stage (it % SPC) of its chunk number (it / SPC)
  // (32-bit counters: chunk numbers and a CTA's stage count are far below 2^31, and 64-bit div / mod by the ring
  // constants would cost more integer instructions per stage than the likelihood terms)
  const int first = (int)blockIdx.x, step = (int)gridDim.x;
  const int nchunks = (int)a.nchunks;
  const int my_chunks = first < nchunks ? (nchunks - first + step - 1) / step : 0;
  const int my_stages = my_chunks * SPC;
  if (warp == NCW) {
    if (lane != 0) return;
    unsigned long long o_next = my_chunks > 0 ? __ldg(a.occ + first) : 0ull;
    uint32_t gm = 0;
    int slot = 0;
    uint32_t par = 1;                         // parity to wait for on the empty barrier of `slot`
    for (int it = 0; it < my_stages; ++it) {
      const int chunk = first + (it / SPC) * step;
      const int64_t row = (int64_t)chunk * 64 + (it % SPC) * LK_KB;
      if ((it % SPC) == 0) {
        const unsigned long long o = o_next;
        if ((it / SPC) + 1 < my_chunks) o_next = __ldg(a.occ + chunk + step);   // prefetch: consumed a chunk later
        gm = 0;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if ((o >> (4 * j)) & 0xfull) gm |= 1u << j;
      }
      const uint32_t fb = full_base + 8 * slot;
      const uint32_t sb = base + slot * STAGE_BYTES;
      mbar_wait(empty_base + 8 * slot, par);
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(meta_base + 8 * slot), "r"(gm) : "memory");
      const bool want_y = a.rvec == nullptr;
      mbar_expect_tx(fb, (uint32_t)__popc(gm) * LK_GROUP_BYTES + (want_y ? 2 * LK_KB * 8 : 0) + (want_y && a.size ? LK_KB * 8 : 0));
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if ((gm >> j) & 1u) tma_load_2d(sb + j * LK_GROUP_BYTES, &tmA, 64 * j, (int)row, fb);
      if (want_y) {
        bulk_load_1d(sb + NJ * LK_GROUP_BYTES, a.y + row, LK_KB * 8, fb);
        if (a.size) bulk_load_1d(sb + NJ * LK_GROUP_BYTES + LK_KB * 8, a.size + row, LK_KB * 8, fb);
        bulk_load_1d(sb + NJ * LK_GROUP_BYTES + 2 * LK_KB * 8, a.eta + row, LK_KB * 8, fb);   // previous eta
      }
      if (++slot == STAGES) {
        slot = 0;
        par ^= 1u;
      }
    }
    return;
  }

  // ---- consumers: warp w of team t handles observations w * R .. w * R + R - 1 of stages t, t + R, ... ----------
  const int team = warp / TEAM;
  const int wrow = (warp % TEAM) * R;          // first observation of the stage this warp owns
  double2 ga[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) ga[j] = make_double2(0.0, 0.0);
  double ll = 0.0, sumsq = 0.0, dmax = 0.0;
  int bad = 0;
  static_assert(R <= STAGES, "a team's first stage must lie in the first lap of the ring");
  int slot = team;
  uint32_t par = 0;                           // parity to wait for on the full barrier of `slot`
  for (int it = team; it < my_stages; it += R) {
    mbar_wait(full_base + 8 * slot, par);
    const int64_t row = (int64_t)(first + (it / SPC) * step) * 64 + (it % SPC) * LK_KB + wrow;
    uint32_t gm;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(gm) : "r"(meta_base + 8 * slot) : "memory");
    const uint32_t sb = base + slot * STAGE_BYTES;
    const uint32_t ra = sb + wrow * 512 + lane * 16;
    if (row < a.n) {
      RowCtx c;
      c.ra = ra;
      c.w_base = w_base;
      c.aux = sb + NJ * LK_GROUP_BYTES;
      c.wrow = wrow;
      c.lane = lane;
      c.row = row;
      const int g = 32 - __clz(gm);                       // groups 0 .. g-1 occupied <=> gm == 2^g - 1
      if (NJ <= 8 && gm != 0u && gm == ((1u << g) - 1u))
        lik_row_dispatch<NJ, (NJ <= 8 ? NJ : 1), R>(g, a, c, LK_GROUP_BYTES, LK_KB, ga, ll, sumsq, dmax, bad);
      else if constexpr (R == 2)
        lik_row2<NJ, NJ, true>(a, c, gm, LK_GROUP_BYTES, LK_KB, ga, ll, sumsq, dmax, bad);
      else
        lik_row<NJ, NJ, true>(a, c, gm, LK_GROUP_BYTES, LK_KB, ga, ll, sumsq, dmax, bad);
    }
    __syncwarp();
    if (lane == 0) lk_arrive(empty_base + 8 * slot);
    slot += R;
    if (slot >= STAGES) {
      slot -= STAGES;
      par ^= 1u;
    }
  }
  // ---- block reduction in a fixed order (the ring is drained: reuse its memory) -------------------------
  asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
  double* sg = sm_gen + (size_t)warp * (NJ * 64);
#pragma unroll
  for (int j = 0; j < NJ; ++j) *reinterpret_cast<double2*>(sg + 64 * j + 2 * lane) = ga[j];
  double* ss = sm_gen + (size_t)NCW * (NJ * 64);
  if (R == 2) {                                // lane 0: first observations, lane 1: second ones (fixed order)
    ll += __shfl_sync(0xffffffffu, ll, 1);
    sumsq += __shfl_sync(0xffffffffu, sumsq, 1);
    bad |= __shfl_sync(0xffffffffu, bad, 1);
    dmax = fmax(dmax, __shfl_sync(0xffffffffu, dmax, 1));
  }
  if (lane == 0) {
    ss[warp * 4 + 0] = ll;
    ss[warp * 4 + 1] = sumsq;
    ss[warp * 4 + 2] = (double)bad;
    ss[warp * 4 + 3] = dmax;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
  for (int c = tid; c < lda; c += NCW * 32) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < NCW; ++w8) s += sm_gen[(size_t)w8 * (NJ * 64) + c];
    a.part_g[(size_t)blockIdx.x * lda + c] = s;
  }
  if (tid < 3) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < NCW; ++w8) s += ss[w8 * 4 + tid];
    a.part_s[(size_t)blockIdx.x * 4 + tid] = s;
  }
  if (tid == 3) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < NCW; ++w8) s = fmax(s, ss[w8 * 4 + 3]);
    a.part_s[(size_t)blockIdx.x * 4 + 3] = s;
  }
}

template <int NJ, int R>
static int launch_lik_tr(bgp_model* m, const LikArgs& a) {
  constexpr int smem = lk_stages(NJ) * lk_stage_bytes(NJ) + NJ * 512 + 16 * lk_stages(NJ) + 256;
  static_assert(lk_stages(NJ) >= 2, "likelihood ring needs at least two stages");
  static_assert(LK_CONSUMERS * lk_groups(NJ) * (NJ * 64 * 8 + 32) <= lk_stages(NJ) * lk_stage_bytes(NJ), "reduction scratch must fit in the ring");
  // per device and cheap: set on every launch (a process may drive models on several devices)
  BGP_CUDA(cudaFuncSetAttribute(lik_kernel<NJ, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const LikPlan* pl = (const LikPlan*)m->lik_plan;
  lik_kernel<NJ, R><<<m->lik_blocks, lk_threads(NJ, R), smem, m->stream>>>(pl->tmA, a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}
template <int NJ>
static int launch_lik_t(bgp_model* m, const LikArgs& a) {
  static const bool one_row = getenv("BGP_LIK_R1") != nullptr;      // env: diagnostics (one observation per warp)
  if constexpr (lk_rows_per_warp(NJ) == 2) {
    if (!one_row) return launch_lik_tr<NJ, 2>(m, a);
  }
  return launch_lik_tr<NJ, 1>(m, a);
}

int lik_max_lda() { return 1024; }

int lik_plan_create(bgp_model* m) {
  LikPlan* pl = new LikPlan();
  m->lik_plan = pl;
  if (make_tensormap_f64(&pl->tmA, m->A, (uint64_t)m->lda, (uint64_t)m->n, (uint64_t)m->lda, 64,
                         (uint32_t)lk_kb((m->lda + 63) / 64), false) != 0) {
    set_error("cuTensorMapEncodeTiled failed for the likelihood kernel");
    return BGP_ERR_CUDA;
  }
  return BGP_OK;
}

void lik_plan_destroy(bgp_model* m) {
  delete (LikPlan*)m->lik_plan;
  m->lik_plan = nullptr;
}

int launch_lik(bgp_model* m, const double* W_dev, bool want_c3, double tau, const double* rvec) {
  if (m->osp_on && !rvec) return osp_launch_lik(m, W_dev, tau);
  LikArgs a;
  a.rvec = rvec;
  a.lda = m->lda;
  a.n = m->n;
  a.nchunks = m->nchunks;
  a.W = W_dev;
  a.y = m->y;
  a.size = m->family == BGP_FAMILY_BINOMIAL ? m->size : nullptr;
  a.family = m->family;
  a.tau = tau;
  a.eta = m->eta;
  a.wobs = m->wobs;
  a.c3 = want_c3 ? m->c3 : nullptr;
  a.part_g = m->part_g;
  a.part_s = m->part_s;
  a.occ = (const unsigned long long*)m->occ_dev;
  const int nj = (m->lda + 63) / 64;
  switch (nj) {
    case 1: return launch_lik_t<1>(m, a);
    case 2: return launch_lik_t<2>(m, a);
    case 3: return launch_lik_t<3>(m, a);
    case 4: return launch_lik_t<4>(m, a);
    case 5: return launch_lik_t<5>(m, a);
    case 6: return launch_lik_t<6>(m, a);
    case 7: return launch_lik_t<7>(m, a);
    case 8: return launch_lik_t<8>(m, a);
    case 9: case 10: return launch_lik_t<10>(m, a);
    case 11: case 12: return launch_lik_t<12>(m, a);
    case 13: case 14: return launch_lik_t<14>(m, a);
    case 15: case 16: return launch_lik_t<16>(m, a);
    default:
      set_error("latent dimension p = %d exceeds the supported maximum of %d", m->p, lik_max_lda());
      return BGP_ERR_ARG;
  }
}

}  // namespace bgp
