// syrk.cu — H = A^T diag(w) A + Q(theta): the dense FP64 contraction of every Newton iteration.
//
// Replaces the sparse-Hessian AD sweep TMB runs on objective_function::operator()
// (/root/reference/src/BayesGP.cpp:30-253; ff$env$spHess(random = TRUE), call site
// /root/reference/R/02_model_fit.R:276-284).  n*p*(p+1) flops per evaluation when A is dense.
//
// Design (sm_100a):
//   * FP64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — tcgen05 has no FP64 kind; the larger f64 mma
//     shapes lower to the same instruction).  Measured issue rate: one DMMA per 16.1 clk per SM sub-partition
//     = 37.0 TFLOP/s at 1965 MHz (scripts/ubench/dmma_bench.cu).
//   * Persistent, warp-specialised CTAs (3 per SM): one producer warp + 4 consumer warps.  A CTA tile is
//     up to 4 row boxes x one N panel (64 x 64 of H); consumer warp w owns column box w of the panel and runs over
//     the tile's row boxes (a 64 x 16 strip, 32 accumulator doubles per thread).  The diag(w) scaling is
//     applied to the warp's own column-box fragments only (2 DMUL per 16 DMMA — a DMUL costs ~4.7 clk of the
//     same pipe).
//     (Round 2 measured the alternative with 64 x 128 tiles and a pair of column boxes per warp, 2 CTAs per SM:
//     41 % fewer shared-memory reads per DMMA but 2.18 ms instead of 1.27 ms — ncu: 27.6 % of the stall samples
//     `no_instruction` (45 unrolled mask variants overflow the instruction cache), 18 % full-barrier waits, tensor
//     pipe 49 % busy; shared-memory wavefronts of THIS layout are only 31 % of peak, so they were never the limit.)
//   * Operands are TMA-staged: A is observation-major, so a TMA box of {16 columns, 16 observations} lands as
//     16 lines of 128 B with the hardware 128B swizzle; a stage is 4 + 4 boxes + 16 weights.  The producer
//     warp runs a 4-stage mbarrier ring ahead of the consumers (full / empty barriers, no CTA-wide barrier).
//     Fragment loads are LDS.128 with a column permutation chosen so the swizzled lines are read
//     conflict-free; the permutation is undone in the epilogue.
//   * Work list: the unit of work is (tile, 64-observation chunk).  Cells of the {chunk x column-box}
//     occupancy map (rowsort.cu) that are structurally zero are skipped — whole chunks by leaving them out of
//     the list, single boxes by a per-stage 4-bit mask that selects a compile-time-specialised loop body
//     (a predicated-off DMMA still occupies the pipe, so skipping has to be a branch).  Boxes above the
//     diagonal or outside the lda x lda matrix are masked the same way.
//   * Scheduling: the work list is cut into units = (tile, range of chunks) of roughly equal cost, ordered
//     range-major (units that read the same observations run at the same time and share them through L2),
//     with finer units at the end of the queue.  Persistent CTAs take units from an atomic counter, so the
//     load balances whatever the co-resident CTAs do; every unit owns a private partial slot and the
//     slots of a tile are summed in a fixed order, so the result does not depend on which CTA ran what
//     (bit-reproducible).  The sum is mirrored to the upper triangle, then Q(theta) is added.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "bgp_internal.h"
#include "ptx.cuh"

namespace bgp {

#ifndef SK_MBOX_N
#define SK_MBOX_N 4
#endif
static_assert(SK_MBOX_N == 4, "the consumer layout (one warp per N box, four row boxes) assumes 4 x 4 box tiles");
constexpr int SK_MBOX = SK_MBOX_N;         // strips (16-row boxes of H) per CTA tile = consumer warps
constexpr int SK_NBOX = 4;                 // N panel: 4 boxes = 64 columns of H
constexpr int SK_KB = 16;                  // observations per pipeline stage
constexpr int SK_SPC = 4;                  // stages per 64-observation chunk
constexpr int SK_STAGES = 4;
constexpr int SK_CTAS_PER_SM = 3;
constexpr int SK_CONSUMERS = SK_NBOX;     // one consumer warp per box of the N panel
constexpr int SK_THREADS = 32 * (SK_CONSUMERS + 1);
constexpr int SK_BOX_BYTES = 16 * SK_KB * 8;                       // 2048
constexpr int SK_STAGE_BYTES = (SK_MBOX + SK_NBOX) * SK_BOX_BYTES;
constexpr int SK_W_BYTES = SK_KB * 8;                              // 128
constexpr int SK_TILE_ELEMS = 16 * SK_MBOX * 16 * SK_NBOX;
constexpr int SK_SMEM = SK_STAGES * (SK_STAGE_BYTES + SK_W_BYTES + 16) + 16 * SK_STAGES + 1024;

// A CTA tile = one N panel (columns 64J .. 64J+63 of H) against up to SK_MBOX strips; strip w is the
// 16-row box `rows[w]` of H (any box on or below the panel's diagonal — strips need not be adjacent,
// every box is its own TMA copy).  Strips are grouped so that the warps of a tile carry equal work.
struct SkTile {
  int J;
  int slot_off, slot_cnt;   // partial slots of this tile in tile_slots
  uint32_t smask;           // bit 4*s + b: box (rows[s], 4J + b) is on or below the diagonal and inside the matrix
  int8_t rows[8];           // box row of row slot s, -1 = unused
  uint8_t wcol[4];          // consumer warp w works on column box wcol[w] of the N panel ...
  uint8_t wrows[4];         // ... against the row slots in this mask (every live box has exactly one owner)
};
static_assert(sizeof(SkTile) == 32, "SkTile layout");

struct SyrkPlan {
  CUtensorMap tmA;
  int ntiles = 0, G = 0, nslots = 0;
  SkTile* tiles_dev = nullptr;
  uint32_t* entries_dev = nullptr; // tile << 24 | chunk
  int2* units_dev = nullptr;       // per unit: [entry begin, entry end)
  int nunits = 0;
  int* counter_dev = nullptr;      // work queue head (reset before every launch)
  int* tile_slots_dev = nullptr;   // CSR of partial slots per tile
  double useful_flops = 0.0;       // structurally non-zero flops per launch
  int64_t nentries = 0;
  unsigned long long* dbg_dev = nullptr;   // BGP_SK_DEBUG: per-CTA start / end timestamps
  std::vector<int64_t> unit_cost_dbg;
  std::vector<int> unit_tile_dbg;
  int dbg_left = 0;
};

using namespace ptx;

// column permutation inside a 16-column box: fragment row j reads 16-byte chunk ch(j)
__host__ __device__ __forceinline__ int sy_chunk(int j) { return (j >> 1) + 4 * (j & 1); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t lo, uint32_t hi) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// one pipeline stage (16 observations) of a warp's 64 x 16 column strip: the warp owns one box of the N panel
// (its weighted fragments are formed once per k-step) and runs over the tile's row boxes; MASK selects them
template <int MASK>
__device__ __forceinline__ void sk_stage(double (&acc)[4][2][2][2], const uint32_t (&pa)[4], uint32_t pb, uint32_t pw,
                                         int fk, int ch) {
#pragma unroll
  for (int kk = 0; kk < SK_KB / 4; ++kk) {
    const int row = kk * 4 + fk;
    const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
    double2 b = lds128(pb + off);
    const double wk = lds64(pw + row * 8);
    b.x *= wk;
    b.y *= wk;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (MASK & (1 << s)) {
        const double2 a = lds128(pa[s] + off);
        dmma884(acc[s][0][0][0], acc[s][0][0][1], a.x, b.x);
        dmma884(acc[s][0][1][0], acc[s][0][1][1], a.x, b.y);
        dmma884(acc[s][1][0][0], acc[s][1][0][1], a.y, b.x);
        dmma884(acc[s][1][1][0], acc[s][1][1][1], a.y, b.y);
      }
    }
  }
}

// per-chunk activity of a tile: which strips are present in this chunk (and not aliased to the N panel =>
// need their own copy), which N boxes any present strip wants
__host__ __device__ __forceinline__ void sk_chunk_masks(const SkTile& t, unsigned long long occ, uint32_t& act_m,
                                                        uint32_t& load_m, uint32_t& need_n) {
  const uint32_t ncol = (uint32_t)(occ >> (4 * t.J)) & 0xfu;
  act_m = load_m = need_n = 0;
#pragma unroll
  for (int w = 0; w < SK_MBOX; ++w) {
    const int r = t.rows[w];
    if (r < 0) continue;
    const uint32_t sm = (t.smask >> (4 * w)) & 0xfu & ncol;
    if (((occ >> r) & 1ull) && sm) {
      act_m |= 1u << w;
      need_n |= sm;
      if (r < 4 * t.J || r >= 4 * t.J + 4) load_m |= 1u << w;
    }
  }
  // a strip that lives inside the N panel reads the panel's own copy of its box
#pragma unroll
  for (int w = 0; w < SK_MBOX; ++w) {
    const int r = t.rows[w];
    if (((act_m >> w) & 1u) && r >= 4 * t.J && r < 4 * t.J + 4) need_n |= 1u << (r - 4 * t.J);
  }
}

__global__ void __launch_bounds__(SK_THREADS, SK_CTAS_PER_SM)
    syrk_kernel(const __grid_constant__ CUtensorMap tmA, const double* __restrict__ wobs, double* __restrict__ part,
                const SkTile* __restrict__ tiles, const uint32_t* __restrict__ entries, const int2* __restrict__ units,
                int nunits, int* __restrict__ counter, const unsigned long long* __restrict__ occ,
                unsigned long long* __restrict__ dbg, int rot_units) {
  extern __shared__ uint8_t smem_raw[];
  if (dbg && threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    dbg[2 * blockIdx.x] = t0;
  }
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base + SK_STAGES * SK_STAGE_BYTES;
  const uint32_t meta_base = w_base + SK_STAGES * SK_W_BYTES;      // 16 B per stage: {masks, tile, unit, -}
  const uint32_t full_base = meta_base + 16 * SK_STAGES;
  const uint32_t empty_base = full_base + 8 * SK_STAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < SK_STAGES; ++s) {
      mbar_init(full_base + 8 * s, 1);
      mbar_init(empty_base + 8 * s, SK_CONSUMERS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == SK_CONSUMERS) {
    // ---------------- producer warp: one elected lane takes units from the queue, issues the TMA boxes
    if (lane != 0) return;
    int it = 0;
    int unit = atomicAdd(counter, 1);
    while (unit < nunits) {
      const int unit_next = atomicAdd(counter, 1);      // in flight while this unit is being issued
      if (dbg) {
        unsigned long long tu;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tu));
        dbg[2 * gridDim.x + 2 * unit] = tu;
        dbg[2 * gridDim.x + 2 * unit + 1] = blockIdx.x;
      }
      const int2 ur = units[unit];
      const int e_begin = ur.x, e_end = ur.y;
      // software prefetch: the entry two chunks ahead, its tile descriptor and occupancy word one chunk ahead
      // (three dependent L2 round trips would otherwise sit in front of every chunk)
      uint32_t ent_n = __ldg(entries + e_begin);
      uint32_t ent_nn = e_begin + 1 < e_end ? __ldg(entries + e_begin + 1) : 0u;
      SkTile t_n = tiles[ent_n >> 24];
      unsigned long long occ_n = __ldg(occ + (ent_n & 0xffffffu));
      for (int e = e_begin; e < e_end; ++e) {
        const uint32_t ent = ent_n;
        const int tile = (int)(ent >> 24), chunk = (int)(ent & 0xffffffu);
        const SkTile t = t_n;
        const unsigned long long occ_c = occ_n;
        if (e + 1 < e_end) {
          ent_n = ent_nn;
          t_n = tiles[ent_n >> 24];
          occ_n = __ldg(occ + (ent_n & 0xffffffu));
          if (e + 2 < e_end) ent_nn = __ldg(entries + e + 2);
        }
        uint32_t act_m, load_m, need_n;
        sk_chunk_masks(t, occ_c, act_m, load_m, need_n);
        const uint32_t tx = (uint32_t)(__popc(load_m) + __popc(need_n)) * SK_BOX_BYTES + SK_W_BYTES;
        const int ncol0 = 4 * t.J * 16;
        for (int s = 0; s < SK_SPC; ++s, ++it) {
          const int slot = it % SK_STAGES;
          const uint32_t fb = full_base + 8 * slot;
          const uint32_t sb = base + slot * SK_STAGE_BYTES;
          mbar_wait(empty_base + 8 * slot, (uint32_t)(((it / SK_STAGES) & 1) ^ 1));
          const int row = chunk * 64 + s * SK_KB;
          sts128(meta_base + 16 * slot, act_m | (need_n << 8), (uint32_t)tile, (uint32_t)unit, 0u);   // rows | N boxes present
          mbar_expect_tx(fb, tx);
#pragma unroll
          for (int b = 0; b < SK_MBOX; ++b)
            if ((load_m >> b) & 1u) tma_load_2d(sb + b * SK_BOX_BYTES, &tmA, 16 * t.rows[b], row, fb);
#pragma unroll
          for (int b = 0; b < SK_NBOX; ++b)
            if ((need_n >> b) & 1u) tma_load_2d(sb + (SK_MBOX + b) * SK_BOX_BYTES, &tmA, ncol0 + b * 16, row, fb);
          bulk_load_1d(w_base + slot * SK_W_BYTES, wobs + row, SK_W_BYTES, fb);
        }
      }
      unit = unit_next;
    }
    // end of queue: one empty stage tells the consumers to flush and leave
    {
      const int slot = it % SK_STAGES;
      mbar_wait(empty_base + 8 * slot, (uint32_t)(((it / SK_STAGES) & 1) ^ 1));
      sts128(meta_base + 16 * slot, 0u, 0u, 0xffffffffu, 1u);
      mbar_arrive(full_base + 8 * slot);
    }
    return;
  }

  // ---------------- consumer warps: warp w owns column box w of the N panel ----------------------------
  int wc = warp;                 // column box of the N panel this warp works on (per tile)
  const int fj = lane >> 2, fk = lane & 3;
  const int ch = sy_chunk(fj);
  double acc[4][2][2][2];
  auto zero_acc = [&]() {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[a][b][c][0] = acc[a][b][c][1] = 0.0;
  };
  zero_acc();
  int cur_unit = -1;
  uint32_t smask_w = 0;       // bit s: box (rows[s], 4J + wc) belongs to the lower triangle
  uint32_t a_off[4] = {0, 0, 0, 0};
  auto flush = [&]() {
    // undo the column permutation, write this column strip of the partial tile (row-major [16 * s + m][n])
    double* out = part + (size_t)cur_unit * SK_TILE_ELEMS;
#pragma unroll
    for (int sr = 0; sr < 4; ++sr) {
      if (!((smask_w >> sr) & 1u)) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int M = sr * 16 + 2 * ch + e;
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int N = wc * 16 + 2 * sy_chunk(2 * fk + c) + f;
            out[M * (16 * SK_NBOX) + N] = acc[sr][e][f][c];
          }
      }
    }
  };
  for (int it = 0;; ++it) {
    const int slot = it % SK_STAGES;
    mbar_wait(full_base + 8 * slot, (uint32_t)((it / SK_STAGES) & 1));
    const uint4 meta = lds_u4(meta_base + 16 * slot);
    const int unit = (int)meta.z;
    if (meta.w) {                 // end of queue
      if (cur_unit >= 0) flush();
      break;
    }
    if (unit != cur_unit) {
      if (cur_unit >= 0) {
        flush();
        zero_acc();
      }
      cur_unit = unit;
      const SkTile* tp = tiles + meta.y;
      const int j4 = 4 * tp->J;
      // Column boxes are dealt to the warps with a per-unit rotation.  After the zero-pattern sort the occupied
      // column boxes of a chunk are a prefix, so box 0 of a panel carries ~31 % of a tile's DMMAs and box 3 ~20 %
      // (C3); warp w of every CTA sits on SM sub-partition w, each with its own FP64 tensor pipe, so a fixed
      // box -> warp map keeps one pipe saturated while another idles a third of the time.
      const int wi = (warp + (rot_units ? unit : 0)) & 3;
      smask_w = tp->wrows[wi];
      wc = tp->wcol[wi];
#pragma unroll
      for (int sr = 0; sr < 4; ++sr) {
        const int r = tp->rows[sr];
        // a row box is read from its private copy, or from the N panel's copy when it lies inside the panel
        a_off[sr] = (r >= j4 && r < j4 + 4) ? (uint32_t)(SK_MBOX + (r - j4)) * SK_BOX_BYTES : (uint32_t)sr * SK_BOX_BYTES;
      }
    }
    const uint32_t mask = ((meta.x >> (8 + wc)) & 1u) ? (smask_w & meta.x & 0xfu) : 0u;
    const uint32_t sb = base + slot * SK_STAGE_BYTES;
    const uint32_t pa[4] = {sb + a_off[0], sb + a_off[1], sb + a_off[2], sb + a_off[3]};
    const uint32_t pb = sb + (SK_MBOX + wc) * SK_BOX_BYTES;
    const uint32_t pw = w_base + slot * SK_W_BYTES;
    switch (mask) {
      case 1: sk_stage<1>(acc, pa, pb, pw, fk, ch); break;
      case 2: sk_stage<2>(acc, pa, pb, pw, fk, ch); break;
      case 3: sk_stage<3>(acc, pa, pb, pw, fk, ch); break;
      case 4: sk_stage<4>(acc, pa, pb, pw, fk, ch); break;
      case 5: sk_stage<5>(acc, pa, pb, pw, fk, ch); break;
      case 6: sk_stage<6>(acc, pa, pb, pw, fk, ch); break;
      case 7: sk_stage<7>(acc, pa, pb, pw, fk, ch); break;
      case 8: sk_stage<8>(acc, pa, pb, pw, fk, ch); break;
      case 9: sk_stage<9>(acc, pa, pb, pw, fk, ch); break;
      case 10: sk_stage<10>(acc, pa, pb, pw, fk, ch); break;
      case 11: sk_stage<11>(acc, pa, pb, pw, fk, ch); break;
      case 12: sk_stage<12>(acc, pa, pb, pw, fk, ch); break;
      case 13: sk_stage<13>(acc, pa, pb, pw, fk, ch); break;
      case 14: sk_stage<14>(acc, pa, pb, pw, fk, ch); break;
      case 15: sk_stage<15>(acc, pa, pb, pw, fk, ch); break;
      default: break;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_base + 8 * slot);
  }
  if (dbg && warp == 0 && lane == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    dbg[2 * blockIdx.x + 1] = t1;
  }
}

// sum the partial slots of each tile in a fixed order, write the lower triangle and its mirror.
// 64 elements x 8 slot lanes per block: lane q adds slots q, q + 8, ... (independent loads in flight), the
// eight partial sums are then combined in lane order, so the result is independent of the launch shape.
constexpr int SR_ELEMS = 64, SR_LANES = 8;
__global__ void __launch_bounds__(SR_ELEMS * SR_LANES)
    syrk_reduce_kernel(const double* __restrict__ part, const SkTile* __restrict__ tiles, const int* __restrict__ tile_slots,
                       int p, int ldh, double* __restrict__ H) {
  __shared__ double sm[SR_LANES][SR_ELEMS + 1];
  const int tile = blockIdx.x / (SK_TILE_ELEMS / SR_ELEMS);
  const int el = threadIdx.x % SR_ELEMS, q = threadIdx.x / SR_ELEMS;
  const int e = (blockIdx.x % (SK_TILE_ELEMS / SR_ELEMS)) * SR_ELEMS + el;
  const int M = e / (16 * SK_NBOX), N = e % (16 * SK_NBOX);
  const SkTile t = tiles[tile];
  const int r = t.rows[M / 16];
  const int gr = r * 16 + (M % 16), gc = t.J * 16 * SK_NBOX + N;
  const bool live = r >= 0 && gr < p && gc < p && gc <= gr;
  double s = 0.0;
  if (live) {
    const int* sl = tile_slots + t.slot_off;
    for (int k = q; k < t.slot_cnt; k += SR_LANES) s += part[(size_t)sl[k] * SK_TILE_ELEMS + e];
  }
  sm[q][el] = s;
  __syncthreads();
  if (q == 0 && live) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < SR_LANES; ++k) tot += sm[k][el];
    H[(size_t)gc * ldh + gr] = tot;
    H[(size_t)gr * ldh + gc] = tot;
  }
}

struct AddQArgs {
  double* H;
  int ldh, p;
  const double* qfix;
  int nrnd;
  struct {
    int off, d, diag;
    const double* P;
    double etheta;
  } rnd[16];
};

__global__ void add_q_kernel(const AddQArgs a) {
  // block (bx) handles one random block (bx < nrnd) or the fixed diagonal (bx == nrnd)
  const int b = blockIdx.y;
  if (b == a.nrnd) {
    // only the boundary / fixed-effect entries: the spline diagonals belong to the other blocks
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.p; c += gridDim.x * blockDim.x) {
      const double q = a.qfix[c];                    // zero on the spline entries, which the other blocks own
      if (q != 0.0) a.H[(size_t)c * a.ldh + c] += q;
    }
    return;
  }
  const int off = a.rnd[b].off, d = a.rnd[b].d;
  const double et = a.rnd[b].etheta;
  if (a.rnd[b].diag) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d; c += gridDim.x * blockDim.x)
      a.H[(size_t)(off + c) * a.ldh + off + c] += et * a.rnd[b].P[c];
  } else {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)d * d; e += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(e % d), c = (int)(e / d);
      a.H[(size_t)(off + c) * a.ldh + off + r] += et * a.rnd[b].P[e];
    }
  }
}

// ---- host ---------------------------------------------------------------------------------------
// Builds the tile list, the (tile, chunk) work list from the occupancy map and its stream-K partition.
int syrk_plan_create(bgp_model* m) {
  SyrkPlan* pl = new SyrkPlan();
  m->syrk_plan = pl;
  if (make_tensormap_f64(&pl->tmA, m->A, (uint64_t)m->lda, (uint64_t)m->n, (uint64_t)m->lda, 16, SK_KB) != 0) {
    set_error("cuTensorMapEncodeTiled failed for the Hessian kernel");
    return BGP_ERR_CUDA;
  }
  if (m->nchunks >= (1 << 24)) {
    set_error("more than 2^30 observations per device are not supported by the Hessian work list");
    return BGP_ERR_ARG;
  }
  const int nbox = m->lda / 16;
  const int nJ = (nbox + SK_NBOX - 1) / SK_NBOX;
  // Row boxes of N panel J: r >= 4J.  Rows r >= 4J+3 meet all four boxes of the panel ("full"); they are
  // grouped four at a time in row order, so that observations sorted by zero pattern switch the rows of a
  // tile on one after the other while every consumer warp (= column box) keeps the same share.  The three
  // rows that cross the diagonal (1, 2, 3 boxes) go with the left-over full rows.
  std::vector<SkTile> tiles;
  for (int J = 0; J < nJ; ++J) {
    const int c_last = std::min(nbox, SK_NBOX * J + SK_NBOX) - 1;
    std::vector<int> full, diag;
    for (int r = SK_NBOX * J; r < nbox; ++r) (r >= c_last ? full : diag).push_back(r);
    std::vector<std::vector<int>> groups;
    size_t i = 0;
    for (; i + SK_MBOX <= full.size(); i += SK_MBOX) groups.push_back(std::vector<int>(full.begin() + i, full.begin() + i + SK_MBOX));
    std::vector<int> rest(full.begin() + i, full.end());
    if (rest.size() + diag.size() <= (size_t)SK_MBOX) {
      rest.insert(rest.end(), diag.begin(), diag.end());
      if (!rest.empty()) groups.push_back(rest);
    } else {
      if (!rest.empty()) groups.push_back(rest);
      if (!diag.empty()) groups.push_back(diag);
    }
    for (const auto& g : groups) {
      SkTile t;
      memset(&t, 0, sizeof(t));
      t.J = J;
      for (int w = 0; w < 8; ++w) t.rows[w] = -1;
      for (size_t w = 0; w < g.size(); ++w) {
        const int r = g[w];
        t.rows[w] = (int8_t)r;
        for (int b = 0; b < SK_NBOX; ++b) {
          const int c = SK_NBOX * J + b;
          if (c < nbox && r >= c) t.smask |= 1u << (4 * w + b);
        }
      }
      if (getenv("BGP_SK_FULL")) t.smask = 0xffffu;   // diagnostics only
      // box ownership: warp b takes column box b; while a warp is idle and another holds >= 2 boxes more,
      // the idle warp takes over half of the busiest warp's row slots (same column box)
      for (int b = 0; b < SK_NBOX; ++b) {
        t.wcol[b] = (uint8_t)b;
        t.wrows[b] = 0;
        for (int w = 0; w < SK_MBOX; ++w) t.wrows[b] |= (uint8_t)(((t.smask >> (4 * w + b)) & 1u) << w);
      }
      for (int round = 0; round < 4; ++round) {
        int lo = 0, hi = 0;
        for (int b = 1; b < SK_NBOX; ++b) {
          if (__builtin_popcount(t.wrows[b]) < __builtin_popcount(t.wrows[lo])) lo = b;
          if (__builtin_popcount(t.wrows[b]) > __builtin_popcount(t.wrows[hi])) hi = b;
        }
        const int nhi = __builtin_popcount(t.wrows[hi]);
        if (t.wrows[lo] != 0 || nhi < 2) break;
        uint8_t moved = 0;
        int left = nhi / 2;
        for (int w = SK_MBOX - 1; w >= 0 && left > 0; --w)
          if ((t.wrows[hi] >> w) & 1u) {
            moved |= (uint8_t)(1u << w);
            --left;
          }
        t.wrows[hi] &= (uint8_t)~moved;
        t.wrows[lo] = moved;
        t.wcol[lo] = t.wcol[hi];
      }
      tiles.push_back(t);
    }
  }
  pl->ntiles = (int)tiles.size();
  if (pl->ntiles > 255) {
    set_error("Hessian tile count %d exceeds the work-list encoding", pl->ntiles);
    return BGP_ERR_ARG;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  pl->G = SK_CTAS_PER_SM * sms;
  // Work list.  A unit is (tile, run of chunks) of bounded cost; the queue walks the observations from the
  // last block to the first (after the zero-pattern sort the late blocks are the densest) and visits every
  // tile per block, so the CTAs that run at the same time read the same observations (L2 reuse).  Unit
  // cost: total / (4 G) for the first 70 % of the work, then 1/2, 1/4, 1/8 and 1/16 of that (short queue tail).
  // cost of a chunk = fixed overhead + DMMA time of the busiest consumer warp (warp b owns column box b).
  auto chunk_cost = [&](const SkTile& t, uint64_t o, int& boxes) -> int {
    uint32_t act_m, load_m, need_n;
    sk_chunk_masks(t, o, act_m, load_m, need_n);
    boxes = 0;
    if (!act_m) return 0;
    const uint32_t ncol = (uint32_t)(o >> (4 * t.J)) & 0xfu;
    int mx = 0;
    for (int w = 0; w < SK_NBOX; ++w) {
      const int bw = ((ncol >> t.wcol[w]) & 1u) ? __builtin_popcount(t.wrows[w] & act_m) : 0;
      boxes += bw;
      mx = std::max(mx, bw);
    }
    return boxes ? 2 + 4 * mx : 0;
  };
  double boxes_total = 0.0;
  int64_t cost_total = 0;
  for (int ti = 0; ti < pl->ntiles; ++ti)
    for (int64_t c = 0; c < m->nchunks; ++c) {
      int boxes;
      cost_total += chunk_cost(tiles[ti], m->occ_host[(size_t)c], boxes);
      boxes_total += boxes;
    }
  const int64_t unit_target = std::max<int64_t>(64, cost_total / ((int64_t)4 * pl->G));
  const int64_t blk = 8;                                   // chunks per block of the walk
  std::vector<uint32_t> entries;
  std::vector<int2> units;
  std::vector<int64_t> unit_cost_fine;
  std::vector<int> unit_tile_fine;
  entries.reserve((size_t)pl->ntiles * m->nchunks / 2);
  std::vector<std::vector<uint32_t>> pend((size_t)pl->ntiles);
  std::vector<int64_t> pend_cost((size_t)pl->ntiles, 0);
  int64_t done_cost = 0;
  auto emit = [&](int ti) {
    if (pend[(size_t)ti].empty()) return;
    const int e0 = (int)entries.size();
    entries.insert(entries.end(), pend[(size_t)ti].begin(), pend[(size_t)ti].end());
    units.push_back(make_int2(e0, (int)entries.size()));
    unit_cost_fine.push_back(pend_cost[(size_t)ti]);
    done_cost += pend_cost[(size_t)ti];
    pend[(size_t)ti].clear();
    pend_cost[(size_t)ti] = 0;
  };
  const int64_t nblk = (m->nchunks + blk - 1) / blk;
  // Walk order of the observation blocks.  After the zero-pattern sort the late blocks are the densest, the early
  // ones touch a few column boxes only: their stages carry so little work that the 4-stage ring cannot cover the
  // TMA latency (full-barrier waits; dense rows run 8 % faster per executed flop).  Walking dense -> sparse left all
  // of that latency-bound work for the end of the queue, when nothing else shares the SM (measured: CTA end times
  // spread over 7 % of the kernel whatever the unit size).  Alternating blocks from both ends keeps sparse and
  // dense units resident together (3 CTAs per SM), so one CTA's waits are the others' tensor-pipe time, and ends
  // the queue in the middle of the range.  BGP_SK_WALK=0 restores dense -> sparse (diagnostics).
  std::vector<int64_t> walk;
  {
    const char* e = getenv("BGP_SK_WALK");
    const bool interleave = !(e && e[0] == '0');
    int64_t lo = 0, hi = nblk - 1;
    while (lo <= hi) {
      walk.push_back(hi--);
      if (interleave && lo <= hi) walk.push_back(lo++);
    }
  }
  for (int64_t bk : walk) {
    const int64_t c0 = bk * blk, c1 = std::min<int64_t>(m->nchunks, c0 + blk);
    for (int ti = 0; ti < pl->ntiles; ++ti) {
      for (int64_t c = c0; c < c1; ++c) {
        int boxes;
        const int cst = chunk_cost(tiles[ti], m->occ_host[(size_t)c], boxes);
        if (!cst) continue;
        pend[(size_t)ti].push_back(((uint32_t)ti << 24) | (uint32_t)c);
        pend_cost[(size_t)ti] += cst;
      }
      // the queue tail: CTAs finish within one (small) unit of each other.  With 1/4-size units to the end the CTA
      // end times spread over 7 % of the kernel (ncu / BGP_SK_DEBUG, round 1); 1/8 and 1/16 units for the last 7 %
      // of the work cost ~1 k more partial slots
      const int64_t pct = done_cost * 100 / std::max<int64_t>(1, cost_total);
      const int64_t target = pct < 70 ? unit_target
                             : (pct < 85 ? unit_target / 2 : (pct < 93 ? unit_target / 4 : (pct < 97 ? unit_target / 8 : unit_target / 16)));
      if (pend_cost[(size_t)ti] >= target) emit(ti);
    }
  }
  for (int ti = 0; ti < pl->ntiles; ++ti) emit(ti);
  pl->nunits = (int)units.size();
  pl->nslots = pl->nunits;
  pl->nentries = (int64_t)entries.size();
  pl->useful_flops = boxes_total * 64.0 * 16.0 * 16.0 * 2.0;
  // partial slot of a unit = its index; the slots of a tile are summed in queue order
  std::vector<std::vector<int>> slots_of_tile((size_t)pl->ntiles);
  for (int u = 0; u < pl->nunits; ++u) {
    const int ti = (int)(entries[(size_t)units[(size_t)u].x] >> 24);
    slots_of_tile[(size_t)ti].push_back(u);
    unit_tile_fine.push_back(ti);
  }
  std::vector<int> tile_slots;
  for (int ti = 0; ti < pl->ntiles; ++ti) {
    tiles[(size_t)ti].slot_off = (int)tile_slots.size();
    tiles[(size_t)ti].slot_cnt = (int)slots_of_tile[(size_t)ti].size();
    tile_slots.insert(tile_slots.end(), slots_of_tile[(size_t)ti].begin(), slots_of_tile[(size_t)ti].end());
  }
  auto upload = [&](auto** dst, const auto& v) -> int {
    const size_t bytes = std::max<size_t>(1, v.size()) * sizeof(v[0]);
    BGP_CUDA(cudaMalloc((void**)dst, bytes));
    if (!v.empty()) BGP_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
    return BGP_OK;
  };
  BGP_TRY(upload(&pl->tiles_dev, tiles));
  BGP_TRY(upload(&pl->entries_dev, entries));
  BGP_TRY(upload(&pl->units_dev, units));
  BGP_CUDA(cudaMalloc(&pl->counter_dev, sizeof(int)));
  BGP_TRY(upload(&pl->tile_slots_dev, tile_slots));
  m->part_H_bytes = (size_t)std::max(1, pl->nslots) * SK_TILE_ELEMS * sizeof(double);
  BGP_CUDA(cudaMalloc(&m->part_H, m->part_H_bytes));
  BGP_CUDA(cudaMemsetAsync(m->part_H, 0, m->part_H_bytes, m->stream));
  BGP_CUDA(cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM));
  m->hess_useful_flops = pl->useful_flops;
  if (getenv("BGP_SK_DEBUG")) {
    BGP_CUDA(cudaMalloc(&pl->dbg_dev, ((size_t)pl->G * 2 + (size_t)pl->nunits * 2) * sizeof(unsigned long long)));
    pl->unit_cost_dbg = unit_cost_fine;
    pl->unit_tile_dbg = unit_tile_fine;
    pl->dbg_left = 3;
    fprintf(stderr, "[syrk] tiles %d entries %lld units %d G %d\n", pl->ntiles, (long long)pl->nentries, pl->nunits, pl->G);
  }
  return BGP_OK;
}

void syrk_plan_destroy(bgp_model* m) {
  SyrkPlan* pl = (SyrkPlan*)m->syrk_plan;
  if (!pl) return;
  for (void* ptr : {(void*)pl->tiles_dev, (void*)pl->entries_dev, (void*)pl->units_dev, (void*)pl->counter_dev,
                    (void*)pl->tile_slots_dev, (void*)pl->dbg_dev})
    if (ptr) cudaFree(ptr);
  delete pl;
  m->syrk_plan = nullptr;
}

// H_lik = A^T diag(w) A (both triangles); Q is added by launch_add_q after the optional allreduce
int launch_syrk(bgp_model* m) {
  SyrkPlan* pl = (SyrkPlan*)m->syrk_plan;
  static const bool sk_rotate = getenv("BGP_SK_NOROT") == nullptr;     // env: diagnostics (fixed box -> warp map)
  BGP_CUDA(cudaMemsetAsync(pl->counter_dev, 0, sizeof(int), m->stream));
  syrk_kernel<<<pl->G, SK_THREADS, SK_SMEM, m->stream>>>(pl->tmA, m->wobs, m->part_H, pl->tiles_dev, pl->entries_dev,
                                                        pl->units_dev, pl->nunits, pl->counter_dev,
                                                        (const unsigned long long*)m->occ_dev, pl->dbg_dev,
                                                        sk_rotate ? 1 : 0);
  count_launch();
  if (pl->dbg_dev && pl->dbg_left > 0) {
    --pl->dbg_left;
    std::vector<unsigned long long> h((size_t)pl->G * 2 + (size_t)pl->nunits * 2);
    cudaStreamSynchronize(m->stream);
    cudaMemcpy(h.data(), pl->dbg_dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t1 = 0;
    for (int g = 0; g < pl->G; ++g) {
      t0 = std::min(t0, h[2 * g]);
      t1 = std::max(t1, h[2 * g + 1]);
    }
    std::vector<double> dur((size_t)pl->G), endt((size_t)pl->G);
    for (int g = 0; g < pl->G; ++g) {
      dur[g] = (double)(h[2 * g + 1] - h[2 * g]) * 1e-3;
      endt[g] = (double)(h[2 * g + 1] - t0) * 1e-3;
    }
    std::vector<double> sd = dur, se = endt;
    std::sort(sd.begin(), sd.end());
    std::sort(se.begin(), se.end());
    fprintf(stderr, "[syrk] span %.1f us; CTA duration min %.1f p10 %.1f med %.1f p90 %.1f max %.1f; end time p10 %.1f med %.1f p90 %.1f\n",
            (double)(t1 - t0) * 1e-3, sd.front(), sd[sd.size() / 10], sd[sd.size() / 2], sd[sd.size() * 9 / 10], sd.back(),
            se[se.size() / 10], se[se.size() / 2], se[se.size() * 9 / 10]);
    if (pl->dbg_left == 0) {
      // per-unit durations: next unit start on the same CTA (or CTA end) minus this unit's start
      std::vector<std::vector<std::pair<unsigned long long, int>>> per_cta((size_t)pl->G);
      for (int u = 0; u < pl->nunits; ++u)
        per_cta[(size_t)h[2 * pl->G + 2 * u + 1]].push_back({h[2 * pl->G + 2 * u], u});
      std::vector<double> tile_time((size_t)pl->ntiles, 0.0), tile_cost((size_t)pl->ntiles, 0.0);
      double ttot = 0, ctot = 0;
      for (int g = 0; g < pl->G; ++g) {
        auto& v = per_cta[(size_t)g];
        std::sort(v.begin(), v.end());
        for (size_t i = 0; i < v.size(); ++i) {
          const unsigned long long e = i + 1 < v.size() ? v[i + 1].first : h[2 * g + 1];
          const double d = (double)(e - v[i].first) * 1e-3;
          const int u = v[i].second;
          tile_time[(size_t)pl->unit_tile_dbg[(size_t)u]] += d;
          tile_cost[(size_t)pl->unit_tile_dbg[(size_t)u]] += (double)pl->unit_cost_dbg[(size_t)u];
          ttot += d;
          ctot += (double)pl->unit_cost_dbg[(size_t)u];
        }
      }
      // the end of the queue: the units that were started last, and when the CTAs that ran them finished
      {
        std::vector<std::pair<unsigned long long, int>> by_start;
        for (int u = 0; u < pl->nunits; ++u) by_start.push_back({h[2 * pl->G + 2 * u], u});
        std::sort(by_start.begin(), by_start.end());
        fprintf(stderr, "[syrk] last units (start us, cost, tile, cta, cta end us):");
        for (size_t i = by_start.size() > 24 ? by_start.size() - 24 : 0; i < by_start.size(); ++i) {
          const int u = by_start[i].second;
          const int g = (int)h[2 * pl->G + 2 * u + 1];
          fprintf(stderr, " (%.0f, %lld, %d, %d, %.0f)", (double)(by_start[i].first - t0) * 1e-3,
                  (long long)pl->unit_cost_dbg[(size_t)u], pl->unit_tile_dbg[(size_t)u], g, (double)(h[2 * g + 1] - t0) * 1e-3);
        }
        fprintf(stderr, "\n[syrk] unit cost quantiles: ");
        std::vector<int64_t> uc(pl->unit_cost_dbg.begin(), pl->unit_cost_dbg.end());
        std::sort(uc.begin(), uc.end());
        for (double q : {0.0, 0.1, 0.5, 0.9, 1.0}) fprintf(stderr, " %lld", (long long)uc[(size_t)(q * (uc.size() - 1))]);
        // the slowest CTAs: their units in order
        std::vector<std::pair<double, int>> ends;
        for (int g = 0; g < pl->G; ++g) ends.push_back({endt[g], g});
        std::sort(ends.begin(), ends.end());
        for (int k = 0; k < 3; ++k) {
          const int g = ends[ends.size() - 1 - k].second;
          fprintf(stderr, "\n[syrk] slow CTA %d (end %.0f us) units (start us, cost, tile):", g, endt[g]);
          for (auto& pr : per_cta[(size_t)g])
            fprintf(stderr, " (%.0f, %lld, %d)", (double)(pr.first - t0) * 1e-3, (long long)pl->unit_cost_dbg[(size_t)pr.second],
                    pl->unit_tile_dbg[(size_t)pr.second]);
        }
        fprintf(stderr, "\n");
      }
      fprintf(stderr, "[syrk] per tile: time share / cost share (ratio):");
      for (int t = 0; t < pl->ntiles; ++t)
        fprintf(stderr, " %d:%.3f/%.3f(%.2f)", t, tile_time[t] / ttot, tile_cost[t] / ctot,
                (tile_time[t] / ttot) / std::max(1e-12, tile_cost[t] / ctot));
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "[syrk] per-CTA duration by index (every 37th):");
    for (int g = 0; g < pl->G; g += 37) fprintf(stderr, " %d:%.0f", g, dur[g]);
    fprintf(stderr, "\n");
  }
  syrk_reduce_kernel<<<pl->ntiles * (SK_TILE_ELEMS / SR_ELEMS), SR_ELEMS * SR_LANES, 0, m->stream>>>(m->part_H, pl->tiles_dev,
                                                                               pl->tile_slots_dev, m->p, m->ldh, m->H);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

int launch_add_q(bgp_model* m, const double* theta) {
  AddQArgs a;
  a.H = m->H;
  a.ldh = m->ldh;
  a.p = m->p;
  a.qfix = m->qfix;
  a.nrnd = m->J;
  for (int j = 0; j < m->J; ++j) {
    a.rnd[j].off = m->rnd[j].off;
    a.rnd[j].d = m->rnd[j].d;
    a.rnd[j].diag = m->rnd[j].diag ? 1 : 0;
    a.rnd[j].P = m->rnd[j].P_dev;
    a.rnd[j].etheta = std::exp(theta[j]);
  }
  dim3 grid(8, m->J + 1);
  add_q_kernel<<<grid, 256, 0, m->stream>>>(a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

// lower triangle of H, column by column: element (r, c), r >= c, at c p - c (c - 1) / 2 + (r - c)
__global__ void pack_lower_kernel(const double* __restrict__ H, int p, int ldh, double* __restrict__ out) {
  const int c = blockIdx.x;
  double* dst = out + ((size_t)c * p - (size_t)c * (c - 1) / 2);
  for (int r = c + threadIdx.x; r < p; r += blockDim.x) dst[r - c] = H[(size_t)c * ldh + r];
}
__global__ void unpack_lower_kernel(const double* __restrict__ in, int p, int ldh, double* __restrict__ H) {
  const int c = blockIdx.x;
  const double* src = in + ((size_t)c * p - (size_t)c * (c - 1) / 2);
  for (int r = c + threadIdx.x; r < p; r += blockDim.x) {
    const double v = src[r - c];
    H[(size_t)c * ldh + r] = v;
    H[(size_t)r * ldh + c] = v;
  }
}

int launch_hessian(bgp_model* m, const double* theta) {
  m->L_holds_H = false;
  if (m->osp_on) {
    // moment path: H_lik and, on a single device, Q(theta) in one kernel
    BGP_TRY(osp_launch_hessian(m, m->world > 1 ? nullptr : theta));
    if (m->world == 1) return BGP_OK;
  } else {
    BGP_TRY(launch_syrk(m));
  }
  if (m->world > 1) {
    // observation shards: the likelihood Hessians of the ranks are summed over NVLink; only the packed lower
    // triangle (p (p + 1) / 2 doubles) travels, both triangles are rewritten from the sum so that every rank holds
    // the same symmetric matrix bit for bit
    const size_t np = (size_t)m->p * (m->p + 1) / 2;
    if (!m->hpack) BGP_CUDA(cudaMalloc(&m->hpack, np * sizeof(double)));
    pack_lower_kernel<<<m->p, 128, 0, m->stream>>>(m->H, m->p, m->ldh, m->hpack);
    count_launch();
    BGP_TRY(comm_allreduce_sum(m, m->hpack, np));
    unpack_lower_kernel<<<m->p, 128, 0, m->stream>>>(m->hpack, m->p, m->ldh, m->H);
    count_launch();
    BGP_CUDA(cudaGetLastError());
  }
  return launch_add_q(m, theta);
}

}  // namespace bgp
