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
//   * Persistent, warp-specialised CTAs (2 per SM): one producer warp + 8 consumer warps.  A CTA tile is
//     128 rows x 64 columns of H; consumer warp w owns a 16 x 64 strip (one 16-column box of the M panel
//     against the four boxes of the N panel), 32 accumulator doubles per thread.  The diag(w) scaling is
//     applied to the strip's two A fragments only (2 DMUL per 16 DMMA — a DMUL costs ~4.7 clk of the same pipe).
//   * Operands are TMA-staged: A is observation-major, so a TMA box of {16 columns, 16 observations} lands as
//     16 lines of 128 B with the hardware 128B swizzle; a stage is 8 + 4 boxes + 16 weights.  The producer
//     warp runs a 4-stage mbarrier ring ahead of the consumers (full / empty barriers, no CTA-wide barrier).
//     Fragment loads are LDS.128 with a column permutation chosen so the swizzled lines are read
//     conflict-free; the permutation is undone in the epilogue.
//   * Work list: the unit of work is (tile, 64-observation chunk).  Cells of the {chunk x column-box}
//     occupancy map (rowsort.cu) that are structurally zero are skipped — whole chunks by leaving them out of
//     the list, single boxes by a per-stage 4-bit mask that selects a compile-time-specialised loop body
//     (a predicated-off DMMA still occupies the pipe, so skipping has to be a branch).  Boxes above the
//     diagonal or outside the lda x lda matrix are masked the same way.
//   * Stream-K scheduling: the host cuts the cost-weighted work list into one contiguous slice per CTA;
//     a CTA flushes its accumulators to a private partial slot whenever the tile changes.  Partials are
//     summed in a fixed order (deterministic), mirrored to the upper triangle, then Q(theta) is added.
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
constexpr int SK_STAGES = SK_MBOX == 4 ? 3 : 4;
constexpr int SK_CTAS_PER_SM = SK_MBOX == 4 ? 4 : 2;
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
  uint32_t smask;           // bit 4*w + b: box (rows[w], 4J + b) is on or below the diagonal and inside the matrix
  int8_t rows[8];           // box row of strip w, -1 = unused
  int pad[2];
};
static_assert(sizeof(SkTile) == 32, "SkTile layout");

struct SyrkPlan {
  CUtensorMap tmA;
  int ntiles = 0, G = 0, nslots = 0;
  SkTile* tiles_dev = nullptr;
  uint32_t* entries_dev = nullptr; // tile << 24 | chunk
  int* cta_begin_dev = nullptr;    // G + 1
  int* cta_slot_dev = nullptr;     // first partial slot of each CTA
  int* tile_slots_dev = nullptr;   // CSR of partial slots per tile
  double useful_flops = 0.0;       // structurally non-zero flops per launch
  int64_t nentries = 0;
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
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
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
                const SkTile* __restrict__ tiles, const uint32_t* __restrict__ entries, const int* __restrict__ cta_begin,
                const int* __restrict__ cta_slot, const unsigned long long* __restrict__ occ) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base + SK_STAGES * SK_STAGE_BYTES;
  const uint32_t meta_base = w_base + SK_STAGES * SK_W_BYTES;      // 16 B per stage: {masks, tile}
  const uint32_t full_base = meta_base + 16 * SK_STAGES;
  const uint32_t empty_base = full_base + 8 * SK_STAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e_begin = cta_begin[blockIdx.x], e_end = cta_begin[blockIdx.x + 1];
  const int nstage = (e_end - e_begin) * SK_SPC;

  if (tid == 0) {
    for (int s = 0; s < SK_STAGES; ++s) {
      mbar_init(full_base + 8 * s, 1);
      mbar_init(empty_base + 8 * s, SK_CONSUMERS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (nstage == 0) return;

  if (warp == SK_CONSUMERS) {
    // ---------------- producer warp: one elected lane issues the TMA boxes and the weights ---------
    if (lane != 0) return;
    int it = 0;
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
        sts64(meta_base + 16 * slot, act_m | (need_n << 8), (uint32_t)tile);   // rows present | N boxes present
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
    return;
  }

  // ---------------- consumer warps: warp w owns column box w of the N panel ----------------------------
  const int wc = warp;
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
  int cur_tile = -1, nflush = 0;
  uint32_t smask_w = 0;       // bit s: box (rows[s], 4J + wc) belongs to the lower triangle
  uint32_t a_off[4] = {0, 0, 0, 0};
  auto flush = [&]() {
    // undo the column permutation, write this column strip of the partial tile (row-major [16 * s + m][n])
    double* out = part + (size_t)(cta_slot[blockIdx.x] + nflush) * SK_TILE_ELEMS;
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
    ++nflush;
  };
  for (int it = 0; it < nstage; ++it) {
    const int slot = it % SK_STAGES;
    mbar_wait(full_base + 8 * slot, (uint32_t)((it / SK_STAGES) & 1));
    const uint2 meta = lds_u2(meta_base + 16 * slot);
    const int tile = (int)meta.y;
    if (tile != cur_tile) {
      if (cur_tile >= 0) {
        flush();
        zero_acc();
      }
      cur_tile = tile;
      const SkTile* tp = tiles + tile;
      const uint32_t sm = tp->smask;
      const int j4 = 4 * tp->J;
      smask_w = 0;
#pragma unroll
      for (int sr = 0; sr < 4; ++sr) {
        smask_w |= ((sm >> (4 * sr + wc)) & 1u) << sr;
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
  flush();
}

// sum the partial slots of each tile in a fixed order, write the lower triangle and its mirror
__global__ void __launch_bounds__(256)
    syrk_reduce_kernel(const double* __restrict__ part, const SkTile* __restrict__ tiles, const int* __restrict__ tile_slots,
                       int p, int ldh, double* __restrict__ H) {
  const int tile = blockIdx.x / (SK_TILE_ELEMS / 256);
  const int e = (blockIdx.x % (SK_TILE_ELEMS / 256)) * 256 + threadIdx.x;
  const int M = e / (16 * SK_NBOX), N = e % (16 * SK_NBOX);
  const SkTile t = tiles[tile];
  const int r = t.rows[M / 16];
  if (r < 0) return;
  const int gr = r * 16 + (M % 16), gc = t.J * 16 * SK_NBOX + N;
  if (gr >= p || gc >= p || gc > gr) return;
  double s = 0.0;
  for (int k = 0; k < t.slot_cnt; ++k) s += part[(size_t)tile_slots[t.slot_off + k] * SK_TILE_ELEMS + e];
  H[(size_t)gc * ldh + gr] = s;
  H[(size_t)gr * ldh + gc] = s;
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
  // work list: tile-major, chunks ascending; cost = fixed per-chunk overhead + DMMA work
  int cost_mode = 1;
  if (const char* e = getenv("BGP_SK_COST")) cost_mode = atoi(e);
  std::vector<uint32_t> entries;
  std::vector<uint8_t> cost;
  entries.reserve((size_t)pl->ntiles * m->nchunks / 2);
  cost.reserve((size_t)pl->ntiles * m->nchunks / 2);
  double boxes_total = 0.0;
  int64_t cost_total = 0;
  for (int ti = 0; ti < pl->ntiles; ++ti) {
    const SkTile& t = tiles[ti];
    for (int64_t c = 0; c < m->nchunks; ++c) {
      const uint64_t o = m->occ_host[(size_t)c];
      uint32_t act_m, load_m, need_n;
      sk_chunk_masks(t, o, act_m, load_m, need_n);
      if (!act_m) continue;
      const uint32_t ncol = (uint32_t)(o >> (4 * t.J)) & 0xfu;
      // consumer warp b owns column box b: a chunk takes as long as the busiest warp
      int boxes = 0, mx = 0;
      for (int b = 0; b < SK_NBOX; ++b) {
        int bw = 0;
        if ((ncol >> b) & 1u)
          for (int w = 0; w < SK_MBOX; ++w) bw += (int)((act_m >> w) & 1u) & (int)((t.smask >> (4 * w + b)) & 1u);
        boxes += bw;
        mx = std::max(mx, bw);
      }
      const int cst = 2 + (cost_mode == 0 ? boxes : (cost_mode == 1 ? 4 * mx : 2 * mx + boxes / 2));
      entries.push_back(((uint32_t)ti << 24) | (uint32_t)c);
      cost.push_back((uint8_t)cst);
      boxes_total += boxes;
      cost_total += cst;
    }
  }
  pl->nentries = (int64_t)entries.size();
  pl->useful_flops = boxes_total * 64.0 * 16.0 * 16.0 * 2.0;
  // stream-K: contiguous slices of equal cost
  std::vector<int> cta_begin((size_t)pl->G + 1, 0), cta_slot((size_t)pl->G, 0);
  std::vector<std::vector<int>> slots_of_tile((size_t)pl->ntiles);
  {
    int64_t acc = 0;
    size_t e = 0;
    int nslots = 0;
    for (int g = 0; g < pl->G; ++g) {
      cta_begin[(size_t)g] = (int)e;
      cta_slot[(size_t)g] = nslots;
      const int64_t target = (cost_total * (int64_t)(g + 1)) / pl->G;
      int last_tile = -1;
      while (e < entries.size() && (acc < target || g == pl->G - 1)) {
        const int ti = (int)(entries[e] >> 24);
        if (ti != last_tile) {
          slots_of_tile[(size_t)ti].push_back(nslots++);
          last_tile = ti;
        }
        acc += cost[e];
        ++e;
      }
    }
    cta_begin[(size_t)pl->G] = (int)e;
    pl->nslots = nslots;
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
  BGP_TRY(upload(&pl->cta_begin_dev, cta_begin));
  BGP_TRY(upload(&pl->cta_slot_dev, cta_slot));
  BGP_TRY(upload(&pl->tile_slots_dev, tile_slots));
  m->part_H_bytes = (size_t)std::max(1, pl->nslots) * SK_TILE_ELEMS * sizeof(double);
  BGP_CUDA(cudaMalloc(&m->part_H, m->part_H_bytes));
  BGP_CUDA(cudaMemset(m->part_H, 0, m->part_H_bytes));
  BGP_CUDA(cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM));
  m->hess_useful_flops = pl->useful_flops;
  return BGP_OK;
}

void syrk_plan_destroy(bgp_model* m) {
  SyrkPlan* pl = (SyrkPlan*)m->syrk_plan;
  if (!pl) return;
  for (void* ptr : {(void*)pl->tiles_dev, (void*)pl->entries_dev, (void*)pl->cta_begin_dev, (void*)pl->cta_slot_dev,
                    (void*)pl->tile_slots_dev})
    if (ptr) cudaFree(ptr);
  delete pl;
  m->syrk_plan = nullptr;
}

// H_lik = A^T diag(w) A (both triangles); Q is added by launch_add_q after the optional allreduce
int launch_syrk(bgp_model* m) {
  SyrkPlan* pl = (SyrkPlan*)m->syrk_plan;
  syrk_kernel<<<pl->G, SK_THREADS, SK_SMEM, m->stream>>>(pl->tmA, m->wobs, m->part_H, pl->tiles_dev, pl->entries_dev,
                                                        pl->cta_begin_dev, pl->cta_slot_dev,
                                                        (const unsigned long long*)m->occ_dev);
  count_launch();
  syrk_reduce_kernel<<<pl->ntiles * (SK_TILE_ELEMS / 256), 256, 0, m->stream>>>(m->part_H, pl->tiles_dev,
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

int launch_hessian(bgp_model* m, const double* theta) {
  BGP_TRY(launch_syrk(m));
  if (m->world > 1) BGP_TRY(comm_allreduce_sum(m, m->H, (size_t)m->ldh * m->p));
  return launch_add_q(m, theta);
}

}  // namespace bgp
