// syrk.cu — H = A^T diag(w) A + Q(theta): the dense FP64 contraction of every Newton iteration.
//
// Replaces the sparse-Hessian AD sweep TMB runs on objective_function::operator()
// (/root/reference/src/BayesGP.cpp:30-253; ff$env$spHess(random = TRUE), call site
// /root/reference/R/02_model_fit.R:276-284).  n*p*(p+1) flops per evaluation.
//
// Design (sm_100a):
//   * FP64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — tcgen05 has no FP64 kind).
//   * The lower triangle of H is cut into 64x64 tiles; each CTA owns one (tile, observation
//     split) and accumulates 64x64 in registers (4 warps x 32x32, 64 accumulator doubles/thread).
//   * Operands are TMA-staged: A is observation-major, so a TMA box of {16 columns, 16 rows}
//     lands as 16 lines of 128 B with the hardware 128B swizzle; a 64-column panel is 4 boxes.
//     3-stage mbarrier pipeline, one elected thread issues the copies.  Fragment loads are
//     LDS.128 with a column permutation chosen so the swizzled lines are read conflict-free;
//     the permutation is undone in the epilogue.
//   * diag(w) is applied to the A fragment in registers (one DMUL per fragment element).
//   * Split-K partials go to a workspace and are reduced in a fixed order (deterministic),
//     mirrored to the upper triangle, then Q(theta) is added.
//   * CTAs of the same observation split are adjacent in the grid, so the panels they share
//     are served by L2: HBM traffic stays ~8*n*lda bytes per Hessian.
#include <cuda.h>

#include <algorithm>

#include "bgp_internal.h"
#include "ptx.cuh"

namespace bgp {

constexpr int SY_T = 64;          // tile edge
constexpr int SY_KB = 16;         // observations per pipeline stage
constexpr int SY_STAGES = 3;
constexpr int SY_THREADS = 128;
constexpr int SY_BOX_BYTES = 16 * SY_KB * 8;            // 2048
constexpr int SY_PANEL_BYTES = 4 * SY_BOX_BYTES;        // 8192
constexpr int SY_STAGE_BYTES = 2 * SY_PANEL_BYTES;      // 16384
constexpr int SY_W_BYTES = SY_KB * 8;                   // 128
constexpr int SY_SMEM = SY_STAGES * SY_STAGE_BYTES + SY_STAGES * SY_W_BYTES + 16 * SY_STAGES + 16 + 1024;

struct SyrkPlan {
  CUtensorMap tmA;
  int nt = 0, ntiles = 0, nsplit = 0;
  int64_t chunk = 0;
  int2* tiles_dev = nullptr;
};

using namespace ptx;

// column permutation inside a 16-column box: fragment row j reads 16-byte chunk ch(j)
__host__ __device__ __forceinline__ int sy_chunk(int j) { return (j >> 1) + 4 * (j & 1); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Work is skipped at the granularity of 16 x 16 boxes (the fragment permutation interleaves the two
// 8-row fragments of a box): a box is computed iff it touches the lower triangle (incl. diagonal)
// and lies inside the lda x lda matrix.
__global__ void __launch_bounds__(SY_THREADS, 4)
    syrk_kernel(const __grid_constant__ CUtensorMap tmA, const double* __restrict__ wobs, double* __restrict__ part,
                const int2* __restrict__ tiles, int ntiles, int64_t n, int64_t chunk, int lda) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base + SY_STAGES * SY_STAGE_BYTES;
  const uint32_t full_base = w_base + SY_STAGES * SY_W_BYTES;
  const uint32_t empty_base = full_base + 8 * SY_STAGES;

  const int tile = blockIdx.x % ntiles, split = blockIdx.x / ntiles;
  const int ti = tiles[tile].x, tj = tiles[tile].y;
  const bool diag = ti == tj;
  const int64_t k_begin = (int64_t)split * chunk;
  const int64_t k_end = k_begin + chunk < n ? k_begin + chunk : n;
  const int niter = k_end > k_begin ? (int)((k_end - k_begin + SY_KB - 1) / SY_KB) : 0;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  // 2 x 2 boxes per warp: bit (bi * 2 + bj)
  int bmask = 0;
#pragma unroll
  for (int bi = 0; bi < 2; ++bi)
#pragma unroll
    for (int bj = 0; bj < 2; ++bj) {
      const int r0 = ti * SY_T + (wm * 2 + bi) * 16, c0 = tj * SY_T + (wn * 2 + bj) * 16;
      if (r0 < lda && c0 < lda && r0 + 15 >= c0) bmask |= 1 << (bi * 2 + bj);
    }

  if (tid == 0) {
    for (int s = 0; s < SY_STAGES; ++s) {
      mbar_init(full_base + 8 * s, 1);
      mbar_init(empty_base + 8 * s, SY_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t stage_tx = (diag ? SY_PANEL_BYTES : 2 * SY_PANEL_BYTES) + SY_W_BYTES;
  auto issue = [&](int it) {
    const int s = it % SY_STAGES;
    const uint32_t bar = full_base + 8 * s;
    const uint32_t sa = base + s * SY_STAGE_BYTES;
    const int row = (int)(k_begin + (int64_t)it * SY_KB);
    mbar_expect_tx(bar, stage_tx);
#pragma unroll
    for (int b = 0; b < 4; ++b) tma_load_2d(sa + b * SY_BOX_BYTES, &tmA, ti * SY_T + b * 16, row, bar);
    if (!diag) {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        tma_load_2d(sa + SY_PANEL_BYTES + b * SY_BOX_BYTES, &tmA, tj * SY_T + b * 16, row, bar);
    }
    bulk_load_1d(w_base + s * SY_W_BYTES, wobs + row, SY_W_BYTES, bar);
  };
  if (tid == 0 && niter > 0) issue(0);

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int fj = lane >> 2, fk = lane & 3;
  const int ch = sy_chunk(fj);

  // mbarrier ring, no CTA-wide barrier in the loop: the elected producer thread prefetches one stage
  // ahead into the slot whose readers (all 4 warps, iteration it - 2) released it via empty[slot].
  for (int it = 0; it < niter; ++it) {
    if (tid == 0 && it + 1 < niter) {
      const int j = it + 1;
      if (j >= SY_STAGES) mbar_wait(empty_base + 8 * (j % SY_STAGES), (uint32_t)(((j / SY_STAGES) - 1) & 1));
      issue(j);
    }
    const int s = it % SY_STAGES;
    mbar_wait(full_base + 8 * s, (uint32_t)((it / SY_STAGES) & 1));
    if (bmask) {
      const uint32_t sa = base + s * SY_STAGE_BYTES;
      const uint32_t pa = sa + (wm * 2) * SY_BOX_BYTES;
      const uint32_t pb = (diag ? sa : sa + SY_PANEL_BYTES) + (wn * 2) * SY_BOX_BYTES;
      const uint32_t pw = w_base + s * SY_W_BYTES;
#pragma unroll
      for (int kk = 0; kk < SY_KB / 4; ++kk) {
        const int row = kk * 4 + fk;
        const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
        const double2 a0 = lds128(pa + off), a1 = lds128(pa + SY_BOX_BYTES + off);
        const double2 b0 = lds128(pb + off), b1 = lds128(pb + SY_BOX_BYTES + off);
        const double wk = lds64(pw + row * 8);
        const double af[4] = {a0.x * wk, a0.y * wk, a1.x * wk, a1.y * wk};
        const double bf[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
        for (int bi = 0; bi < 2; ++bi)
#pragma unroll
          for (int bj = 0; bj < 2; ++bj)
            if (bmask & (1 << (bi * 2 + bj))) {
#pragma unroll
              for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int f = 0; f < 2; ++f)
                  dmma884(acc[2 * bi + e][2 * bj + f][0], acc[2 * bi + e][2 * bj + f][1], af[2 * bi + e], bf[2 * bj + f]);
            }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_base + 8 * s);
  }

  // ---- epilogue: undo the column permutation, write the 64x64 partial (row-major [M][N]) --------
  if (bmask) {
    double* out = part + ((size_t)split * ntiles + tile) * (SY_T * SY_T);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int M = (wm * 2 + (mi >> 1)) * 16 + 2 * ch + (mi & 1);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        if (!(bmask & (1 << ((mi >> 1) * 2 + (ni >> 1))))) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jn = 2 * fk + e;
          const int N = (wn * 2 + (ni >> 1)) * 16 + 2 * sy_chunk(jn) + (ni & 1);
          out[M * SY_T + N] = acc[mi][ni][e];
        }
      }
    }
  }
}

// sum the split-K partials in a fixed order, write the lower triangle and its mirror
__global__ void __launch_bounds__(256)
    syrk_reduce_kernel(const double* __restrict__ part, const int2* __restrict__ tiles, int ntiles, int nsplit, int p,
                       int ldh, double* __restrict__ H) {
  const int tile = blockIdx.x / (SY_T * SY_T / 256);
  const int e = (blockIdx.x % (SY_T * SY_T / 256)) * 256 + threadIdx.x;
  const int M = e / SY_T, N = e % SY_T;
  const int ti = tiles[tile].x, tj = tiles[tile].y;
  const int gr = ti * SY_T + M, gc = tj * SY_T + N;
  if (gr >= p || gc >= p || gc > gr) return;
  double s = 0.0;
  for (int sp = 0; sp < nsplit; ++sp) s += part[((size_t)sp * ntiles + tile) * (SY_T * SY_T) + e];
  H[(size_t)gc * ldh + gr] = s;
  H[(size_t)gr * ldh + gc] = s;
}

struct AddQArgs {
  double* H;
  int ldh, p;
  int fix_start;       // first W index that is not a spline coefficient
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
    for (int c = a.fix_start + blockIdx.x * blockDim.x + threadIdx.x; c < a.p; c += gridDim.x * blockDim.x)
      a.H[(size_t)c * a.ldh + c] += a.qfix[c];
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
int syrk_plan_create(bgp_model* m) {
  SyrkPlan* pl = new SyrkPlan();
  m->syrk_plan = pl;
  if (make_tensormap_f64(&pl->tmA, m->A, (uint64_t)m->lda, (uint64_t)m->n, (uint64_t)m->lda, 16, SY_KB) != 0) {
    set_error("cuTensorMapEncodeTiled failed for the Hessian kernel");
    return BGP_ERR_CUDA;
  }
  pl->nt = (m->p + SY_T - 1) / SY_T;
  pl->ntiles = pl->nt * (pl->nt + 1) / 2;
  std::vector<int2> tiles;
  for (int i = 0; i < pl->nt; ++i)
    for (int j = 0; j <= i; ++j) tiles.push_back(make_int2(i, j));
  BGP_CUDA(cudaMalloc(&pl->tiles_dev, tiles.size() * sizeof(int2)));
  BGP_CUDA(cudaMemcpy(pl->tiles_dev, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice));
  // split the observations so the grid fills 148 SMs x 4 resident CTAs about twice over
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  const int64_t target = (int64_t)sms * 4 * 2;
  int64_t nsplit = std::max<int64_t>(1, target / pl->ntiles);
  const int64_t max_split = std::max<int64_t>(1, m->n / (8 * SY_KB));
  nsplit = std::min(nsplit, max_split);
  int64_t chunk = (m->n + nsplit - 1) / nsplit;
  chunk = round_up64(chunk, SY_KB);
  nsplit = (m->n + chunk - 1) / chunk;
  if (nsplit < 1) nsplit = 1;
  pl->nsplit = (int)nsplit;
  pl->chunk = chunk;
  m->part_H_bytes = (size_t)pl->nsplit * pl->ntiles * SY_T * SY_T * sizeof(double);
  BGP_CUDA(cudaMalloc(&m->part_H, m->part_H_bytes));
  BGP_CUDA(cudaMemset(m->part_H, 0, m->part_H_bytes));
  BGP_CUDA(cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM));
  return BGP_OK;
}

void syrk_plan_destroy(bgp_model* m) {
  SyrkPlan* pl = (SyrkPlan*)m->syrk_plan;
  if (!pl) return;
  if (pl->tiles_dev) cudaFree(pl->tiles_dev);
  delete pl;
  m->syrk_plan = nullptr;
}

// H_lik = A^T diag(w) A (both triangles); Q is added by launch_add_q after the optional allreduce
int launch_syrk(bgp_model* m) {
  SyrkPlan* pl = (SyrkPlan*)m->syrk_plan;
  syrk_kernel<<<pl->ntiles * pl->nsplit, SY_THREADS, SY_SMEM, m->stream>>>(pl->tmA, m->wobs, m->part_H, pl->tiles_dev,
                                                                            pl->ntiles, m->n, pl->chunk, m->lda);
  count_launch();
  syrk_reduce_kernel<<<pl->ntiles * (SY_T * SY_T / 256), 256, 0, m->stream>>>(m->part_H, pl->tiles_dev, pl->ntiles,
                                                                               pl->nsplit, m->p, m->ldh, m->H);
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
  a.fix_start = 0;
  for (int j = 0; j < m->J; ++j) {
    a.rnd[j].off = m->rnd[j].off;
    a.rnd[j].d = m->rnd[j].d;
    a.rnd[j].diag = m->rnd[j].diag ? 1 : 0;
    a.rnd[j].P = m->rnd[j].P_dev;
    a.rnd[j].etheta = std::exp(theta[j]);
    a.fix_start += m->rnd[j].d;
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
