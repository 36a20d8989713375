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

namespace bgp {

constexpr int SY_T = 64;          // tile edge
constexpr int SY_KB = 16;         // observations per pipeline stage
constexpr int SY_STAGES = 3;
constexpr int SY_THREADS = 128;
constexpr int SY_BOX_BYTES = 16 * SY_KB * 8;            // 2048
constexpr int SY_PANEL_BYTES = 4 * SY_BOX_BYTES;        // 8192
constexpr int SY_STAGE_BYTES = 2 * SY_PANEL_BYTES;      // 16384
constexpr int SY_W_BYTES = SY_KB * 8;                   // 128
constexpr int SY_SMEM = SY_STAGES * SY_STAGE_BYTES + SY_STAGES * SY_W_BYTES + 64 + 1024;

struct SyrkPlan {
  CUtensorMap tmA;
  int nt = 0, ntiles = 0, nsplit = 0;
  int64_t chunk = 0;
  int2* tiles_dev = nullptr;
};

// ---- PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// column permutation inside a 16-column box: fragment row j reads 16-byte chunk ch(j)
__host__ __device__ __forceinline__ int sy_chunk(int j) { return (j >> 1) + 4 * (j & 1); }

__global__ void __launch_bounds__(SY_THREADS, 4)
    syrk_kernel(const __grid_constant__ CUtensorMap tmA, const double* __restrict__ wobs, double* __restrict__ part,
                const int2* __restrict__ tiles, int ntiles, int64_t n, int64_t chunk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base + SY_STAGES * SY_STAGE_BYTES;
  const uint32_t bar_base = w_base + SY_STAGES * SY_W_BYTES;

  const int tile = blockIdx.x % ntiles, split = blockIdx.x / ntiles;
  const int ti = tiles[tile].x, tj = tiles[tile].y;
  const bool diag = ti == tj;
  const int64_t k_begin = (int64_t)split * chunk;
  const int64_t k_end = k_begin + chunk < n ? k_begin + chunk : n;
  const int niter = k_end > k_begin ? (int)((k_end - k_begin + SY_KB - 1) / SY_KB) : 0;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const bool active = !(diag && wm == 0 && wn == 1);   // strictly-upper quadrant of a diagonal tile

  if (tid == 0) {
    for (int s = 0; s < SY_STAGES; ++s) mbar_init(bar_base + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t stage_tx = (diag ? SY_PANEL_BYTES : 2 * SY_PANEL_BYTES) + SY_W_BYTES;
  auto issue = [&](int it) {
    const int s = it % SY_STAGES;
    const uint32_t bar = bar_base + 8 * s;
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

  if (tid == 0) {
    for (int it = 0; it < SY_STAGES - 1 && it < niter; ++it) issue(it);
  }

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int fj = lane >> 2, fk = lane & 3;
  const int ch = sy_chunk(fj);

  for (int it = 0; it < niter; ++it) {
    __syncthreads();   // every warp is done with stage (it-1) % STAGES
    if (tid == 0 && it + SY_STAGES - 1 < niter) issue(it + SY_STAGES - 1);
    const int s = it % SY_STAGES;
    mbar_wait(bar_base + 8 * s, (uint32_t)((it / SY_STAGES) & 1));
    if (active) {
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
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
      }
    }
  }

  // ---- epilogue: undo the column permutation, write the 64x64 partial (row-major [M][N]) --------
  if (active) {
    double* out = part + ((size_t)split * ntiles + tile) * (SY_T * SY_T);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int M = (wm * 2 + (mi >> 1)) * 16 + 2 * ch + (mi & 1);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
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
  if (ti == tj && M < 32 && N >= 32) return;
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int syrk_plan_create(bgp_model* m) {
  SyrkPlan* pl = new SyrkPlan();
  m->syrk_plan = pl;
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return BGP_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)m->lda, (cuuint64_t)m->n};
  const cuuint64_t gstr[1] = {(cuuint64_t)m->lda * 8};
  const cuuint32_t box[2] = {16, (cuuint32_t)SY_KB};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)m->A, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
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
                                                                            pl->ntiles, m->n, pl->chunk);
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
