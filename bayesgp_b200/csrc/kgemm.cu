// kgemm.cu — FP64 tensor-pipe GEMM  C[m][n] = bias[m] + sum_k A[m][k] B[n][k]  (both operands K-major).
//
// This is the DGEMM-shaped kernel behind
//   * posterior sampling     W = mode + chol(H)^-1 Z           (aghq::sample_marginal, call site
//                                                               /root/reference/R/02_model_fit.R:687-689)
//   * sample -> function     F = [X | B](x_new) [global; coef]   (compute_post_fun_IWP / _sGP,
//                                                               /root/reference/R/03_post_fit.R:235, :272)
// 128 x 64 CTA tile, 8 warps x (32 x 32) DMMA.8x8x4 accumulators, operands staged by TMA as
// {16 k, rows} boxes with the 128-byte swizzle (3-stage mbarrier ring), conflict-free LDS.64 fragment
// loads through a row permutation, C staged through shared memory for coalesced stores.
#include "bgp_internal.h"
#include "ptx.cuh"

namespace bgp {

using namespace ptx;

constexpr int KG_TM = 128, KG_TN = 64, KG_KB = 16, KG_STAGES = 3, KG_THREADS = 256;
constexpr int KG_A_BYTES = KG_TM * 128, KG_B_BYTES = KG_TN * 128, KG_STAGE_BYTES = KG_A_BYTES + KG_B_BYTES;
constexpr int KG_CPITCH = KG_TN + 1;
constexpr int KG_SMEM_PIPE = KG_STAGES * KG_STAGE_BYTES;                 // 73728
constexpr int KG_SMEM_C = KG_TM * KG_CPITCH * 8;                         // 66560
constexpr int KG_SMEM = (KG_SMEM_PIPE > KG_SMEM_C ? KG_SMEM_PIPE : KG_SMEM_C) + 64 + 1024;

struct KgemmArgs {
  int64_t M, N;
  int K;
  const double* bias;
  double* out;
  int64_t ldc;
  int out_col_major;
  const int32_t* col_perm;
  const unsigned long long* a_occ;   // per 128-row tile of A: bit s set iff k-slice s (16 wide) of the tile is not all zero
};

__global__ void __launch_bounds__(KG_THREADS, 2)
    kgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KgemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int pipe_c = KG_SMEM_PIPE > KG_SMEM_C ? KG_SMEM_PIPE : KG_SMEM_C;
  const uint32_t bar_base = base + pipe_c;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int fj = lane >> 2, fk = lane & 3;
  const int64_t m0 = (int64_t)blockIdx.x * KG_TM, n0 = (int64_t)blockIdx.y * KG_TN;
  const int nslices = (a.K + KG_KB - 1) / KG_KB;
  // k-slices of this row tile that carry anything: structural zeros of A (a prediction design B(x_new) is ~50 %
  // zeros in prefix / band form, R/01_utility.R:346-364) are neither copied nor multiplied
  __shared__ unsigned char s_list[64];
  __shared__ int s_niter;
  if (tid == 0) {
    for (int s = 0; s < KG_STAGES; ++s) mbar_init(bar_base + 8 * s, 1);
    mbar_fence_init();
    int cnt = 0;
    if (a.a_occ) {
      const unsigned long long o = a.a_occ[blockIdx.x];
      for (int sl = 0; sl < nslices; ++sl)
        if ((o >> sl) & 1ull) s_list[cnt++] = (unsigned char)sl;
    } else {
      cnt = nslices;
    }
    s_niter = cnt;
  }
  __syncthreads();
  const int niter = s_niter;
  const bool listed = a.a_occ != nullptr;
  auto issue = [&](int it) {
    const int s = it % KG_STAGES;
    const uint32_t bar = bar_base + 8 * s, sa = base + s * KG_STAGE_BYTES;
    const int sl = listed ? (int)s_list[it] : it;
    mbar_expect_tx(bar, KG_STAGE_BYTES);
    tma_load_2d(sa, &tmA, sl * KG_KB, (int)m0, bar);
    tma_load_2d(sa + KG_A_BYTES, &tmB, sl * KG_KB, (int)n0, bar);
  };
  if (tid == 0)
    for (int it = 0; it < KG_STAGES - 1 && it < niter; ++it) issue(it);

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  uint32_t a_off[4], b_off[4];
  int a_sw[4], b_sw[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int ra = wm * 32 + 16 * (f >> 1) + 2 * fj + (f & 1);
    const int rb = wn * 32 + 16 * (f >> 1) + 2 * fj + (f & 1);
    a_off[f] = ra * 128;
    a_sw[f] = ra & 7;
    b_off[f] = KG_A_BYTES + rb * 128;
    b_sw[f] = rb & 7;
  }
  for (int it = 0; it < niter; ++it) {
    __syncthreads();
    if (tid == 0 && it + KG_STAGES - 1 < niter) issue(it + KG_STAGES - 1);
    const int s = it % KG_STAGES;
    mbar_wait(bar_base + 8 * s, (uint32_t)((it / KG_STAGES) & 1));
    const uint32_t sa = base + s * KG_STAGE_BYTES;
#pragma unroll
    for (int kk = 0; kk < KG_KB / 4; ++kk) {
      const int chunk = 2 * kk + (fk >> 1);
      const uint32_t lo = (fk & 1) * 8;
      double af[4], bf[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        af[f] = lds64(sa + a_off[f] + ((chunk ^ a_sw[f]) << 4) + lo);
        bf[f] = lds64(sa + b_off[f] + ((chunk ^ b_sw[f]) << 4) + lo);
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
    }
  }
  // ---- epilogue: accumulators -> shared tile -> coalesced global stores -------------------------------
  __syncthreads();
  double* sC = reinterpret_cast<double*>(base_ptr);
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int r = wm * 32 + 16 * (mi >> 1) + 2 * fj + (mi & 1);
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = wn * 32 + 16 * (ni >> 1) + 2 * (2 * fk + e) + (ni & 1);
        sC[r * KG_CPITCH + c] = acc[mi][ni][e];
      }
  }
  __syncthreads();
  if (a.out_col_major) {
    // out[n' * ldc + m]: consecutive threads walk m
    for (int e = tid; e < KG_TM * KG_TN; e += KG_THREADS) {
      const int r = e % KG_TM, c = e / KG_TM;
      const int64_t m = m0 + r, n = n0 + c;
      if (m < a.M && n < a.N) {
        const int64_t nn = a.col_perm ? (int64_t)a.col_perm[n] : n;
        a.out[nn * a.ldc + m] = sC[r * KG_CPITCH + c] + (a.bias ? a.bias[m] : 0.0);
      }
    }
  } else {
    for (int e = tid; e < KG_TM * KG_TN; e += KG_THREADS) {
      const int r = e / KG_TN, c = e % KG_TN;
      const int64_t m = m0 + r, n = n0 + c;
      if (m < a.M && n < a.N) a.out[m * a.ldc + n] = sC[r * KG_CPITCH + c] + (a.bias ? a.bias[m] : 0.0);
    }
  }
}

// per 128-row tile: OR over the rows of (row has a non-zero in k-slice s) << s
__global__ void __launch_bounds__(256) kgemm_occ_kernel(const double* __restrict__ A, int64_t M, int64_t lda, int K,
                                                        unsigned long long* __restrict__ occ) {
  __shared__ unsigned long long sm[8];
  const int64_t m0 = (int64_t)blockIdx.x * KG_TM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long o = 0ull;
  for (int r = warp; r < KG_TM && m0 + r < M; r += 8) {
    const double* row = A + (size_t)(m0 + r) * lda;
    for (int c = lane; c < K; c += 32)
      if (row[c] != 0.0) o |= 1ull << (c / KG_KB);
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) o |= __shfl_xor_sync(0xffffffffu, o, sft);
  if (lane == 0) sm[warp] = o;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0ull;
    for (int w = 0; w < 8; ++w) t |= sm[w];
    occ[blockIdx.x] = t;
  }
}

int launch_kgemm_occ(const double* A, int64_t M, int64_t lda, int K, unsigned long long* occ, cudaStream_t st) {
  if (K > 64 * KG_KB) {
    set_error("kgemm occupancy map: K = %d exceeds 1024", K);
    return BGP_ERR_ARG;
  }
  kgemm_occ_kernel<<<(unsigned)((M + KG_TM - 1) / KG_TM), 256, 0, st>>>(A, M, lda, K, occ);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

int launch_kgemm(const double* A, int64_t M, int64_t lda, const double* B, int64_t N, int64_t ldb, int K,
                 const double* bias, double* out, int64_t ldc, bool out_col_major, const int32_t* col_perm,
                 cudaStream_t st, const unsigned long long* a_occ) {
  if (M <= 0 || N <= 0) return BGP_OK;
  if ((lda & 1) || (ldb & 1) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) {
    set_error("kgemm: operands must be 16-byte aligned with an even row pitch");
    return BGP_ERR_ARG;
  }
  CUtensorMap tmA, tmB;
  if (make_tensormap_f64(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 16, KG_TM) != 0 ||
      make_tensormap_f64(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 16, KG_TN) != 0) {
    set_error("cuTensorMapEncodeTiled failed in kgemm");
    return BGP_ERR_CUDA;
  }
  // per device and cheap: set on every launch (a process may drive several devices)
  BGP_CUDA(cudaFuncSetAttribute(kgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KG_SMEM));
  KgemmArgs a;
  a.M = M;
  a.N = N;
  a.K = K;
  a.bias = bias;
  a.out = out;
  a.ldc = ldc;
  a.out_col_major = out_col_major ? 1 : 0;
  a.col_perm = col_perm;
  a.a_occ = a_occ;
  dim3 grid((unsigned)((M + KG_TM - 1) / KG_TM), (unsigned)((N + KG_TN - 1) / KG_TN));
  kgemm_kernel<<<grid, KG_THREADS, KG_SMEM, st>>>(tmA, tmB, a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

}  // namespace bgp
