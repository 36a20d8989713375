// grad.cu — exact gradient of the Laplace objective w.r.t. theta (ff$gr).
//
// Replaces the reverse-mode AD sweep TMB performs for MakeADFun(random = "W")$gr (call sites
// /root/reference/R/02_model_fit.R:276-284: used by aghq's BFGS and by numDeriv::jacobian(ff$gr, .)
// at :283).  Closed form, with Hi = H^-1 at the mode w_hat:
//   dL/dtheta_k = df/dtheta_k + 1/2 tr(Hi dH/dtheta_k) - 1/2 v^T Hi c_k,
//   v = A^T (c3 * q),  q_i = a_i^T Hi a_i (leverages),  c_k = d2f / dW dtheta_k.
// Device work:
//   1. H = U U^T (Cholesky in reversed order, chol.cu), V = U^-1 by block forward substitution (32x32 blocks,
//      one CTA per block column) — upper triangular, so structural zeros of a_i cut BOTH loops of V a_i;
//   2. leverages q_i = || V a_i ||^2 : a TRMM-shaped FP64 DMMA kernel, TMA-staged operands (both K-major, 128B
//      swizzle), Y tiles never leave registers, fused row norms, empty {chunk x column box} cells skipped;
//   3. v = A^T (c3 * q) with the streaming kernel of lik.cu;
//   4. the p-sized algebra (V v, V^T V v, diag H^-1, block traces) in one CTA.  S doubles go back to the host.
#include <algorithm>

#include "bgp_internal.h"
#include "ptx.cuh"

namespace bgp {

using namespace ptx;

// ---- 1. triangular inverse ----------------------------------------------------------------------
// L: p x ldh column-major lower; Linv: p x ldl row-major lower.
__global__ void __launch_bounds__(32) trtri_diag_kernel(const double* __restrict__ L, int p, int ldh,
                                                        double* __restrict__ Linv, int ldl, double* __restrict__ LinvT) {
  __shared__ double sD[32][33];
  __shared__ double sX[32][33];
  const int b0 = blockIdx.x * 32, c = threadIdx.x;
  const int bn = (p - b0) < 32 ? (p - b0) : 32;
  for (int k = 0; k < 32; ++k) sD[c][k] = (c < bn && k <= c) ? L[(size_t)(b0 + k) * ldh + b0 + c] : 0.0;
  __syncwarp();
  if (c < bn) {
    for (int i = c; i < bn; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int j = c; j < i; ++j) s = fma(-sD[i][j], sX[j][c], s);
      sX[i][c] = s / sD[i][i];
    }
  }
  __syncwarp();
  for (int i = 0; i < bn; ++i)
    if (c < bn) {
      const double v = (c <= i) ? sX[i][c] : 0.0;
      Linv[(size_t)(b0 + i) * ldl + b0 + c] = v;
      if (LinvT) LinvT[(size_t)(b0 + c) * ldl + b0 + i] = v;
    }
}

__global__ void __launch_bounds__(256) trtri_offdiag_kernel(const double* __restrict__ L, int p, int ldh,
                                                            double* __restrict__ Linv, int ldl,
                                                            double* __restrict__ LinvT) {
  __shared__ double sL[32][33];
  __shared__ double sX[32][33];
  __shared__ double sA[32][33];
  __shared__ double sDi[32][33];
  const int cb = blockIdx.x, nb = (p + 31) / 32, t = threadIdx.x;
  const int li = t & 31, lr = t >> 5;                 // load mapping
  const int oi = t >> 3, oj = (t & 7) * 4;            // output mapping: row oi, columns oj..oj+3
  for (int b = cb + 1; b < nb; ++b) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = cb; k < b; ++k) {
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int kk = lr + 8 * r;
        const int gr = b * 32 + li, gk = k * 32 + kk;
        sL[li][kk] = (gr < p && gk < p) ? L[(size_t)gk * ldh + gr] : 0.0;
        const int gc = cb * 32 + li;
        sX[kk][li] = (gk < p && gc < p) ? Linv[(size_t)gk * ldl + gc] : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const double l = sL[oi][kk];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fma(l, sX[kk][oj + e], acc[e]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) sA[oi][oj + e] = acc[e];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = lr + 8 * r;
      const int gi = b * 32 + kk, gj = b * 32 + li;
      sDi[kk][li] = (gi < p && gj < p) ? Linv[(size_t)gi * ldl + gj] : 0.0;
    }
    __syncthreads();
    double out[4] = {0.0, 0.0, 0.0, 0.0};
    for (int mm = 0; mm <= oi; ++mm) {
      const double dv = sDi[oi][mm];
#pragma unroll
      for (int e = 0; e < 4; ++e) out[e] = fma(dv, sA[mm][oj + e], out[e]);
    }
    const int gi = b * 32 + oi;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gj = cb * 32 + oj + e;
      if (gi < p && gj < p) {
        Linv[(size_t)gi * ldl + gj] = -out[e];
        if (LinvT) LinvT[(size_t)gj * ldl + gi] = -out[e];
      }
    }
  }
}

// ---- 2. leverages -----------------------------------------------------------------------------------
// q_i = a_i^T H^-1 a_i = || V a_i ||^2 with V = U^-1 UPPER triangular, H = U U^T (the Cholesky factor of H taken
// in reversed order).  With an upper factor y = V a reads y_k = sum_{c >= k} V[k][c] a_c, so a row whose non-zeros
// end at column m (after the zero-pattern sort the occupied column boxes of a chunk are a prefix) costs m^2 / 2
// instead of the p^2 / 2 of the lower factor: both the contraction slices c and the output rows k beyond the last
// occupied box are skipped, box by box, from the same occupancy words the Hessian kernel uses.
constexpr int LV_TM = 128;    // observations per CTA (two 64-observation chunks of the occupancy map)
constexpr int LV_TN = 64;     // rows of V per pass
constexpr int LV_KB = 16;     // contraction slice per stage (one 128-byte line, one column box)
constexpr int LV_STAGES = 3;
constexpr int LV_THREADS = 256;
constexpr int LV_A_BYTES = LV_TM * 128;
constexpr int LV_B_BYTES = LV_TN * 128;
constexpr int LV_STAGE_BYTES = LV_A_BYTES + LV_B_BYTES;
constexpr int LV_MAXLIST = 16 * 64 / 2 + 64;     // (row block, slice) pairs on or above the diagonal, p <= 1024
constexpr int LV_SMEM = LV_STAGES * LV_STAGE_BYTES + 64 + 1024 + 2 * LV_TM * 8 + LV_MAXLIST * 4;

__global__ void __launch_bounds__(LV_THREADS, 2)
    leverage_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmV,
                    const double* __restrict__ c3, const unsigned long long* __restrict__ occ, int64_t nchunks,
                    double* __restrict__ z, int64_t n, int p, int ldl) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + LV_STAGES * LV_STAGE_BYTES;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw)) + LV_STAGES * LV_STAGE_BYTES + 64;
  double* sQ = reinterpret_cast<double*>(gen);
  // work list: entry = row block << 16 | slice << 1 | last-of-block
  uint32_t* list = reinterpret_cast<uint32_t*>(gen + 2 * LV_TM * 8);
  __shared__ int s_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int fj = lane >> 2, fk = lane & 3;
  const int obs0 = blockIdx.x * LV_TM;
  const int NT = (p + LV_TN - 1) / LV_TN;
  const int nslices = ldl / LV_KB;

  if (tid == 0) {
    for (int s = 0; s < LV_STAGES; ++s) mbar_init(bar_base + 8 * s, 1);
    mbar_fence_init();
    const int64_t c0 = 2 * (int64_t)blockIdx.x;
    unsigned long long o = c0 < nchunks ? occ[c0] : 0ull;
    if (c0 + 1 < nchunks) o |= occ[c0 + 1];
    int cnt = 0;
    for (int nb = 0; nb < NT; ++nb) {
      int first = cnt;
      for (int ks = 4 * nb; ks < nslices; ++ks)
        if ((o >> ks) & 1ull) list[cnt++] = ((uint32_t)nb << 16) | ((uint32_t)ks << 1);
      if (cnt > first) list[cnt - 1] |= 1u;
    }
    s_total = cnt;
  }
  __syncthreads();
  const int total = s_total;

  int p_it = 0;
  auto issue = [&]() {
    const int s = p_it % LV_STAGES;
    const uint32_t bar = bar_base + 8 * s;
    const uint32_t sa = base + s * LV_STAGE_BYTES;
    const uint32_t e = list[p_it];
    const int nb = (int)(e >> 16), ks = (int)((e >> 1) & 0x7fffu);
    mbar_expect_tx(bar, LV_STAGE_BYTES);
    tma_load_2d(sa, &tmA, ks * LV_KB, obs0, bar);
    tma_load_2d(sa + LV_A_BYTES, &tmV, ks * LV_KB, nb * LV_TN, bar);
    ++p_it;
  };
  if (tid == 0)
    for (int i = 0; i < LV_STAGES - 1 && p_it < total; ++i) issue();

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double qacc[4] = {0.0, 0.0, 0.0, 0.0};

  // fragment row offsets inside the boxes (row permutation r = 2 j + (f & 1) + 16 (f >> 1))
  uint32_t a_off[4], b_off[4];
  int a_sw[4], b_sw[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int ra = wm * 32 + 16 * (f >> 1) + 2 * fj + (f & 1);
    const int rb = wn * 32 + 16 * (f >> 1) + 2 * fj + (f & 1);
    a_off[f] = ra * 128;
    a_sw[f] = ra & 7;
    b_off[f] = LV_A_BYTES + rb * 128;
    b_sw[f] = rb & 7;
  }

  for (int it = 0; it < total; ++it) {
    __syncthreads();
    if (tid == 0 && p_it < total) issue();
    const int s = it % LV_STAGES;
    mbar_wait(bar_base + 8 * s, (uint32_t)((it / LV_STAGES) & 1));
    const uint32_t sa = base + s * LV_STAGE_BYTES;
#pragma unroll
    for (int kk = 0; kk < LV_KB / 4; ++kk) {
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
    if (list[it] & 1u) {
      // this 128 x 64 slab of Y = A V^T is complete: fold its squares into the row norms
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          qacc[mi] = fma(acc[mi][ni][0], acc[mi][ni][0], qacc[mi]);
          qacc[mi] = fma(acc[mi][ni][1], acc[mi][ni][1], qacc[mi]);
          acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        }
    }
  }
  // rows are shared by the 4 lanes of a quad and by the two N-halves (wn)
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    qacc[mi] += __shfl_xor_sync(0xffffffffu, qacc[mi], 1);
    qacc[mi] += __shfl_xor_sync(0xffffffffu, qacc[mi], 2);
  }
  __syncthreads();
  if (fk == 0) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int r = wm * 32 + 16 * (mi >> 1) + 2 * fj + (mi & 1);
      sQ[wn * LV_TM + r] = qacc[mi];
    }
  }
  __syncthreads();
  if (tid < LV_TM) {
    const int64_t obs = (int64_t)obs0 + tid;
    if (obs < n) z[obs] = c3[obs] * (sQ[tid] + sQ[LV_TM + tid]);
  }
}

// H in reversed index order (both triangles): Hr[i][j] = H[p-1-i][p-1-j]
__global__ void reverse_sym_kernel(const double* __restrict__ H, int p, int ldh, double* __restrict__ Hr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < p) Hr[(size_t)j * ldh + i] = H[(size_t)(p - 1 - j) * ldh + (p - 1 - i)];
}

// V = U^-1 (row-major p x ldl, upper) from the inverse of the reversed factor (row-major lower):
// V[k][c] = Linv_r[p-1-k][p-1-c]
__global__ void unreverse_tri_kernel(const double* __restrict__ Lr, int p, int ldl, double* __restrict__ V) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (c < ldl) V[(size_t)k * ldl + c] = (c >= k && c < p) ? Lr[(size_t)(p - 1 - k) * ldl + (p - 1 - c)] : 0.0;
}

// ---- 3. the p-sized algebra, one CTA ------------------------------------------------------------------
struct GradSmallArgs {
  const double* V;          // p x ldl row-major upper, H^-1 = V^T V
  int p, ldl, lda, S, J, gaussian;
  const double* v;          // A^T (c3 * q) (lda) or NULL
  const double* W;          // mode (lda)
  const double* qfix;       // theta-independent diagonal of Q
  const EvalScalars* sc;    // sumsq at the mode
  double n_total;
  struct {
    int off, d, diag;
    const double* P;
    double etheta, theta, phi;
  } rnd[16];
  double noise_theta, noise_phi;
  double* scratch;          // 4 * lda doubles
  double* grad;             // S
};

// t1 = V v: one warp per row k, lanes over the columns c >= k
__global__ void __launch_bounds__(256) grad_vv_kernel(const double* __restrict__ V, int p, int ldl, const double* __restrict__ v,
                                                      double* __restrict__ t1) {
  const int lane = threadIdx.x & 31, k = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= p) return;
  double s = 0.0;
  if (v)
    for (int c = k + lane; c < p; c += 32) s = fma(V[(size_t)k * ldl + c], v[c], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) t1[k] = s;
}

// Hiv = V^T t1 and diag(H^-1) = column square norms of V: a CTA owns 32 columns, its 8 warps split the rows
// k <= c (row r goes to warp r % 8), partial sums combined in warp order
__global__ void __launch_bounds__(256) grad_vt_kernel(const double* __restrict__ V, int p, int ldl, const double* __restrict__ t1,
                                                      double* __restrict__ Hiv, double* __restrict__ dHi) {
  __shared__ double sm[2][8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int kmax = min(p, blockIdx.x * 32 + 32);
  double s = 0.0, d = 0.0;
  if (c < p)
    for (int k = warp; k < kmax; k += 8) {
      const double x = V[(size_t)k * ldl + c];          // zero below the diagonal (k > c)
      s = fma(x, t1[k], s);
      d = fma(x, x, d);
    }
  sm[0][warp][lane] = s;
  sm[1][warp][lane] = d;
  __syncthreads();
  if (warp == 0 && c < p) {
    double ts = 0.0, td = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      ts += sm[0][w][lane];
      td += sm[1][w][lane];
    }
    Hiv[c] = ts;
    dHi[c] = td;
  }
}

__global__ void __launch_bounds__(1024) grad_small_kernel(const GradSmallArgs a) {
  __shared__ double s_red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double* Hiv = a.scratch + a.lda;        // V^T (V v), from grad_vt_kernel
  double* dHi = a.scratch + 2 * a.lda;    // diag(H^-1)
  double* PU = a.scratch + 3 * a.lda;     // P_k U_k per block entry
  auto block_sum = [&](double x) -> double {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += s_red[w];       // fixed order
    return t;
  };
  double trHiQ = 0.0;
  for (int b = 0; b < a.J; ++b) {
    const int off = a.rnd[b].off, d = a.rnd[b].d;
    const double* P = a.rnd[b].P;
    double upu = 0.0, hpu = 0.0, tr = 0.0;
    if (a.rnd[b].diag) {
      for (int c = tid; c < d; c += blockDim.x) {
        const double pu = P[c] * a.W[off + c];
        upu = fma(a.W[off + c], pu, upu);
        hpu = fma(Hiv[off + c], pu, hpu);
        tr = fma(P[c], dHi[off + c], tr);
      }
    } else {
      for (int c = tid; c < d; c += blockDim.x) {
        double s = 0.0;
        for (int e = 0; e < d; ++e) s = fma(P[(size_t)e * d + c], a.W[off + e], s);
        PU[c] = s;
        upu = fma(a.W[off + c], s, upu);
        hpu = fma(Hiv[off + c], s, hpu);
      }
      // tr(H^-1_bb P) = sum_k x_k^T P x_k, x_k = V[k][block] (zero left of the diagonal): one row per warp
      for (int k = warp; k < off + d && k < a.p; k += nw) {
        const double* x = a.V + (size_t)k * a.ldl + off;
        const int c_lo = k > off ? k - off : 0;
        double s = 0.0;
        for (int c = c_lo + lane; c < d; c += 32) {
          double t = 0.0;
          for (int e = c_lo; e < d; ++e) t = fma(P[(size_t)e * d + c], x[e], t);
          s = fma(x[c], t, s);
        }
        tr += s;
      }
    }
    upu = block_sum(upu);
    hpu = block_sum(hpu);
    tr = block_sum(tr);
    const double ek = a.rnd[b].etheta;
    if (tid == 0) {
      const double dfdth = 0.5 * ek * upu - 0.5 * d - 0.5 * a.rnd[b].phi * exp(-0.5 * a.rnd[b].theta) + 0.5;
      a.grad[b] = dfdth + 0.5 * ek * tr - 0.5 * ek * hpu;
    }
    trHiQ += ek * tr;
  }
  if (a.gaussian) {
    double t = 0.0;
    for (int c = tid; c < a.p; c += blockDim.x) t = fma(a.qfix[c], dHi[c], t);
    t = block_sum(t);
    if (tid == 0) {
      const double tau = exp(a.noise_theta);
      const double dfdth = -0.5 * a.n_total + 0.5 * tau * a.sc->sumsq - 0.5 * a.noise_phi * exp(-0.5 * a.noise_theta) + 0.5;
      a.grad[a.S - 1] = dfdth + 0.5 * ((double)a.p - (trHiQ + t));
    }
  }
}

int launch_trtri(bgp_model* m, double* Linv, int ldl, double* LinvT) {
  const int p = m->p, nb = (p + 31) / 32;
  trtri_diag_kernel<<<nb, 32, 0, m->stream>>>(m->L, p, m->ldh, Linv, ldl, LinvT);
  count_launch();
  if (nb > 1) {
    trtri_offdiag_kernel<<<nb - 1, 256, 0, m->stream>>>(m->L, p, m->ldh, Linv, ldl, LinvT);
    count_launch();
  }
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

struct GradPlan {
  CUtensorMap tmA, tmV;
  double* Hrev = nullptr;      // p x ldh: H in reversed index order
  double* V = nullptr;         // p x ldl row-major upper: U^-1
  double* scratch = nullptr;   // 4 lda + S doubles
  double* grad_host = nullptr; // pinned, S doubles + chol info
};

static int grad_plan_get(bgp_model* m, GradPlan** out) {
  if (m->grad_plan) {
    *out = (GradPlan*)m->grad_plan;
    return BGP_OK;
  }
  GradPlan* gp = new GradPlan();
  m->ldl = round_up(m->p, 16);
  const size_t zb = (size_t)(round_up64(m->n, 64) + 64) * sizeof(double);
  BGP_CUDA(cudaMalloc(&m->Linv, (size_t)m->p * m->ldl * sizeof(double)));
  BGP_CUDA(cudaMemsetAsync(m->Linv, 0, (size_t)m->p * m->ldl * sizeof(double), m->stream));   // in stream order: the stream does not wait for the legacy default stream
  BGP_CUDA(cudaMalloc(&gp->V, (size_t)round_up(m->p, 64) * m->ldl * sizeof(double)));
  BGP_CUDA(cudaMemsetAsync(gp->V, 0, (size_t)round_up(m->p, 64) * m->ldl * sizeof(double), m->stream));
  BGP_CUDA(cudaMalloc(&gp->Hrev, (size_t)m->p * m->ldh * sizeof(double)));
  BGP_CUDA(cudaMemsetAsync(gp->Hrev, 0, (size_t)m->p * m->ldh * sizeof(double), m->stream));
  BGP_CUDA(cudaMalloc(&gp->scratch, ((size_t)4 * m->lda + 32) * sizeof(double)));
  BGP_CUDA(cudaMemsetAsync(gp->scratch, 0, ((size_t)4 * m->lda + 32) * sizeof(double), m->stream));
  BGP_CUDA(cudaMallocHost(&gp->grad_host, 32 * sizeof(double)));
  BGP_CUDA(cudaMalloc(&m->zobs, zb));
  BGP_CUDA(cudaMemsetAsync(m->zobs, 0, zb, m->stream));
  if (make_tensormap_f64(&gp->tmA, m->A, (uint64_t)m->lda, (uint64_t)m->n, (uint64_t)m->lda, 16, LV_TM) != 0 ||
      make_tensormap_f64(&gp->tmV, gp->V, (uint64_t)m->ldl, (uint64_t)m->p, (uint64_t)m->ldl, 16, LV_TN) != 0) {
    delete gp;
    set_error("cuTensorMapEncodeTiled failed for the leverage kernel");
    return BGP_ERR_CUDA;
  }
  BGP_CUDA(cudaFuncSetAttribute(leverage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LV_SMEM));
  {
    // executed work of one leverage launch: (row block, slice) pairs on or above the diagonal whose slice is
    // occupied in the CTA's two chunks, 128 x 64 x 16 multiply-adds each
    const int NT = (m->p + LV_TN - 1) / LV_TN, nslices = m->ldl / LV_KB;
    double pairs = 0.0;
    for (int64_t c = 0; c < m->nchunks; c += 2) {
      unsigned long long o = m->occ_host[(size_t)c];
      if (c + 1 < m->nchunks) o |= m->occ_host[(size_t)c + 1];
      for (int nb = 0; nb < NT; ++nb)
        for (int ks = 4 * nb; ks < nslices; ++ks) pairs += (double)((o >> ks) & 1ull);
    }
    m->lev_flops = pairs * 2.0 * LV_TM * LV_TN * LV_KB;
  }
  m->grad_plan = gp;
  *out = gp;
  return BGP_OK;
}

void grad_plan_destroy(bgp_model* m) {
  if (m->grad_plan) {
    GradPlan* gp = (GradPlan*)m->grad_plan;
    if (gp->Hrev) cudaFree(gp->Hrev);
    if (gp->V) cudaFree(gp->V);
    if (gp->scratch) cudaFree(gp->scratch);
    if (gp->grad_host) cudaFreeHost(gp->grad_host);
    delete gp;
  }
  m->grad_plan = nullptr;
  if (m->Linv) cudaFree(m->Linv);
  if (m->zobs) cudaFree(m->zobs);
  m->Linv = m->zobs = nullptr;
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ part_g, int nblocks, int lda,
                                                              double* __restrict__ red) {
  __shared__ double sm[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (c < lda)
    for (int b = warp; b < nblocks; b += 8) s += part_g[(size_t)b * lda + c];
  sm[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < lda) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    red[c] = t;
  }
}

// Everything stays on the device; S doubles come back.  Sequence: (likelihood pass at the mode unless the inner
// solve just left one there) -> (Hessian at the mode unless it is in memory) -> reversed-order Cholesky ->
// triangular inverse -> V -> leverages -> A^T (c3 q) -> the p-sized algebra in one CTA.
int laplace_gradient(bgp_model* m, const double* theta, double* grad) {
  GradPlan* gp = nullptr;
  BGP_TRY(grad_plan_get(m, &gp));
  const int p = m->p, ldl = m->ldl;
  const bool gaussian = m->family == BGP_FAMILY_GAUSSIAN;
  const bool has_c3 = m->family == BGP_FAMILY_POISSON || m->family == BGP_FAMILY_BINOMIAL;
  if (p > 1024) {
    set_error("laplace gradient: p = %d exceeds the leverage kernel's work list (1024)", p);
    return BGP_ERR_ARG;
  }
  // A model on the O-spline moment path (ospline.cu) stays on it: pass and Hessian at the mode from the moments, the
  // leverages from per-interval quadratic forms (osp_launch_leverage).  bgp_model_set_ospline(m, 2) keeps the dense
  // passes for the gradient (A/B): the design rows are resident either way.
  const bool osp_lev = m->osp_on && !m->osp_dense_grad;
  struct DenseScope {
    bgp_model* m;
    bool was;
    DenseScope(bgp_model* mm, bool act) : m(mm), was(act) {
      if (was) {
        m->osp_on = false;
        m->obs_at_mode = false;
      }
    }
    ~DenseScope() {
      if (was) {
        m->osp_on = true;
        m->obs_at_mode = false;      // the dense arrays, not the moments, hold the mode's quantities
      }
    }
  } dense_scope(m, m->osp_on && m->osp_dense_grad);
  if (dense_scope.was && !m->factor_is_exact) {
    // Hessian at the mode from the moment path (one pass over 40-60 bytes per observation) before switching over
    m->osp_on = true;
    BGP_TRY(eval_fg_async(m, m->Wmode, theta, false));
    phase_mark(m, PH_HESS);
    BGP_TRY(launch_hessian(m, theta));
    m->n_hess++;
    m->factor_is_exact = true;
    m->osp_on = false;
  }
  // exact per-observation quantities at the mode (w, c3; sumsq for the Gaussian noise theta)
  if (!m->obs_at_mode) {
    BGP_TRY(eval_fg_async(m, m->Wmode, theta, true));
    m->obs_at_mode = true;
  }
  if (!m->factor_is_exact) {
    // the inner solve kept the Hessian of its last Newton iteration (newton.cu); the gradient differentiates
    // through H^-1, so it gets the Hessian at the mode itself
    phase_mark(m, PH_HESS);
    BGP_TRY(launch_hessian(m, theta));
    m->n_hess++;
    m->factor_is_exact = true;
  }
  // H = U U^T: Cholesky of H in reversed order (the factor left in m->L is that of the reversed matrix from here on)
  phase_mark(m, PH_CHOL);
  {
    dim3 grid((p + 255) / 256, p);
    reverse_sym_kernel<<<grid, 256, 0, m->stream>>>(m->H, p, m->ldh, gp->Hrev);
    count_launch();
    double* keep = m->H;
    m->L_holds_H = false;            // the factor below is that of the reversed matrix
    m->H = gp->Hrev;
    const int st = launch_chol_solve(m, false);
    m->H = keep;
    m->n_chol++;
    BGP_TRY(st);
    m->L_is_reversed = true;
  }
  BGP_TRY(launch_trtri(m, m->Linv, ldl, nullptr));
  {
    dim3 grid((ldl + 255) / 256, p);
    unreverse_tri_kernel<<<grid, 256, 0, m->stream>>>(m->Linv, p, ldl, gp->V);
    count_launch();
  }
  phase_mark(m, PH_OTHER);
  const double* v_dev = nullptr;
  if (has_c3 && osp_lev) {
    phase_mark(m, PH_LEV);
    BGP_TRY(osp_launch_leverage(m, gp->V, ldl));
    m->n_lev++;
    if (m->world > 1) BGP_TRY(comm_allreduce_sum(m, m->red_buf, (size_t)m->lda));
    v_dev = m->red_buf;
    m->obs_at_mode = false;          // the interval moments now belong to the leverage pass
    phase_mark(m, PH_OTHER);
  } else if (has_c3) {
    phase_mark(m, PH_LEV);
    leverage_kernel<<<(unsigned)((m->n + LV_TM - 1) / LV_TM), LV_THREADS, LV_SMEM, m->stream>>>(
        gp->tmA, gp->tmV, m->c3, (const unsigned long long*)m->occ_dev, m->nchunks, m->zobs, m->n, p, ldl);
    count_launch();
    BGP_CUDA(cudaGetLastError());
    m->n_lev++;
    phase_mark(m, PH_LIK);
    m->n_lik++;
    BGP_TRY(launch_lik(m, m->Wmode, false, 1.0, m->zobs));
    reduce_partials_kernel<<<(m->lda + 31) / 32, 256, 0, m->stream>>>(m->part_g, m->lik_blocks, m->lda, m->red_buf);
    count_launch();
    if (m->world > 1) BGP_TRY(comm_allreduce_sum(m, m->red_buf, (size_t)m->lda));
    v_dev = m->red_buf;
    phase_mark(m, PH_OTHER);
  }
  GradSmallArgs a;
  memset(&a, 0, sizeof(a));
  a.V = gp->V;
  a.p = p;
  a.ldl = ldl;
  a.lda = m->lda;
  a.S = m->S;
  a.J = m->J;
  a.gaussian = gaussian ? 1 : 0;
  a.v = v_dev;
  a.W = m->Wmode;
  a.qfix = m->qfix;
  a.sc = m->sc_dev;
  a.n_total = (double)m->n_total;
  for (int k = 0; k < m->J; ++k) {
    const RandomBlock& rb = m->rnd[k];
    a.rnd[k].off = rb.off;
    a.rnd[k].d = rb.d;
    a.rnd[k].diag = rb.diag ? 1 : 0;
    a.rnd[k].P = rb.P_dev;
    a.rnd[k].etheta = std::exp(theta[k]);
    a.rnd[k].theta = theta[k];
    a.rnd[k].phi = -std::log(rb.alpha) / rb.u;
  }
  if (gaussian) {
    a.noise_theta = theta[m->S - 1];
    a.noise_phi = -std::log(m->theta_alpha[m->S - 1]) / m->theta_u[m->S - 1];
  }
  a.scratch = gp->scratch;
  a.grad = gp->scratch + 4 * m->lda;
  grad_vv_kernel<<<(p + 7) / 8, 256, 0, m->stream>>>(gp->V, p, ldl, v_dev, a.scratch);
  grad_vt_kernel<<<(p + 31) / 32, 256, 0, m->stream>>>(gp->V, p, ldl, a.scratch, a.scratch + m->lda, a.scratch + 2 * m->lda);
  count_launch(2);
  grad_small_kernel<<<1, 1024, 0, m->stream>>>(a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  BGP_CUDA(cudaMemcpyAsync(gp->grad_host, a.grad, (size_t)m->S * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  BGP_CUDA(cudaMemcpyAsync(m->sc_host, m->sc_dev, sizeof(EvalScalars), cudaMemcpyDeviceToHost, m->stream));
  phase_mark(m, PH_OTHER);
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  phase_harvest(m);
  static const int grad_debug = getenv("BGP_GRAD_DEBUG") ? atoi(getenv("BGP_GRAD_DEBUG")) : 0;     // env: diagnostics (2: NaN only)
  if (grad_debug == 1 || (grad_debug == 2 && (m->sc_host->chol_info != 0 || !(gp->grad_host[0] == gp->grad_host[0]))))
    fprintf(stderr, "[grad] chol_info %d logdet %.17g f %.17g sumsq %.17g nonfinite %d grad0 %.17g exact %d at_mode %d\n",
            m->sc_host->chol_info, m->sc_host->logdet, m->sc_host->f, m->sc_host->sumsq, m->sc_host->nonfinite, gp->grad_host[0],
            (int)m->factor_is_exact, (int)m->obs_at_mode);
  if (m->sc_host->chol_info != 0) {
    // H was positive definite in the natural order; a failure of the reversed factorisation is a rounding accident
    // on a numerically singular H: NaN gradient, as TMB answers when its factorisation fails
    for (int k = 0; k < m->S; ++k) grad[k] = NAN;
    return BGP_OK;
  }
  for (int k = 0; k < m->S; ++k) grad[k] = gp->grad_host[k];
  return BGP_OK;
}

}  // namespace bgp
