// grad.cu — exact gradient of the Laplace objective w.r.t. theta (ff$gr).
//
// Replaces the reverse-mode AD sweep TMB performs for MakeADFun(random = "W")$gr (call sites
// /root/reference/R/02_model_fit.R:276-284: used by aghq's BFGS and by numDeriv::jacobian(ff$gr, .)
// at :283).  Closed form, with Hi = H^-1 at the mode w_hat:
//   dL/dtheta_k = df/dtheta_k + 1/2 tr(Hi dH/dtheta_k) - 1/2 v^T Hi c_k,
//   v = A^T (c3 * q),  q_i = a_i^T Hi a_i (leverages),  c_k = d2f / dW dtheta_k.
// Device work:
//   1. L^-1 by block forward substitution (32x32 blocks, one CTA per block column);
//   2. leverages q_i = || L^-1 a_i ||^2 : a TRMM-shaped FP64 DMMA kernel (n p^2 flops), TMA-staged
//      operands (both K-major, 128B swizzle), Y tiles never leave registers, fused row norms;
//   3. v = A^T (c3 * q) with the streaming kernel of lik.cu.
// The p-sized algebra that remains (traces, two triangular products) runs on the host.
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
constexpr int LV_TM = 128;    // observations per CTA
constexpr int LV_TN = 64;     // rows of L^-1 per pass
constexpr int LV_KB = 16;     // contraction slice per stage (one 128-byte line)
constexpr int LV_STAGES = 3;
constexpr int LV_THREADS = 256;
constexpr int LV_A_BYTES = LV_TM * 128;
constexpr int LV_B_BYTES = LV_TN * 128;
constexpr int LV_STAGE_BYTES = LV_A_BYTES + LV_B_BYTES;
constexpr int LV_SMEM = LV_STAGES * LV_STAGE_BYTES + 64 + 1024 + 2 * LV_TM * 8;

__global__ void __launch_bounds__(LV_THREADS, 2)
    leverage_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmL,
                    const double* __restrict__ c3, double* __restrict__ z, int64_t n, int p, int ldl) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + LV_STAGES * LV_STAGE_BYTES;
  double* sQ = reinterpret_cast<double*>(smem_raw + (base - smem_u32(smem_raw)) + LV_STAGES * LV_STAGE_BYTES + 64);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int fj = lane >> 2, fk = lane & 3;
  const int obs0 = blockIdx.x * LV_TM;
  const int NT = (p + LV_TN - 1) / LV_TN;

  if (tid == 0) {
    for (int s = 0; s < LV_STAGES; ++s) mbar_init(bar_base + 8 * s, 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto ksteps = [&](int nb) {
    const int kmax = min(ldl, (nb + 1) * LV_TN);
    return (kmax + LV_KB - 1) / LV_KB;
  };
  int total = 0;
  for (int nb = 0; nb < NT; ++nb) total += ksteps(nb);

  // producer state (thread 0 only)
  int p_nb = 0, p_ks = 0, p_it = 0;
  auto issue = [&]() {
    const int s = p_it % LV_STAGES;
    const uint32_t bar = bar_base + 8 * s;
    const uint32_t sa = base + s * LV_STAGE_BYTES;
    mbar_expect_tx(bar, LV_STAGE_BYTES);
    tma_load_2d(sa, &tmA, p_ks * LV_KB, obs0, bar);
    tma_load_2d(sa + LV_A_BYTES, &tmL, p_ks * LV_KB, p_nb * LV_TN, bar);
    ++p_it;
    if (++p_ks == ksteps(p_nb)) {
      p_ks = 0;
      ++p_nb;
    }
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

  int c_nb = 0, c_ks = 0;
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
    if (++c_ks == ksteps(c_nb)) {
      // this 128 x 64 slab of Y = A L^-T is complete: fold its squares into the row norms
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          qacc[mi] = fma(acc[mi][ni][0], acc[mi][ni][0], qacc[mi]);
          qacc[mi] = fma(acc[mi][ni][1], acc[mi][ni][1], qacc[mi]);
          acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        }
      c_ks = 0;
      ++c_nb;
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
  CUtensorMap tmA, tmL;
  std::vector<double> hLinv, hv, hw;
};

static int grad_plan_get(bgp_model* m, GradPlan** out) {
  if (m->grad_plan) {
    *out = (GradPlan*)m->grad_plan;
    return BGP_OK;
  }
  GradPlan* gp = new GradPlan();
  m->ldl = round_up(m->p, 16);
  BGP_CUDA(cudaMalloc(&m->Linv, (size_t)m->p * m->ldl * sizeof(double)));
  BGP_CUDA(cudaMemset(m->Linv, 0, (size_t)m->p * m->ldl * sizeof(double)));
  BGP_CUDA(cudaMalloc(&m->zobs, (size_t)(round_up64(m->n, 64) + 64) * sizeof(double)));
  BGP_CUDA(cudaMemset(m->zobs, 0, (size_t)(round_up64(m->n, 64) + 64) * sizeof(double)));
  if (make_tensormap_f64(&gp->tmA, m->A, (uint64_t)m->lda, (uint64_t)m->n, (uint64_t)m->lda, 16, LV_TM) != 0 ||
      make_tensormap_f64(&gp->tmL, m->Linv, (uint64_t)m->ldl, (uint64_t)m->p, (uint64_t)m->ldl, 16, LV_TN) != 0) {
    delete gp;
    set_error("cuTensorMapEncodeTiled failed for the leverage kernel");
    return BGP_ERR_CUDA;
  }
  BGP_CUDA(cudaFuncSetAttribute(leverage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LV_SMEM));
  gp->hLinv.resize((size_t)m->p * m->ldl);
  gp->hv.resize(m->lda);
  gp->hw.resize(m->lda);
  m->grad_plan = gp;
  *out = gp;
  return BGP_OK;
}

void grad_plan_destroy(bgp_model* m) {
  if (m->grad_plan) delete (GradPlan*)m->grad_plan;
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

int laplace_gradient(bgp_model* m, const double* theta, double* grad) {
  GradPlan* gp = nullptr;
  BGP_TRY(grad_plan_get(m, &gp));
  const int p = m->p, ldl = m->ldl;
  const bool gaussian = m->family == BGP_FAMILY_GAUSSIAN;
  const bool has_c3 = m->family == BGP_FAMILY_POISSON || m->family == BGP_FAMILY_BINOMIAL;
  // exact per-observation quantities at the mode (also fixes sumsq for the Gaussian noise theta)
  BGP_TRY(eval_fg_async(m, m->Wmode, theta, true));
  if (!m->factor_is_exact) {
    // the inner solve kept the factor of its last Newton iteration (newton.cu); the gradient differentiates
    // through H^-1, so it gets the factor at the mode itself
    BGP_TRY(launch_hessian(m, theta));
    BGP_TRY(launch_chol_solve(m, false));
    m->factor_is_exact = true;
  }
  // 1. L^-1
  BGP_TRY(launch_trtri(m, m->Linv, ldl, nullptr));
  // 2./3. leverage term v = A^T (c3 * q)
  std::fill(gp->hv.begin(), gp->hv.end(), 0.0);
  if (has_c3) {
    leverage_kernel<<<(unsigned)((m->n + LV_TM - 1) / LV_TM), LV_THREADS, LV_SMEM, m->stream>>>(gp->tmA, gp->tmL, m->c3,
                                                                                                 m->zobs, m->n, p, ldl);
    count_launch();
    BGP_CUDA(cudaGetLastError());
    BGP_TRY(launch_lik(m, m->Wmode, false, 1.0, m->zobs));
    reduce_partials_kernel<<<(m->lda + 31) / 32, 256, 0, m->stream>>>(m->part_g, m->lik_blocks, m->lda, m->red_buf);
    count_launch();
    if (m->world > 1) BGP_TRY(comm_allreduce_sum(m, m->red_buf, (size_t)m->lda));
    BGP_CUDA(cudaMemcpyAsync(gp->hv.data(), m->red_buf, (size_t)m->lda * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  }
  BGP_CUDA(cudaMemcpyAsync(gp->hLinv.data(), m->Linv, (size_t)p * ldl * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  BGP_CUDA(cudaMemcpyAsync(gp->hw.data(), m->Wmode, (size_t)m->lda * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  BGP_CUDA(cudaMemcpyAsync(m->sc_host, m->sc_dev, sizeof(EvalScalars), cudaMemcpyDeviceToHost, m->stream));
  phase_mark(m, PH_OTHER);
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  phase_harvest(m);
  const double sumsq = m->sc_host->sumsq;
  const double* Li = gp->hLinv.data();
  const double* w = gp->hw.data();
  // Hi v = L^-T (L^-1 v)
  std::vector<double> t1(p, 0.0), Hiv(p, 0.0);
  if (has_c3) {
    for (int i = 0; i < p; ++i) {
      double s = 0.0;
      for (int a = 0; a <= i; ++a) s += Li[(size_t)i * ldl + a] * gp->hv[a];
      t1[i] = s;
    }
    for (int i = 0; i < p; ++i) {
      const double ti = t1[i];
      for (int a = 0; a <= i; ++a) Hiv[a] += Li[(size_t)i * ldl + a] * ti;
    }
  }
  double trHiQ = 0.0;
  for (int k = 0; k < m->J; ++k) {
    const RandomBlock& rb = m->rnd[k];
    const int off = rb.off, d = rb.d;
    const double* P = rb.P_host.data();
    const double ek = std::exp(theta[k]);
    std::vector<double> PU(d, 0.0);
    double tr = 0.0;
    if (rb.diag) {
      for (int c = 0; c < d; ++c) PU[c] = P[c] * w[off + c];
      for (int c = 0; c < d; ++c) {
        double s = 0.0;
        for (int i = off + c; i < p; ++i) {
          const double v = Li[(size_t)i * ldl + off + c];
          s += v * v;
        }
        tr += P[c] * s;
      }
    } else {
      for (int c = 0; c < d; ++c) {
        double s = 0.0;
        for (int b = 0; b < d; ++b) s += P[(size_t)b * d + c] * w[off + b];
        PU[c] = s;
      }
      std::vector<double> tmp(d);
      for (int i = off; i < p; ++i) {
        const double* mi = Li + (size_t)i * ldl + off;
        const int dd = std::min(d, i - off + 1);     // L^-1 is lower triangular
        for (int c = 0; c < dd; ++c) {
          double s = 0.0;
          for (int b = 0; b < dd; ++b) s += P[(size_t)b * d + c] * mi[b];
          tmp[c] = s;
        }
        double s = 0.0;
        for (int c = 0; c < dd; ++c) s += mi[c] * tmp[c];
        tr += s;
      }
    }
    double upu = 0.0, hpu = 0.0;
    for (int c = 0; c < d; ++c) {
      upu += w[off + c] * PU[c];
      hpu += Hiv[off + c] * PU[c];
    }
    const double phi = -std::log(rb.alpha) / rb.u;
    const double dfdth = 0.5 * ek * upu - 0.5 * d - 0.5 * phi * std::exp(-0.5 * theta[k]) + 0.5;
    grad[k] = dfdth + 0.5 * ek * tr - 0.5 * ek * hpu;
    trHiQ += ek * tr;
  }
  if (gaussian) {
    const int k = m->S - 1;
    const double tau = std::exp(theta[k]);
    const std::vector<double>& qfix = m->qfix_host;
    for (int c = 0; c < p; ++c) {
      if (qfix[c] == 0.0) continue;
      double s = 0.0;
      for (int i = c; i < p; ++i) {
        const double v = Li[(size_t)i * ldl + c];
        s += v * v;
      }
      trHiQ += qfix[c] * s;
    }
    const double phi = -std::log(m->theta_alpha[k]) / m->theta_u[k];
    const double dfdth = -0.5 * (double)m->n_total + 0.5 * tau * sumsq - 0.5 * phi * std::exp(-0.5 * theta[k]) + 0.5;
    grad[k] = dfdth + 0.5 * ((double)p - trHiQ);
  }
  return BGP_OK;
}

}  // namespace bgp
