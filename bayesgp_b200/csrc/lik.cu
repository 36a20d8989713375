// lik.cu — the per-observation pass of the latent-Gaussian objective (HBM-streaming).
//
// Replaces, for one (W, theta):
//   eta = sum Xf*beta_f + sum X*beta + sum B*U          /root/reference/src/BayesGP.cpp:133-145
//   ll  = Gaussian / Poisson / Binomial log-likelihood   /root/reference/src/BayesGP.cpp:155-168
// and the first-order AD sweep TMB runs on them: r = d ll/d eta, w = -d2 ll/d eta2,
// c3 = d w/d eta, g_lik = A^T r.
//
// Data layout: A is observation-major (n x lda, lda % 16 == 0, zero padded), so one warp
// streams one observation row with fully coalesced 16-byte loads, keeps the row in registers,
// reduces eta with warp shuffles, and accumulates its share of A^T r in registers: A is read
// from HBM exactly once per evaluation (8*lda bytes / observation + 8..16 bytes of y/size).
// Per-block partials are written out and reduced in a fixed order by finish.cu, so results are
// bit-reproducible run to run.
#include "bgp_internal.h"

namespace bgp {

struct LikArgs {
  const double* A;
  int lda;
  int64_t n;
  const double* W;
  const double* y;
  const double* size;
  int family;
  double tau;
  double* eta;
  double* wobs;
  double* c3;
  double* part_g;   // [gridDim.x][lda]
  double* part_s;   // [gridDim.x][4] : ll, sumsq, nonfinite, unused
  const double* rvec;   // if set: skip the likelihood and accumulate A^T rvec only (leverage term)
};

constexpr int LIK_THREADS = 256;
constexpr int LIK_WARPS = LIK_THREADS / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// per-observation likelihood pieces; all lanes of the warp compute the same values
__device__ __forceinline__ void obs_terms(int family, double tau, double eta, double yv, double sz, double& ll,
                                          double& sumsq, double& r, double& w, double& c3) {
  if (family == BGP_FAMILY_POISSON) {                 // dpois(y, exp(eta), log): y*eta - exp(eta) - lgamma(y+1)
    const double mu = exp(eta);
    ll += yv * eta - mu;
    r = yv - mu;
    w = mu;
    c3 = mu;
  } else if (family == BGP_FAMILY_BINOMIAL) {         // dbinom_robust(y, size, eta, log)
    const double e = exp(-fabs(eta));
    const double l1p = log1p(e);
    const double lse_pos = fmax(eta, 0.0) + l1p;      // log(1 + e^eta)
    const double lse_neg = fmax(-eta, 0.0) + l1p;     // log(1 + e^-eta)
    ll += -yv * lse_neg - (sz - yv) * lse_pos;
    const double inv = 1.0 / (1.0 + e);
    const double pi = eta >= 0.0 ? inv : e * inv;
    const double om = eta >= 0.0 ? e * inv : inv;     // 1 - pi
    w = sz * pi * om;
    r = yv - sz * pi;
    c3 = w * (om - pi);
  } else if (family == BGP_FAMILY_GAUSSIAN) {         // dnorm(y, eta, exp(-theta_S/2), log)
    const double res = yv - eta;
    sumsq += res * res;
    r = tau * res;
    w = tau;
    c3 = 0.0;
  } else {
    r = 0.0;
    w = 0.0;
    c3 = 0.0;
  }
}

template <int NJ, int R>
__global__ void __launch_bounds__(LIK_THREADS) lik_kernel(const LikArgs a) {
  extern __shared__ double sm[];   // [LIK_WARPS][lda] partial g, then [LIK_WARPS][4] scalars
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lda = a.lda;
  double2 wv[NJ], ga[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = 2 * lane + 64 * j;
    wv[j] = c < lda ? *reinterpret_cast<const double2*>(a.W + c) : make_double2(0.0, 0.0);
    ga[j] = make_double2(0.0, 0.0);
  }
  double ll = 0.0, sumsq = 0.0;
  int bad = 0;
  const int64_t gw = (int64_t)blockIdx.x * LIK_WARPS + warp;
  const int64_t stride = (int64_t)gridDim.x * LIK_WARPS * R;
  for (int64_t base = gw * R; base < a.n; base += stride) {
    double2 av[R][NJ];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = base + r;
      const double* rp = a.A + row * (int64_t)lda;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = 2 * lane + 64 * j;
        av[r][j] = (row < a.n && c < lda) ? __ldcs(reinterpret_cast<const double2*>(rp + c)) : make_double2(0.0, 0.0);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = base + r;
      if (row < a.n) {
        double rr;
        if (a.rvec) {
          rr = __ldg(a.rvec + row);
        } else {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            s = fma(av[r][j].x, wv[j].x, s);
            s = fma(av[r][j].y, wv[j].y, s);
          }
          s = warp_sum(s);
          const double yv = __ldg(a.y + row);
          const double sz = a.size ? __ldg(a.size + row) : 1.0;
          double ww, cc;
          obs_terms(a.family, a.tau, s, yv, sz, ll, sumsq, rr, ww, cc);
          if (!(isfinite(ww) && isfinite(rr) && isfinite(ll))) bad = 1;
          if (lane == 0) {
            a.eta[row] = s;
            a.wobs[row] = ww;
            if (a.c3) a.c3[row] = cc;
          }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          ga[j].x = fma(rr, av[r][j].x, ga[j].x);
          ga[j].y = fma(rr, av[r][j].y, ga[j].y);
        }
      }
    }
  }
  // ---- block reduction in a fixed order ------------------------------------------------------
  double* sg = sm + (size_t)warp * lda;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = 2 * lane + 64 * j;
    if (c < lda) *reinterpret_cast<double2*>(sg + c) = ga[j];
  }
  double* ss = sm + (size_t)LIK_WARPS * lda;
  if (lane == 0) {
    ss[warp * 4 + 0] = ll;
    ss[warp * 4 + 1] = sumsq;
    ss[warp * 4 + 2] = (double)bad;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < lda; c += LIK_THREADS) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < LIK_WARPS; ++w8) s += sm[(size_t)w8 * lda + c];
    a.part_g[(size_t)blockIdx.x * lda + c] = s;
  }
  if (threadIdx.x < 3) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < LIK_WARPS; ++w8) s += ss[w8 * 4 + threadIdx.x];
    a.part_s[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}

template <int NJ, int R>
static int launch_lik_t(bgp_model* m, const LikArgs& a) {
  const size_t smem = ((size_t)LIK_WARPS * a.lda + LIK_WARPS * 4) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    BGP_CUDA(cudaFuncSetAttribute(lik_kernel<NJ, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  lik_kernel<NJ, R><<<m->lik_blocks, LIK_THREADS, smem, m->stream>>>(a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

int lik_max_lda() { return 1024; }

int launch_lik(bgp_model* m, const double* W_dev, bool want_c3, double tau, const double* rvec) {
  LikArgs a;
  a.rvec = rvec;
  a.A = m->A;
  a.lda = m->lda;
  a.n = m->n;
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
  const int nj = (m->lda + 63) / 64;
  switch (nj) {
    case 1: return launch_lik_t<1, 4>(m, a);
    case 2: return launch_lik_t<2, 4>(m, a);
    case 3: return launch_lik_t<3, 2>(m, a);
    case 4: return launch_lik_t<4, 2>(m, a);
    case 5: return launch_lik_t<5, 2>(m, a);
    case 6: return launch_lik_t<6, 2>(m, a);
    case 7: return launch_lik_t<7, 2>(m, a);
    case 8: return launch_lik_t<8, 2>(m, a);
    case 9: case 10: return launch_lik_t<10, 1>(m, a);
    case 11: case 12: return launch_lik_t<12, 1>(m, a);
    case 13: case 14: return launch_lik_t<14, 1>(m, a);
    case 15: case 16: return launch_lik_t<16, 1>(m, a);
    default:
      set_error("latent dimension p = %d exceeds the supported maximum of %d", m->p, lik_max_lda());
      return BGP_ERR_ARG;
  }
}

}  // namespace bgp
