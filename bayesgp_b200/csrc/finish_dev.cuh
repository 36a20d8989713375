// Device side of the Gaussian-prior completion shared by finish.cu (its own kernel) and ospline.cu (run by the last
// CTA of the moment path's assembly kernel):  f = -(ll + lpW + lpT), g = -A^T r + Q(theta)(W - mu0), max|g|
// (src/BayesGP.cpp:219-246).
#pragma once

#include "bgp_internal.h"

namespace bgp {

constexpr int MAX_RND = 16;

struct RndDev {
  int off, d, diag;
  const double* P;
  double etheta;
};

struct PriorArgs {
  const double* red;     // [lda + 4]
  int lda, p;
  const double* W;
  const double* mu0;
  const double* qfix;
  double* g;
  EvalScalars* sc;
  double theta_const;    // lpT + 1/2 sum(d_j theta_j + logPdet_j) + likelihood constants
  double tau;
  int family;
  int nrnd;
  RndDev rnd[MAX_RND];
};

// one CTA of NT threads; `red` is read through L2 (__ldcg): the caller may have produced it in the same kernel
template <int NT>
__device__ __forceinline__ void finish_prior_body(const PriorArgs& a, double* s_quad, double* s_gmax) {
  double quad = 0.0, gmax = 0.0;
  for (int c = threadIdx.x; c < a.lda; c += NT) {
    double gv = 0.0;
    if (c < a.p) {
      const double dW = a.W[c] - a.mu0[c];
      double q = a.qfix[c] * dW;
      for (int b = 0; b < a.nrnd; ++b) {
        const RndDev& rb = a.rnd[b];
        if (c >= rb.off && c < rb.off + rb.d) {
          const int i = c - rb.off;
          if (rb.diag) {
            q = rb.etheta * rb.P[i] * dW;
          } else {
            double s = 0.0;
            for (int k = 0; k < rb.d; ++k) s = fma(rb.P[(size_t)k * rb.d + i], a.W[rb.off + k], s);
            q = rb.etheta * s;
          }
        }
      }
      gv = -__ldcg(a.red + c) + q;
      quad = fma(dW, q, quad);
      gmax = fmax(gmax, fabs(gv));
      if (!isfinite(gv)) gmax = INFINITY;
    }
    a.g[c] = gv;
  }
  s_quad[threadIdx.x] = quad;
  s_gmax[threadIdx.x] = gmax;
  __syncthreads();
  for (int o = NT / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_quad[threadIdx.x] += s_quad[threadIdx.x + o];
      s_gmax[threadIdx.x] = fmax(s_gmax[threadIdx.x], s_gmax[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double ll_raw = __ldcg(a.red + a.lda + 0), sumsq = __ldcg(a.red + a.lda + 1), bad = __ldcg(a.red + a.lda + 2);
    const double ll = a.family == BGP_FAMILY_GAUSSIAN ? -0.5 * a.tau * sumsq : ll_raw;
    const double f = -(ll + a.theta_const - 0.5 * s_quad[0]);
    a.sc->f = f;
    a.sc->ll = ll;
    a.sc->gmax = s_gmax[0];
    a.sc->quad = s_quad[0];
    a.sc->sumsq = sumsq;
    a.sc->pad = __ldcg(a.red + a.lda + 3);
    a.sc->nonfinite = (bad != 0.0 || !isfinite(f)) ? 1 : 0;
  }
}

// host: the prior terms of model m at (W_dev, theta) (finish.cu)
void fill_prior_args(bgp_model* m, const double* W_dev, const double* theta, double tau, PriorArgs* a);

}  // namespace bgp
