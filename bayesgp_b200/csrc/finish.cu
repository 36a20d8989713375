// finish.cu — fixed-order reduction of the likelihood partials and the Gaussian-prior terms.
//
// Completes f(W,theta) = -(ll + lpW + lpT) and g = -A^T r + Q(theta)(W - mu0):
//   lpW, Q(theta) = blockdiag(e^{theta_j} P_j, betaprec, beta_fixed_prec)   src/BayesGP.cpp:219-238
//   lpT (theta only; computed on the host and passed in `theta_const`)      src/BayesGP.cpp:241-246
// Two tiny kernels so that, when observations are sharded across GPUs, the packed buffer
// [g_lik (lda) | ll | sumsq | nonfinite | -] can be all-reduced between them.
#include "bgp_internal.h"
#include "finish_dev.cuh"

namespace bgp {

// grid = ceil(lda / 32) CTAs of 256 threads: CTA c owns 32 columns, its 8 warps split the partial blocks
// (block b goes to warp b % 8), partial sums are combined in a fixed order => deterministic.
__global__ void __launch_bounds__(256) finish_reduce_kernel(const double* __restrict__ part_g,
                                                            const double* __restrict__ part_s, int nblocks, int lda,
                                                            double* __restrict__ red) {
  __shared__ double sm[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (c < lda) {
    // four loads in flight, added in ascending block order (the order of the sum is what makes it reproducible)
    for (int b = warp; b < nblocks; b += 32) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = b + 8 * u < nblocks ? part_g[(size_t)(b + 8 * u) * lda + c] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) s += v[u];
    }
  }
  sm[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < lda) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    red[c] = t;
  }
  if (blockIdx.x == 0 && warp == 2) {
    // the four scalars: lane l takes blocks l, l + 32, ...; lanes are combined by a fixed shuffle tree
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = lane; b < nblocks; b += 32) {
      const double4 q = *reinterpret_cast<const double4*>(part_s + (size_t)b * 4);
      t[0] += q.x;
      t[1] += q.y;
      t[2] += q.z;
      t[3] = fmax(t[3], q.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t[0] += __shfl_xor_sync(0xffffffffu, t[0], o);
      t[1] += __shfl_xor_sync(0xffffffffu, t[1], o);
      t[2] += __shfl_xor_sync(0xffffffffu, t[2], o);
      t[3] = fmax(t[3], __shfl_xor_sync(0xffffffffu, t[3], o));
    }
    // [3]: max |eta - previous eta| over the local rows; the SUM all-reduce of sharded models turns it into an
    // upper bound of the global maximum
    if (lane < 4) red[lda + lane] = lane == 0 ? t[0] : (lane == 1 ? t[1] : (lane == 2 ? t[2] : t[3]));
  }
}

__global__ void __launch_bounds__(1024) finish_prior_kernel(const PriorArgs a) {
  __shared__ double s_quad[1024];
  __shared__ double s_gmax[1024];
  finish_prior_body<1024>(a, s_quad, s_gmax);
}

// host side -------------------------------------------------------------------------------------
double theta_constant(const bgp_model* m, const double* theta) {
  // lpT  (src/BayesGP.cpp:241-246)
  double c = 0.0;
  for (int i = 0; i < m->S; ++i) {
    const double phi = -std::log(m->theta_alpha[i]) / m->theta_u[i];
    c += std::log(0.5 * phi) - phi * std::exp(-0.5 * theta[i]) - 0.5 * theta[i];
  }
  // 1/2 (d_j theta_j + logPdet_j)  (src/BayesGP.cpp:229-231)
  for (int j = 0; j < m->J; ++j) c += 0.5 * (m->rnd[j].d * theta[j] + m->rnd[j].logPdet);
  // likelihood constants: -sum lgamma(y+1) | sum lchoose | -n/2 log(2 pi) + n/2 theta_S
  c += m->ll_const;
  if (m->family == BGP_FAMILY_GAUSSIAN) c += 0.5 * (double)m->n_total * theta[m->S - 1];
  return c;
}

void fill_prior_args(bgp_model* m, const double* W_dev, const double* theta, double tau, PriorArgs* out) {
  PriorArgs& a = *out;
  a.red = m->red_buf;
  a.lda = m->lda;
  a.p = m->p;
  a.W = W_dev;
  a.mu0 = m->mu0;
  a.qfix = m->qfix;
  a.g = m->g;
  a.sc = m->sc_dev;
  a.theta_const = theta_constant(m, theta);
  a.tau = tau;
  a.family = m->family;
  a.nrnd = m->J;
  for (int j = 0; j < m->J; ++j) {
    a.rnd[j].off = m->rnd[j].off;
    a.rnd[j].d = m->rnd[j].d;
    a.rnd[j].diag = m->rnd[j].diag ? 1 : 0;
    a.rnd[j].P = m->rnd[j].P_dev;
    a.rnd[j].etheta = std::exp(theta[j]);
  }
}

int launch_finish(bgp_model* m, const double* W_dev, const double* theta, double tau) {
  if (!m->osp_on) {               // the O-spline pass leaves its sums in red_buf itself
    finish_reduce_kernel<<<(m->lda + 31) / 32, 256, 0, m->stream>>>(m->part_g, m->part_s, m->lik_blocks, m->lda, m->red_buf);
    count_launch();
  }
  if (m->world > 1) BGP_TRY(comm_allreduce_sum(m, m->red_buf, (size_t)m->lda + 4));
  PriorArgs a;
  fill_prior_args(m, W_dev, theta, tau, &a);
  finish_prior_kernel<<<1, 1024, 0, m->stream>>>(a);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

}  // namespace bgp
