// sample.cu — posterior sampling: replaces aghq::sample_marginal(mod, M)
// (call site /root/reference/R/02_model_fit.R:687-689; algorithm SURVEY.md Appendix A.6):
//   node j(m) ~ Multinomial(lambda),  lambda_j = w_j exp(logpost_normalized_j)
//   W_m = mode_j + R_j^-1 z_m,  R_j = chol(H_j) upper  (un-permuted convention),  z_m ~ N(0, I_p)
// Per node: Cholesky (chol.cu) -> explicit L^-T (grad.cu) -> one DMMA GEMM over that node's
// samples (kgemm.cu) with the mode as bias; samples are written straight into their original
// column of the p x M output (as samps$samps).  The result also stays resident on the device
// for the predict kernels.
#include <algorithm>

#include "bgp_internal.h"

namespace bgp {

// Zg[g][i] = Z[perm[g]][i]   (Z: p x M column-major, Zg: M x ld row-major, zero padded)
__global__ void gather_z_kernel(const double* __restrict__ Z, int p, int64_t M, const int32_t* __restrict__ perm,
                                double* __restrict__ Zg, int ld) {
  const int64_t g = blockIdx.x;
  const int64_t src = perm[g];
  for (int i = threadIdx.x; i < ld; i += blockDim.x) Zg[g * ld + i] = i < p ? Z[src * p + i] : 0.0;
}

// counter-based generator: philox4x32-10
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                           uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// Zg[g][i] = N(0,1) draw number (perm[g] * p + i) of stream `seed` (Box-Muller on 2 x 53-bit uniforms)
__global__ void normal_z_kernel(uint64_t seed, int p, int64_t M, const int32_t* __restrict__ perm, double* __restrict__ Zg,
                                int ld) {
  const int64_t g = blockIdx.x;
  const uint64_t src = (uint64_t)perm[g];
  for (int i = threadIdx.x; i < ld; i += blockDim.x) {
    double v = 0.0;
    if (i < p) {
      const uint64_t ctr = src * (uint64_t)p + (uint64_t)i;
      uint32_t o[4];
      philox4x32((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
      const uint64_t a = ((uint64_t)o[0] << 32) | o[1], b = ((uint64_t)o[2] << 32) | o[3];
      const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
      const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
      v = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    Zg[g * ld + i] = v;
  }
}

// first node whose Hessian failed to factor: flag[0] = node + 1, flag[1] = pivot index (doubles: the flag travels
// through the node group's SUM all-reduce)
__global__ void note_chol_info_kernel(const EvalScalars* __restrict__ sc, int node, double* __restrict__ flag) {
  if (sc->chol_info != 0 && flag[0] == 0.0) {
    flag[0] = (double)(node + 1);
    flag[1] = (double)sc->chol_info;
  }
}

void fit_release_device(bgp_fit* f) {
  if (!f) return;
  // a fit may outlive its model (garbage-collected host objects): then the pools are gone and the blocks are freed here
  const bool model_ok = f->model && f->model_alive && *f->model_alive;
  if (model_ok) cudaSetDevice(f->model->device);
  for (double** p : {&f->samps_dev, &f->Linv_dev, &f->LinvT_dev, &f->mode_dev})
    if (*p) {
      cudaFree(*p);
      *p = nullptr;
    }
  if (model_ok) {
    pool_give(f->model, false, f->modes_dev, f->modes_dev_bytes);
    pool_give(f->model, false, f->Hs_dev, f->Hs_dev_bytes);
    pool_give(f->model, true, f->mirror, f->mirror_bytes);
  } else {
    if (f->modes_dev) cudaFree(f->modes_dev);
    if (f->Hs_dev) cudaFree(f->Hs_dev);
    if (f->mirror) cudaFreeHost(f->mirror);
  }
  f->modes_dev = f->Hs_dev = f->mirror = nullptr;
}

static int sample_core(bgp_fit* f, int64_t M, const double* Z_host, uint64_t seed, const int32_t* node_idx,
                       double* samps_host) {
  bgp_model* m = f->model;
  const int p = f->p, K = f->K;
  const int ldl = round_up(p, 16);
  BGP_CUDA(cudaSetDevice(m->device));
  for (int64_t s = 0; s < M; ++s)
    if (node_idx[s] < 0 || node_idx[s] >= K) {
      set_error("bgp_sample: node index %d out of range at sample %lld", node_idx[s], (long long)s);
      return BGP_ERR_ARG;
    }
  // group the samples by node (stable)
  std::vector<int64_t> count(K, 0), offset(K + 1, 0);
  for (int64_t s = 0; s < M; ++s) count[node_idx[s]]++;
  for (int j = 0; j < K; ++j) offset[j + 1] = offset[j] + count[j];
  std::vector<int32_t> perm((size_t)M);
  {
    std::vector<int64_t> cur(offset.begin(), offset.end() - 1);
    for (int64_t s = 0; s < M; ++s) perm[(size_t)cur[node_idx[s]]++] = (int32_t)s;
  }
  if (f->samps_dev && f->samps_M != M) {
    cudaFree(f->samps_dev);
    f->samps_dev = nullptr;
  }
  if (!f->samps_dev) BGP_CUDA(cudaMalloc(&f->samps_dev, (size_t)p * M * sizeof(double)));
  f->samps_M = M;
  if (!f->Linv_dev) {
    BGP_CUDA(cudaMalloc(&f->Linv_dev, (size_t)p * ldl * sizeof(double)));
    BGP_CUDA(cudaMalloc(&f->LinvT_dev, (size_t)p * ldl * sizeof(double)));
    BGP_CUDA(cudaMalloc(&f->mode_dev, (size_t)ldl * sizeof(double)));
    BGP_CUDA(cudaMemsetAsync(f->Linv_dev, 0, (size_t)p * ldl * sizeof(double), m->stream));
    BGP_CUDA(cudaMemsetAsync(f->LinvT_dev, 0, (size_t)p * ldl * sizeof(double), m->stream));
  }
  double *Zg = nullptr, *Zraw = nullptr, *flag_dev = nullptr;
  int32_t* perm_dev = nullptr;
  int st = [&]() -> int {
    BGP_CUDA(cudaMalloc(&Zg, (size_t)M * ldl * sizeof(double)));
    BGP_CUDA(cudaMalloc(&perm_dev, (size_t)M * sizeof(int32_t)));
    BGP_CUDA(cudaMemcpyAsync(perm_dev, perm.data(), (size_t)M * sizeof(int32_t), cudaMemcpyHostToDevice, m->stream));
    if (Z_host) {
      BGP_CUDA(cudaMalloc(&Zraw, (size_t)p * M * sizeof(double)));
      BGP_CUDA(cudaMemcpyAsync(Zraw, Z_host, (size_t)p * M * sizeof(double), cudaMemcpyHostToDevice, m->stream));
      gather_z_kernel<<<(unsigned)M, 128, 0, m->stream>>>(Zraw, p, M, perm_dev, Zg, ldl);
    } else {
      normal_z_kernel<<<(unsigned)M, 128, 0, m->stream>>>(seed, p, M, perm_dev, Zg, ldl);
    }
    count_launch();
    BGP_CUDA(cudaGetLastError());
    BGP_CUDA(cudaMalloc(&flag_dev, 2 * sizeof(double)));
    BGP_CUDA(cudaMemsetAsync(flag_dev, 0, 2 * sizeof(double), m->stream));
    const bool sharded = m->node_world > 1;
    // node shards: every rank draws the samples of the nodes it holds into its columns of a zeroed p x M matrix;
    // the SUM over the node group assembles the whole (one owner per column, zeros elsewhere => bit-exact)
    if (sharded) BGP_CUDA(cudaMemsetAsync(f->samps_dev, 0, (size_t)p * M * sizeof(double), m->stream));
    for (int j = 0; j < K; ++j) {
      if (count[j] == 0 || f->slot[j] < 0) continue;
      // R_j = chol(forceSymmetric(H_j)): the lower factor of the symmetric H_j, transposed.  H_j and the mode are
      // where the grid evaluation left them on this device; the factor must be that of H in the W order of
      // src/BayesGP.cpp:76-127 (the same z gives a different W under any other ordering), so both are rotated
      BGP_TRY(rot_H_dev(m, f->Hs_dev + (size_t)f->slot[j] * p * m->ldh, m->H, m->ldh));
      BGP_TRY(rot_vec_dev(m, f->modes_dev + (size_t)f->slot[j] * m->lda, f->mode_dev));
      m->L_holds_H = false;
      BGP_TRY(launch_chol_solve(m, false));
      note_chol_info_kernel<<<1, 1, 0, m->stream>>>(m->sc_dev, j, flag_dev);
      count_launch();
      BGP_TRY(launch_trtri(m, f->Linv_dev, ldl, f->LinvT_dev));
      BGP_TRY(launch_kgemm(f->LinvT_dev, p, ldl, Zg + (size_t)offset[j] * ldl, count[j], ldl, p, f->mode_dev,
                           f->samps_dev, p, true, perm_dev + offset[j], m->stream));
    }
    if (sharded) {
      BGP_TRY(node_allreduce_sum(m, f->samps_dev, (size_t)p * M));
      BGP_TRY(node_allreduce_sum(m, flag_dev, 2));
    }
    double flag[2] = {0.0, 0.0};
    BGP_CUDA(cudaMemcpyAsync(flag, flag_dev, sizeof(flag), cudaMemcpyDeviceToHost, m->stream));
    if (samps_host)
      BGP_CUDA(cudaMemcpyAsync(samps_host, f->samps_dev, (size_t)p * M * sizeof(double), cudaMemcpyDeviceToHost,
                               m->stream));
    BGP_CUDA(cudaStreamSynchronize(m->stream));
    if (flag[0] != 0.0) {
      // aghq::sample_marginal stops in chol() here; garbage samples must not leave with BGP_OK
      set_error("Hessian of quadrature node %d is not positive definite (pivot %d): cannot sample", (int)flag[0] - 1,
                (int)flag[1]);
      return BGP_ERR_NOT_PD;
    }
    return BGP_OK;
  }();
  if (Zg) cudaFree(Zg);
  if (Zraw) cudaFree(Zraw);
  if (flag_dev) cudaFree(flag_dev);
  if (perm_dev) cudaFree(perm_dev);
  return st;
}

}  // namespace bgp

using namespace bgp;

extern "C" {

int bgp_sample(bgp_fit* f, int64_t M, const double* Z, const int32_t* node_idx, double* samps) {
  if (!f || !f->model || M <= 0 || !Z || !node_idx) {
    set_error("bgp_sample: bad arguments");
    return BGP_ERR_ARG;
  }
  return sample_core(f, M, Z, 0, node_idx, samps);
}

int bgp_sample_draw(bgp_fit* f, int64_t M, uint64_t seed, double* samps, int32_t* node_idx_out) {
  if (!f || !f->model || M <= 0) {
    set_error("bgp_sample_draw: bad arguments");
    return BGP_ERR_ARG;
  }
  // node ~ Multinomial(lambda), lambda_j = w_j exp(logpost_normalized_j); splitmix64 uniforms on the host
  std::vector<double> cum(f->K);
  double tot = 0.0;
  for (int j = 0; j < f->K; ++j) {
    const double lam = f->weights[j] * std::exp(f->logpost_norm[j]);
    tot += std::isfinite(lam) ? lam : 0.0;
    cum[j] = tot;
  }
  std::vector<int32_t> idx((size_t)M);
  uint64_t x = seed ^ 0x9E3779B97F4A7C15ull;
  for (int64_t s = 0; s < M; ++s) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t zz = x;
    zz = (zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9ull;
    zz = (zz ^ (zz >> 27)) * 0x94D049BB133111EBull;
    zz ^= zz >> 31;
    const double u = ((double)(zz >> 11) + 0.5) * (1.0 / 9007199254740992.0) * tot;
    int j = (int)(std::lower_bound(cum.begin(), cum.end(), u) - cum.begin());
    if (j >= f->K) j = f->K - 1;
    idx[(size_t)s] = j;
  }
  if (node_idx_out) std::copy(idx.begin(), idx.end(), node_idx_out);
  return sample_core(f, M, nullptr, seed, idx.data(), samps);
}

}  // extern "C"
