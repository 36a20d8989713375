// io.cu — rotation between the external (ABI) parameter order and the internal column order.
//
// External: W = [U_1..U_J | beta_1..beta_J | beta_fixed..]  (/root/reference/src/BayesGP.cpp:76-127; the order of
// the R index maps, /root/reference/R/02_model_fit.R:644-675).  Internal: dense blocks first (bgp_internal.h).
// Every vector / matrix that crosses include/bgp.h goes through these three helpers.
#include "bgp_internal.h"

namespace bgp {

__global__ void rot_in_kernel(const double* __restrict__ ext, int p, int nD, int lda, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lda) return;
  const int nU = p - nD;
  dst[i] = i < p ? ext[i < nD ? i + nU : i - nD] : 0.0;
}

__global__ void rot_out_kernel(const double* __restrict__ src, int p, int nD, double* __restrict__ ext) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p) return;
  const int nU = p - nD;
  ext[e] = src[e < nU ? e + nD : e - nU];
}

__global__ void rot_H_kernel(const double* __restrict__ H, int p, int nD, int ldh, double* __restrict__ ext, int lde) {
  const int er = blockIdx.x * blockDim.x + threadIdx.x, ec = blockIdx.y;
  if (er >= p) return;
  const int nU = p - nD;
  const int ir = er < nU ? er + nD : er - nU, ic = ec < nU ? ec + nD : ec - nU;
  ext[(size_t)ec * lde + er] = H[(size_t)ic * ldh + ir];
}

int copy_vec_in(bgp_model* m, const double* host_ext, double* dev_int) {
  BGP_CUDA(cudaMemcpyAsync(m->xbuf, host_ext, (size_t)m->p * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  rot_in_kernel<<<(m->lda + 255) / 256, 256, 0, m->stream>>>(m->xbuf, m->p, m->nD, m->lda, dev_int);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

int copy_vec_out(bgp_model* m, const double* dev_int, double* host_ext) {
  rot_out_kernel<<<(m->p + 255) / 256, 256, 0, m->stream>>>(dev_int, m->p, m->nD, m->xbuf);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  BGP_CUDA(cudaMemcpyAsync(host_ext, m->xbuf, (size_t)m->p * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  return BGP_OK;
}

int copy_H_out(bgp_model* m, double* host_ext) {
  dim3 grid((m->p + 255) / 256, m->p);
  rot_H_kernel<<<grid, 256, 0, m->stream>>>(m->H, m->p, m->nD, m->ldh, m->xbuf, m->p);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  BGP_CUDA(cudaMemcpyAsync(host_ext, m->xbuf, (size_t)m->p * m->p * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  return BGP_OK;
}

// device-to-device rotations into the external order (fit getters / node-group gathers)
int rot_vec_dev(bgp_model* m, const double* dev_int, double* dev_ext) {
  rot_out_kernel<<<(m->p + 255) / 256, 256, 0, m->stream>>>(dev_int, m->p, m->nD, dev_ext);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

int rot_H_dev(bgp_model* m, const double* H_int, double* dev_ext, int lde) {
  dim3 grid((m->p + 255) / 256, m->p);
  rot_H_kernel<<<grid, 256, 0, m->stream>>>(H_int, m->p, m->nD, m->ldh, dev_ext, lde);
  count_launch();
  BGP_CUDA(cudaGetLastError());
  return BGP_OK;
}

}  // namespace bgp
