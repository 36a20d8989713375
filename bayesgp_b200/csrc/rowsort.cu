// rowsort.cu — observation re-ordering by zero pattern and the {chunk x column-box} occupancy map.
//
// The reference stores the design blocks as sparse matrices (dgTMatrix built from dense R matrices,
// /root/reference/R/01_utility.R:484-488; tmbdat lists at /root/reference/R/02_model_fit.R:152-173) and TMB's
// sparse Hessian never touches their structural zeros.  O-spline columns are zero for x <= knot_i
// (/root/reference/R/01_utility.R:346-364) and cubic-B-spline (sGP) columns have local support, so about half
// of a typical design matrix is structurally zero.  Here A stays dense and TMA-friendly, and the zeros are
// skipped at the granularity of {64 observations} x {16 columns}:
//   1. every row gets a 64-bit key, bit b set iff the row has a non-zero in columns 16b .. 16b+15;
//   2. rows are radix-sorted by key (stable => deterministic), A / y / size are gathered into that order
//      (f, g and H are sums over observations, so the order is invisible outside the library);
//   3. occ[c] = OR of the keys of rows 64c .. 64c+63.
// The Hessian kernel (syrk.cu) and the likelihood pass (lik.cu) consult occ[] and skip empty cells.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "bgp_internal.h"

namespace bgp {

// one warp per row: lane l tests boxes l and l + 32
__global__ void __launch_bounds__(256) rowkey_kernel(const double* __restrict__ A, int64_t n, int lda,
                                                     unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const double* rp = A + row * (int64_t)lda;
  const int nbox = lda / 16;
  unsigned lo = 0, hi = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int b = lane + 32 * h;
    bool nz = false;
    if (b < nbox) {
      const double2* q = reinterpret_cast<const double2*>(rp + 16 * b);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double2 v = q[i];
        nz |= (v.x != 0.0) | (v.y != 0.0);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, nz);
    if (h == 0) lo = bal; else hi = bal;
  }
  if (lane == 0) {
    keys[row] = ((unsigned long long)hi << 32) | lo;
    idx[row] = (uint32_t)row;
  }
}

// dst row i = src row perm[i]; one warp per row, 16-byte accesses
__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ src, const uint32_t* __restrict__ perm,
                                                          int64_t n, int lda, double* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const double2* s = reinterpret_cast<const double2*>(src + (int64_t)perm[row] * lda);
  double2* d = reinterpret_cast<double2*>(dst + row * (int64_t)lda);
  for (int c = lane; c < lda / 2; c += 32) d[c] = s[c];
}

__global__ void gather_vec_kernel(const double* __restrict__ src, const uint32_t* __restrict__ perm, int64_t n,
                                  double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

__global__ void chunk_occ_kernel(const unsigned long long* __restrict__ keys, int64_t n, int64_t nchunks,
                                 unsigned long long* __restrict__ occ) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  unsigned long long o = 0;
  const int64_t r0 = c * 64, r1 = r0 + 64 < n ? r0 + 64 : n;
  for (int64_t r = r0; r < r1; ++r) o |= keys[r];
  occ[c] = o;
}

int build_row_order(bgp_model* m) {
  const int64_t n = m->n;
  if (n >= ((int64_t)1 << 32)) {
    set_error("more than 2^32 observations per device are not supported");
    return BGP_ERR_ARG;
  }
  bool do_sort = true;
  if (const char* e = getenv("BGP_NO_SORT")) do_sort = !(e[0] == '1');   // diagnostics: keep the caller's row order
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  uint32_t *idx = nullptr, *idx2 = nullptr;
  void* tmp = nullptr;
  int st = [&]() -> int {
    BGP_CUDA(cudaMalloc(&keys, n * sizeof(unsigned long long)));
    BGP_CUDA(cudaMalloc(&idx, n * sizeof(uint32_t)));
    const unsigned blocks = (unsigned)((n + 7) / 8);
    rowkey_kernel<<<blocks, 256, 0, m->stream>>>(m->A, n, m->lda, keys, idx);
    count_launch();
    BGP_CUDA(cudaGetLastError());
    unsigned long long* sorted_keys = keys;
    if (do_sort) {
      BGP_CUDA(cudaMalloc(&keys2, n * sizeof(unsigned long long)));
      BGP_CUDA(cudaMalloc(&idx2, n * sizeof(uint32_t)));
      size_t tb = 0;
      const int end_bit = std::max(1, m->lda / 16);
      BGP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, idx, idx2, (int64_t)n, 0, end_bit, m->stream));
      BGP_CUDA(cudaMalloc(&tmp, tb));
      BGP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, idx, idx2, (int64_t)n, 0, end_bit, m->stream));
      count_launch(4);
      sorted_keys = keys2;
      // gather A, y, size into the sorted order
      double* A2 = nullptr;
      BGP_CUDA(cudaMalloc(&A2, (size_t)n * m->lda * sizeof(double)));
      gather_rows_kernel<<<blocks, 256, 0, m->stream>>>(m->A, idx2, n, m->lda, A2);
      count_launch();
      BGP_CUDA(cudaGetLastError());
      BGP_CUDA(cudaStreamSynchronize(m->stream));
      cudaFree(m->A);
      m->A = A2;
      double* v2 = nullptr;
      const size_t nb_pad = (size_t)(round_up64(n, 64) + 64) * sizeof(double);
      BGP_CUDA(cudaMalloc(&v2, nb_pad));
      BGP_CUDA(cudaMemsetAsync(v2, 0, nb_pad, m->stream));
      const unsigned vb = (unsigned)((n + 255) / 256);
      gather_vec_kernel<<<vb, 256, 0, m->stream>>>(m->y, idx2, n, v2);
      count_launch();
      BGP_CUDA(cudaStreamSynchronize(m->stream));
      std::swap(m->y, v2);
      if (m->size) {
        gather_vec_kernel<<<vb, 256, 0, m->stream>>>(m->size, idx2, n, v2);
        count_launch();
        BGP_CUDA(cudaStreamSynchronize(m->stream));
        std::swap(m->size, v2);
      }
      cudaFree(v2);
    }
    m->nchunks = (n + 63) / 64;
    BGP_CUDA(cudaMalloc(&m->occ_dev, (size_t)m->nchunks * sizeof(unsigned long long)));
    chunk_occ_kernel<<<(unsigned)((m->nchunks + 255) / 256), 256, 0, m->stream>>>(sorted_keys, n, m->nchunks,
                                                                                  (unsigned long long*)m->occ_dev);
    count_launch();
    BGP_CUDA(cudaGetLastError());
    m->occ_host.resize((size_t)m->nchunks);
    BGP_CUDA(cudaMemcpyAsync(m->occ_host.data(), m->occ_dev, (size_t)m->nchunks * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                             m->stream));
    BGP_CUDA(cudaStreamSynchronize(m->stream));
    return BGP_OK;
  }();
  for (void* p : {(void*)keys, (void*)keys2, (void*)idx, (void*)idx2, tmp})
    if (p) cudaFree(p);
  return st;
}

}  // namespace bgp
