// green.cpp — spatial partition of one B200 into {a few SMs, the rest} with CUDA green contexts.
//
// Why: one Laplace evaluation is a chain  pass -> Hessian -> Cholesky -> pass  and the Cholesky (an 8-CTA cluster, 0.23 ms
// at p = 302) leaves 140 of 148 SMs idle.  Two models driven from two host threads could overlap one's Cholesky with the
// other's Hessian, but both big kernels are persistent and fill every SM (the likelihood pass has a static chunk -> CTA
// map), so without a partition the small kernel either waits or delays the big one.  With the partition the big kernels
// of every model run inside the large part (their grids are sized for its SM count) and every Cholesky inside the
// small part.  Opt-in through BGP_GREEN_SMS=<SMs of the small part> at model creation (diagnostics / experiments).
#include <cuda.h>

#include <map>
#include <mutex>

#include "bgp_internal.h"

namespace bgp {

namespace {
struct Part {
  CUgreenCtx small_ctx = nullptr, big_ctx = nullptr;
  int small_sms = 0, big_sms = 0;
};
std::mutex g_mu;
std::map<std::pair<int, int>, Part> g_parts;

template <typename Fn>
bool entry(const char* name, Fn* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
  *out = (Fn)p;
  return true;
}
}  // namespace

int green_streams(int device, int small_sms, cudaStream_t* big, cudaStream_t* small, int* big_sm_count) {
  typedef CUresult (*DevGet)(CUdevice*, int);
  typedef CUresult (*GetRes)(CUdevice, CUdevResource*, CUdevResourceType);
  typedef CUresult (*Split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  typedef CUresult (*GenDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  typedef CUresult (*GreenCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*GreenStream)(CUstream*, CUgreenCtx, unsigned int, int);
  DevGet dev_get;
  GetRes get_res;
  Split split;
  GenDesc gen_desc;
  GreenCreate green_create;
  GreenStream green_stream;
  if (!entry("cuDeviceGet", &dev_get) || !entry("cuDeviceGetDevResource", &get_res) ||
      !entry("cuDevSmResourceSplitByCount", &split) || !entry("cuDevResourceGenerateDesc", &gen_desc) ||
      !entry("cuGreenCtxCreate", &green_create) || !entry("cuGreenCtxStreamCreate", &green_stream)) {
    set_error("green contexts are not available in this driver");
    return BGP_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  Part& part = g_parts[{device, small_sms}];
  if (!part.small_ctx) {
    BGP_CUDA(cudaSetDevice(device));
    BGP_CUDA(cudaFree(0));                       // primary context
    CUdevice dev;
    CUdevResource all, grp, rest;
    unsigned int nb = 1;
    CUdevResourceDesc d_small, d_big;
    if (dev_get(&dev, device) != CUDA_SUCCESS || get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS ||
        split(&grp, &nb, &all, &rest, 0, (unsigned)small_sms) != CUDA_SUCCESS || nb != 1 ||
        gen_desc(&d_small, &grp, 1) != CUDA_SUCCESS || gen_desc(&d_big, &rest, 1) != CUDA_SUCCESS ||
        green_create(&part.small_ctx, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS ||
        green_create(&part.big_ctx, d_big, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
      part = Part();
      set_error("could not split device %d into %d SMs + the rest (green contexts)", device, small_sms);
      return BGP_ERR_CUDA;
    }
    part.small_sms = (int)grp.sm.smCount;
    part.big_sms = (int)rest.sm.smCount;
    if (getenv("BGP_GREEN_DEBUG")) fprintf(stderr, "[green] device %d: %d + %d SMs\n", device, part.small_sms, part.big_sms);
  }
  CUstream sb = nullptr, ss = nullptr;
  if (green_stream(&sb, part.big_ctx, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS ||
      green_stream(&ss, part.small_ctx, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
    set_error("cuGreenCtxStreamCreate failed");
    return BGP_ERR_CUDA;
  }
  *big = (cudaStream_t)sb;
  *small = (cudaStream_t)ss;
  *big_sm_count = part.big_sms;
  return BGP_OK;
}

}  // namespace bgp
