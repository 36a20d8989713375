// Error plumbing and process-wide counters for libbgp.
#include <cstdarg>

#include <cuda_profiler_api.h>

#include "bgp_internal.h"

namespace bgp {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

// Phase timing without holding up the device: marks are event records in stream order; after a synchronize the list
// is only SEALED (phase_harvest); the elapsed times of a sealed list are read later, while the device is busy with the
// next segment (phase_collect: read_scalars before its synchronize, the timing getters).
void phase_mark(bgp_model* m, int phase) {
  if (m->marks.size() >= m->ev_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    m->ev_pool.push_back(e);
  }
  cudaEventRecord(m->ev_pool[m->marks.size()], m->stream);
  m->marks.push_back(phase);
}

void phase_collect(bgp_model* m) {
  for (size_t i = 0; i + 1 < m->sealed_marks.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, m->sealed_pool[i], m->sealed_pool[i + 1]) != cudaSuccess) continue;
    switch (m->sealed_marks[i]) {
      case PH_LIK: m->t_lik += ms; break;
      case PH_HESS: m->t_hess += ms; break;
      case PH_CHOL: m->t_chol += ms; break;
      case PH_LEV: m->t_lev += ms; break;
      default: break;
    }
  }
  m->sealed_marks.clear();
}

void phase_harvest(bgp_model* m) {
  if (!m->sealed_marks.empty()) phase_collect(m);
  m->sealed_marks.swap(m->marks);       // marks is empty now
  m->sealed_pool.swap(m->ev_pool);      // the other set of events serves the next segment
}

}  // namespace bgp

extern "C" {

const char* bgp_last_error(void) { return bgp::g_last_error.c_str(); }
int bgp_version(void) { return 100; }
int64_t bgp_kernel_launch_count(void) { return bgp::g_launch_count.load(); }
// capture window for `ncu --profile-from-start off` (diagnostics)
int bgp_profiler_range(int start) { return (start ? cudaProfilerStart() : cudaProfilerStop()) == cudaSuccess ? BGP_OK : BGP_ERR_CUDA; }

}  // extern "C"
