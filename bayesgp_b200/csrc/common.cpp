// Error plumbing and process-wide counters for libbgp.
#include <cstdarg>

#include "bgp_internal.h"

namespace bgp {

thread_local std::string g_last_error;
int64_t g_launch_count = 0;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

}  // namespace bgp

extern "C" {

const char* bgp_last_error(void) { return bgp::g_last_error.c_str(); }
int bgp_version(void) { return 100; }
int64_t bgp_kernel_launch_count(void) { return bgp::g_launch_count; }

}  // extern "C"
