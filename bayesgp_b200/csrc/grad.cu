// grad.cu — gradient of the Laplace objective w.r.t. theta (placeholder until the leverage kernel lands).
#include "bgp_internal.h"

namespace bgp {

int laplace_gradient(bgp_model* m, const double* theta, double* grad_host) {
  (void)m; (void)theta; (void)grad_host;
  set_error("laplace gradient not implemented yet");
  return BGP_ERR_ARG;
}

}  // namespace bgp
