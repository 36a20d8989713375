// temporary: entry points not implemented yet
#include "bgp_internal.h"
using namespace bgp;
#define NI(name) set_error(#name ": not implemented yet"); return BGP_ERR_ARG;
extern "C" {
int bgp_aghq_fit(bgp_model*, int, const double*, bgp_fit**) { NI(bgp_aghq_fit) }
int bgp_aghq_fit_at(bgp_model*, int, const double*, const double*, bgp_fit**) { NI(bgp_aghq_fit_at) }
void bgp_fit_destroy(bgp_fit*) {}
int bgp_fit_dims(const bgp_fit*, int*, int*, int*, int*) { NI(bgp_fit_dims) }
int bgp_fit_get_opt(const bgp_fit*, double*, double*, int*, int*, int*) { NI(x) }
int bgp_fit_get_grid(const bgp_fit*, double*, double*, double*, double*, double*) { NI(x) }
int bgp_fit_get_modes(const bgp_fit*, double*, double*) { NI(x) }
int bgp_fit_get_marginal(const bgp_fit*, int, double*, double*, double*) { NI(x) }
int bgp_sample(bgp_fit*, int64_t, const double*, const int32_t*, double*) { NI(x) }
int bgp_sample_draw(bgp_fit*, int64_t, uint64_t, double*, int32_t*) { NI(x) }
int bgp_predict_iwp(const double*, const double*, const double*, int64_t, const double*, int, int, int, const double*, int64_t, double, int, double*, double*, double*, double*) { NI(x) }
int bgp_predict_sgp(const double*, const double*, const double*, int64_t, double, int, int, const double*, int, const double*, int64_t, double, int, double*, double*, double*, double*) { NI(x) }
int bgp_basis_iwp(const double*, int, int, const double*, int64_t, int, double*) { NI(x) }
}
