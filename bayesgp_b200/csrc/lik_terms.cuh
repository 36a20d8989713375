// Per-observation likelihood pieces shared by the streaming pass over the dense design (lik.cu) and the
// O-spline moment pass (ospline.cu):  /root/reference/src/BayesGP.cpp:155-168 and the derivatives TMB's AD
// sweep takes of them (r = d ll / d eta, w = -d2 ll / d eta2, c3 = d w / d eta).
#pragma once

#include "../../include/bgp.h"

namespace bgp {

// per-observation likelihood pieces; all lanes of the warp compute the same values
__device__ __forceinline__ void obs_terms(int family, double tau, double eta, double yv, double sz, double& ll,
                                          double& sumsq, double& r, double& w, double& c3) {
  if (family == BGP_FAMILY_POISSON) {                 // dpois(y, exp(eta), log): y*eta - exp(eta) - lgamma(y+1)
    const double mu = exp(eta);
    ll += yv * eta - mu;
    r = yv - mu;
    w = mu;
    c3 = mu;
  } else if (family == BGP_FAMILY_BINOMIAL) {         // dbinom_robust(y, size, eta, log)
    const double e = exp(-fabs(eta));
    const double l1p = log1p(e);
    const double lse_pos = fmax(eta, 0.0) + l1p;      // log(1 + e^eta)
    const double lse_neg = fmax(-eta, 0.0) + l1p;     // log(1 + e^-eta)
    ll += -yv * lse_neg - (sz - yv) * lse_pos;
    const double inv = 1.0 / (1.0 + e);
    const double pi = eta >= 0.0 ? inv : e * inv;
    const double om = eta >= 0.0 ? e * inv : inv;     // 1 - pi
    w = sz * pi * om;
    r = yv - sz * pi;
    c3 = w * (om - pi);
  } else if (family == BGP_FAMILY_GAUSSIAN) {         // dnorm(y, eta, exp(-theta_S/2), log)
    const double res = yv - eta;
    sumsq += res * res;
    r = tau * res;
    w = tau;
    c3 = 0.0;
  } else {
    r = 0.0;
    w = 0.0;
    c3 = 0.0;
  }
}

}  // namespace bgp
