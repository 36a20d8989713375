"""bayesgp_b200 — B200-native inner-inference engine for BayesGP models.

Drop-in replacement for the hot path beneath BayesGP's ``get_result_by_method``
(TMB objective, inner Newton/Laplace, AGHQ grid, posterior sampling, predict).
All numerical work runs in ``libbgp.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/bgp.h``); this package is the thin host-side mirror of the reference's R
interface used by the tests and the benchmark.  There is no CPU fallback.
"""
from ._lib import BgpError, load  # noqa: F401
from .objective import LaplaceObjective, TMBData, make_objective  # noqa: F401
from .terms import Term  # noqa: F401
from .api import (AGHQ, FitResult, build_objective, compute_post_fun_IWP, compute_post_fun_sGP,  # noqa: F401
                  marginal_laplace_tmb, model_fit, model_fit_loop, predict, sample_fixed_effect, sample_marginal)
from .post_fit import compute_pdf_and_cdf, compute_quantiles, fmm_spline, integrate_xy, summary, var_density  # noqa: F401

__all__ = ["BgpError", "load", "LaplaceObjective", "TMBData", "make_objective", "Term", "AGHQ", "FitResult",
           "build_objective", "compute_post_fun_IWP", "compute_post_fun_sGP", "marginal_laplace_tmb", "model_fit",
           "model_fit_loop", "predict", "sample_fixed_effect", "sample_marginal", "compute_pdf_and_cdf", "compute_quantiles", "summary",
           "var_density"]
