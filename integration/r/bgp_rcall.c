/*
 * bgp_rcall.c — R .Call glue over the libbgp C ABI (include/bgp.h).
 *
 * UNTESTED IN THIS REPOSITORY: the build image has no R (no Rinternals.h / libR), so this file is
 * not compiled by __graft_entry__.build().  It is the binding a BayesGP maintainer would add next
 * to src/BayesGP.cpp; it contains no arithmetic, only SEXP <-> pointer marshalling.
 * Build on a machine with R:  R CMD SHLIB bgp_rcall.c -I<repo>/include -L<repo>/bayesgp_b200 -lbgp
 *
 * Replaces the TMB-generated entry points registered by `#define TMB_LIB_INIT R_init_BayesGP`
 * (/root/reference/src/BayesGP.cpp:1-2: MakeADFunObject, EvalADFunObject, ...) at the closure level:
 *   bgpR_model(tmbdat)              <- TMB::MakeADFun(data = tmbdat, ..., random = "W")   R/02_model_fit.R:276-282
 *   bgpR_fn / bgpR_gr               <- ff$fn / ff$gr                                        (used inside aghq)
 *   bgpR_fit(model, k, theta0)      <- aghq::marginal_laplace_tmb(ff, k, startingvalue)     R/02_model_fit.R:284
 *   bgpR_sample(fit, Z, node_idx)   <- aghq::sample_marginal(mod, M)                        R/02_model_fit.R:687-689
 *   bgpR_predict_iwp(...)           <- compute_post_fun_IWP + extract_mean_interval_...     R/03_post_fit.R:200-241,287-296
 * Errors: a non-zero status becomes Rf_error(bgp_last_error()) — the longjmp happens on the R side of
 * the boundary, never inside libbgp.  Inner-Newton failures return NaN (TMB behaviour) instead.
 */
#include <R.h>
#include <Rinternals.h>
#include <string.h>

#include "bgp.h"

static void chk(int st) {
  if (st != BGP_OK) Rf_error("libbgp: %s", bgp_last_error());
}

static SEXP list_get(SEXP lst, const char* name) {
  SEXP nm = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(nm, i)), name) == 0) return VECTOR_ELT(lst, i);
  return R_NilValue;
}

static void model_finalizer(SEXP ptr) {
  bgp_model* m = (bgp_model*)R_ExternalPtrAddr(ptr);
  if (m) bgp_model_destroy(m);
  R_ClearExternalPtr(ptr);
}
static void fit_finalizer(SEXP ptr) {
  bgp_fit* f = (bgp_fit*)R_ExternalPtrAddr(ptr);
  if (f) bgp_fit_destroy(f);
  R_ClearExternalPtr(ptr);
}

/* tmbdat as built at R/02_model_fit.R:152-173, with the design blocks left DENSE (as.matrix), i.e. before
 * the dgTMatrix conversion of R/01_utility.R:484-488.  Lists: X, B, P, logPdet, u, alpha, betaprec,
 * betamean, Xf, beta_fixed_prec, beta_fixed_mean; vectors y, size; scalar family_type. */
SEXP bgpR_model(SEXP tmbdat, SEXP device) {
  SEXP y = list_get(tmbdat, "y"), size = list_get(tmbdat, "size");
  const int family = (int)Rf_asReal(list_get(tmbdat, "family_type"));
  bgp_model* m = NULL;
  chk(bgp_model_new((int64_t)XLENGTH(y), family, REAL(y), Rf_isNull(size) ? NULL : REAL(size), Rf_asInteger(device), &m));
  SEXP ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, model_finalizer, TRUE);
  SEXP B = list_get(tmbdat, "B"), P = list_get(tmbdat, "P"), lpd = list_get(tmbdat, "logPdet");
  SEXP u = list_get(tmbdat, "u"), alpha = list_get(tmbdat, "alpha");
  const R_xlen_t J = XLENGTH(B);
  for (R_xlen_t j = 0; j < J; ++j) {
    SEXP Bj = VECTOR_ELT(B, j), Pj = VECTOR_ELT(P, j);
    chk(bgp_model_add_random(m, Rf_ncols(Bj), REAL(Bj), REAL(Pj), 0, Rf_asReal(VECTOR_ELT(lpd, j)),
                             Rf_asReal(VECTOR_ELT(u, j)), Rf_asReal(VECTOR_ELT(alpha, j))));
  }
  SEXP X = list_get(tmbdat, "X"), bp = list_get(tmbdat, "betaprec"), bm = list_get(tmbdat, "betamean");
  for (R_xlen_t j = 0; j < XLENGTH(X); ++j) {
    SEXP Xj = VECTOR_ELT(X, j);
    const int nc = Rf_isMatrix(Xj) ? Rf_ncols(Xj) : 0;
    chk(bgp_model_add_boundary(m, nc, nc ? REAL(Xj) : NULL, Rf_asReal(VECTOR_ELT(bp, j)), Rf_asReal(VECTOR_ELT(bm, j))));
  }
  SEXP Xf = list_get(tmbdat, "Xf"), fp = list_get(tmbdat, "beta_fixed_prec"), fm = list_get(tmbdat, "beta_fixed_mean");
  for (R_xlen_t j = 0; j < XLENGTH(Xf); ++j) {
    SEXP Xj = VECTOR_ELT(Xf, j);
    chk(bgp_model_add_fixed(m, Rf_isMatrix(Xj) ? Rf_ncols(Xj) : 1, REAL(Xj), Rf_asReal(VECTOR_ELT(fp, j)),
                            Rf_asReal(VECTOR_ELT(fm, j))));
  }
  if (family == BGP_FAMILY_GAUSSIAN)   /* (u, alpha) of the noise theta are appended last, R/02_model_fit.R:120-121 */
    chk(bgp_model_set_noise_prior(m, Rf_asReal(VECTOR_ELT(u, XLENGTH(u) - 1)), Rf_asReal(VECTOR_ELT(alpha, XLENGTH(alpha) - 1))));
  chk(bgp_model_finalize(m));
  UNPROTECT(1);
  return ptr;
}

/* ff$fn(theta): value (NaN on inner failure); attribute-free numeric(1).  last.par / spHess via want_* */
SEXP bgpR_eval(SEXP ptr, SEXP theta, SEXP want_grad, SEXP want_mode, SEXP want_hess) {
  bgp_model* m = (bgp_model*)R_ExternalPtrAddr(ptr);
  int p, S;
  chk(bgp_model_dims(m, NULL, &p, &S));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SEXP val = PROTECT(Rf_allocVector(REALSXP, 1));
  SEXP g = Rf_asLogical(want_grad) ? PROTECT(Rf_allocVector(REALSXP, S)) : PROTECT(R_NilValue);
  SEXP w = Rf_asLogical(want_mode) ? PROTECT(Rf_allocVector(REALSXP, p)) : PROTECT(R_NilValue);
  SEXP H = Rf_asLogical(want_hess) ? PROTECT(Rf_allocMatrix(REALSXP, p, p)) : PROTECT(R_NilValue);
  int iters = 0;
  const int st = bgp_laplace_eval(m, REAL(theta), REAL(val), Rf_isNull(g) ? NULL : REAL(g), Rf_isNull(w) ? NULL : REAL(w),
                                  Rf_isNull(H) ? NULL : REAL(H), &iters);
  if (st == BGP_ERR_NOT_PD || st == BGP_ERR_NONFINITE || st == BGP_ERR_NO_CONVERGENCE) {
    REAL(val)[0] = R_NaN;                 /* TMB: NaN + warning, optim's line search backtracks */
    Rf_warning("libbgp inner problem: %s", bgp_last_error());
  } else {
    chk(st);
  }
  SET_VECTOR_ELT(out, 0, val);
  SET_VECTOR_ELT(out, 1, g);
  SET_VECTOR_ELT(out, 2, w);
  SET_VECTOR_ELT(out, 3, H);
  UNPROTECT(5);
  return out;
}

/* aghq::marginal_laplace_tmb(ff, k, startingvalue): returns an external pointer; getters below */
SEXP bgpR_fit(SEXP ptr, SEXP k, SEXP theta0) {
  bgp_model* m = (bgp_model*)R_ExternalPtrAddr(ptr);
  bgp_fit* f = NULL;
  chk(bgp_aghq_fit(m, Rf_asInteger(k), REAL(theta0), &f));
  SEXP out = PROTECT(R_MakeExternalPtr(f, R_NilValue, ptr));   /* keeps the model alive */
  R_RegisterCFinalizerEx(out, fit_finalizer, TRUE);
  UNPROTECT(1);
  return out;
}

/* list(mode, hessian, convergence, nodes, weights, logpost, logpost_normalized, lognormconst, modes, Hs,
 *      marginals = list(data.frame columns...)) — reshaped by R/bgp_shim.R into the aghq object layout */
SEXP bgpR_fit_get(SEXP fptr) {
  bgp_fit* f = (bgp_fit*)R_ExternalPtrAddr(fptr);
  int S, K, p, k;
  chk(bgp_fit_dims(f, &S, &K, &p, &k));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 11));
  SEXP mode = PROTECT(Rf_allocVector(REALSXP, S)), hess = PROTECT(Rf_allocMatrix(REALSXP, S, S));
  SEXP conv = PROTECT(Rf_allocVector(INTSXP, 1));
  int nfn, ngr;
  chk(bgp_fit_get_opt(f, REAL(mode), REAL(hess), INTEGER(conv), &nfn, &ngr));
  SEXP nodes = PROTECT(Rf_allocMatrix(REALSXP, K, S)), wts = PROTECT(Rf_allocVector(REALSXP, K));
  SEXP lp = PROTECT(Rf_allocVector(REALSXP, K)), lpn = PROTECT(Rf_allocVector(REALSXP, K));
  SEXP lnc = PROTECT(Rf_allocVector(REALSXP, 1));
  chk(bgp_fit_get_grid(f, REAL(nodes), REAL(wts), REAL(lp), REAL(lpn), REAL(lnc)));
  SEXP modes = PROTECT(Rf_allocMatrix(REALSXP, p, K));
  SEXP Hs = PROTECT(Rf_alloc3DArray(REALSXP, p, p, K));
  chk(bgp_fit_get_modes(f, REAL(modes), REAL(Hs)));
  SEXP marg = PROTECT(Rf_allocVector(VECSXP, S));
  for (int j = 0; j < S; ++j) {
    SEXP mj = PROTECT(Rf_allocMatrix(REALSXP, k, 3));   /* columns: theta_j, logmargpost, w */
    chk(bgp_fit_get_marginal(f, j, REAL(mj), REAL(mj) + k, REAL(mj) + 2 * k));
    SET_VECTOR_ELT(marg, j, mj);
    UNPROTECT(1);
  }
  SEXP parts[11] = {mode, hess, conv, nodes, wts, lp, lpn, lnc, modes, Hs, marg};
  for (int i = 0; i < 11; ++i) SET_VECTOR_ELT(out, i, parts[i]);
  UNPROTECT(12);
  return out;
}

/* samps$samps (p x M) from explicit draws: Z p x M standard normal, node_idx 0-based integer(M) */
SEXP bgpR_sample(SEXP fptr, SEXP Z, SEXP node_idx) {
  bgp_fit* f = (bgp_fit*)R_ExternalPtrAddr(fptr);
  const R_xlen_t M = XLENGTH(node_idx);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, Rf_nrows(Z), (int)M));
  chk(bgp_sample(f, (int64_t)M, REAL(Z), (const int32_t*)INTEGER(node_idx), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* compute_post_fun_IWP + extract_mean_interval_given_samps: G x 3 matrix (plower, pupper, mean) */
SEXP bgpR_predict_iwp(SEXP coef, SEXP global, SEXP icpt, SEXP knots, SEXP order, SEXP degree, SEXP x, SEXP level,
                      SEXP device) {
  const R_xlen_t G = XLENGTH(x);
  const int64_t M = Rf_ncols(coef);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)G, 3));
  chk(bgp_predict_iwp(REAL(coef), Rf_isNull(global) ? NULL : REAL(global), Rf_isNull(icpt) ? NULL : REAL(icpt), M,
                      REAL(knots), (int)XLENGTH(knots), Rf_asInteger(order), Rf_asInteger(degree), REAL(x), (int64_t)G,
                      Rf_asReal(level), Rf_asInteger(device), REAL(out) + 2 * G, REAL(out), REAL(out) + G, NULL));
  UNPROTECT(1);
  return out;
}

/* predict.FitResult on the samples bgp_sample left on the device (R/03_post_fit.R:65-76 select the same rows of
 * samps$samps): rows are 0-based first rows of the term's blocks, -1 = absent */
SEXP bgpR_fit_predict_iwp(SEXP fptr, SEXP rows, SEXP knots, SEXP order, SEXP degree, SEXP x, SEXP level) {
  bgp_fit* f = (bgp_fit*)R_ExternalPtrAddr(fptr);
  const R_xlen_t G = XLENGTH(x);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)G, 3));      /* mean | plower | pupper */
  chk(bgp_fit_predict_iwp(f, INTEGER(rows)[0], INTEGER(rows)[1], INTEGER(rows)[2], REAL(knots), (int)XLENGTH(knots),
                          Rf_asInteger(order), Rf_asInteger(degree), REAL(x), (int64_t)G, Rf_asReal(level), REAL(out),
                          REAL(out) + G, REAL(out) + 2 * G));
  UNPROTECT(1);
  return out;
}

/* Compute_Q_sB (R/01_utility.R:67-174) on the device: list(P = d x d, logPdet) */
SEXP bgpR_sgp_precision(SEXP a, SEXP k, SEXP m, SEXP region, SEXP accuracy, SEXP device) {
  const int d = 3 * (Rf_asInteger(k) - 2) * Rf_asInteger(m);
  SEXP P = PROTECT(Rf_allocMatrix(REALSXP, d, d));
  SEXP ld = PROTECT(Rf_allocVector(REALSXP, 1));
  chk(bgp_sgp_precision(Rf_asReal(a), Rf_asInteger(k), Rf_asInteger(m), REAL(region), Rf_asReal(accuracy),
                        Rf_asInteger(device), REAL(P), REAL(ld)));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, P);
  SET_VECTOR_ELT(out, 1, ld);
  UNPROTECT(3);
  return out;
}

/* one process per GPU (e.g. Rmpi / callr workers): join the node group whose ranks split quadrature nodes, sample
 * blocks and prediction rows; `id` is the 128-byte raw vector of bgp_nccl_unique_id() made by rank 0 */
SEXP bgpR_nccl_unique_id(void) {
  SEXP id = PROTECT(Rf_allocVector(RAWSXP, 128));
  chk(bgp_nccl_unique_id(RAW(id)));
  UNPROTECT(1);
  return id;
}
SEXP bgpR_set_node_group(SEXP ptr, SEXP rank, SEXP world, SEXP id) {
  chk(bgp_model_set_node_group((bgp_model*)R_ExternalPtrAddr(ptr), Rf_asInteger(rank), Rf_asInteger(world), RAW(id)));
  return R_NilValue;
}

static const R_CallMethodDef callMethods[] = {
    {"bgpR_fit_predict_iwp", (DL_FUNC)&bgpR_fit_predict_iwp, 7}, {"bgpR_sgp_precision", (DL_FUNC)&bgpR_sgp_precision, 6},
    {"bgpR_nccl_unique_id", (DL_FUNC)&bgpR_nccl_unique_id, 0},   {"bgpR_set_node_group", (DL_FUNC)&bgpR_set_node_group, 4},
    {"bgpR_model", (DL_FUNC)&bgpR_model, 2},       {"bgpR_eval", (DL_FUNC)&bgpR_eval, 5},
    {"bgpR_fit", (DL_FUNC)&bgpR_fit, 3},           {"bgpR_fit_get", (DL_FUNC)&bgpR_fit_get, 1},
    {"bgpR_sample", (DL_FUNC)&bgpR_sample, 3},     {"bgpR_predict_iwp", (DL_FUNC)&bgpR_predict_iwp, 9},
    {NULL, NULL, 0}};

void R_init_bgpshim(DllInfo* dll) {
  R_registerRoutines(dll, NULL, callMethods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
