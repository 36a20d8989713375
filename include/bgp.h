/*
 * bgp.h — C ABI of libbgp, the B200-native replacement for BayesGP's inner
 * inference hot path (everything below `get_result_by_method`,
 * /root/reference/R/02_model_fit.R:275-285, plus the sample -> function
 * evaluation of /root/reference/R/03_post_fit.R:200-296).
 *
 * Conventions
 *   - every matrix crossing this boundary is FP64, COLUMN-MAJOR (R's layout);
 *   - all pointers are HOST pointers unless the name ends in `_dev`;
 *   - the caller owns every buffer; the library copies inputs to the device;
 *   - every entry point returns an int status (BGP_OK == 0); the message of the
 *     last failure on this thread is available from bgp_last_error();
 *   - no exceptions / longjmp cross the boundary; all calls are blocking;
 *   - one host thread drives one model (same as R's single thread).
 *
 * Each group cites the reference interface it replaces.
 */
#ifndef BGP_H
#define BGP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgp_model bgp_model;   /* replaces the TMB ADFun external pointer (`ff`) */
typedef struct bgp_fit bgp_fit;       /* replaces the `marginallaplace`/`aghq` list (`mod`) */

enum {
  BGP_OK = 0,
  BGP_ERR_ARG = 1,           /* bad argument / wrong call order                         */
  BGP_ERR_CUDA = 2,          /* CUDA runtime / driver failure                           */
  BGP_ERR_NOT_PD = 3,        /* Hessian not positive definite (Cholesky pivot <= 0)     */
  BGP_ERR_NONFINITE = 4,     /* objective non-finite at the starting point              */
  BGP_ERR_NO_CONVERGENCE = 5,/* inner Newton hit maxit (value returned as NaN)          */
  BGP_ERR_NCCL = 6,          /* NCCL failure / NCCL not loadable                        */
  BGP_ERR_STATE = 7          /* handle not finalized / already finalized                */
};

/* family codes: /root/reference/R/02_model_fit.R:8-28, src/BayesGP.cpp:155-214 */
enum { BGP_FAMILY_GAUSSIAN = 0, BGP_FAMILY_POISSON = 1, BGP_FAMILY_BINOMIAL = 2, BGP_FAMILY_NONE = -2 };

const char* bgp_last_error(void);
int bgp_version(void);
/* number of CUDA kernels this library has launched in this process (bench `gpu_launches`) */
int64_t bgp_kernel_launch_count(void);
/* cudaProfilerStart / Stop: the capture window of `ncu --profile-from-start off` (scripts/ncu_r02.sh) */
int bgp_profiler_range(int start);

/* ------------------------------------------------------------------------------------------
 * Model construction — replaces the `tmbdat` list + TMB::MakeADFun(data, parameters,
 * random = "W", DLL = "BayesGP")  (/root/reference/R/02_model_fit.R:152-183, :276-282;
 * DATA_* macros of /root/reference/src/BayesGP.cpp:34-52).
 * Call order: bgp_model_new -> add_random* -> add_boundary* -> add_fixed* -> finalize.
 * W layout = [U_1..U_J | beta_1..beta_J | beta_fixed_0..]  (src/BayesGP.cpp:76-127).
 * ---------------------------------------------------------------------------------------- */
int bgp_model_new(int64_t n, int family, const double* y, const double* size /* NULL => 1s */, int device,
                  bgp_model** out);
/* one smoothing term: B (n x d), P (d x d dense, or length-d diagonal when p_is_diag), logPdet,
 * PC-prior (u, alpha): tmbdat$B/P/logPdet/u/alpha  (R/02_model_fit.R:64-68) */
int bgp_model_add_random(bgp_model* m, int d, const double* B, const double* P, int p_is_diag, double logPdet,
                         double u, double alpha);
/* boundary block of a term: X (n x ncol), beta ~ N(mean, 1/prec): tmbdat$X/betaprec/betamean (:53-62) */
int bgp_model_add_boundary(bgp_model* m, int ncol, const double* X, double prec, double mean);
/* fixed effect block (first call = intercept): tmbdat$Xf/beta_fixed_prec/beta_fixed_mean (:137-149) */
int bgp_model_add_fixed(bgp_model* m, int ncol, const double* Xf, double prec, double mean);
/* PC prior of the Gaussian noise theta (appended last, R/02_model_fit.R:120-121) */
int bgp_model_set_noise_prior(bgp_model* m, double u, double alpha);
/* GPU-side constructor of an IWP term from the covariate alone: builds B = local_poly_helper(knots,
 * x - x0, order), X = global_poly(x - x0)[,-1], P = diag(diff(knots)) on the device
 * (/root/reference/R/01_utility.R:278-300,325-401; R/02_model_fit.R:460-462). Adds the random AND the
 * boundary block of the term (boundary blocks of generated terms keep the order of the calls). */
int bgp_model_add_iwp(bgp_model* m, const double* x, double initial_location, const double* knots, int nknots,
                      int order, double u, double alpha, double boundary_prec, double boundary_mean);
/* GPU-side constructor of an sGP term from the covariate: B = cbind over harmonics i = 1..m of
 * [B cos(i a x), B sin(i a x), B] with the cubic B-spline basis on `region` minus its first two functions
 * (Compute_B_sB with boundary = TRUE, as the fit always uses), X = cbind(cos(i a x), sin(i a x))
 * (/root/reference/R/01_utility.R:177-195,224-239,301-312; R/02_model_fit.R:493-569).  P (d x d dense,
 * d = 3 (k-2) m, Compute_Q_sB) and logPdet come from the caller.  Adds the random AND the boundary block. */
int bgp_model_add_sgp(bgp_model* m, const double* x, double initial_location, double a, int k, int nharm,
                      const double* region /* 2 */, const double* P, double logPdet, double u, double alpha,
                      double boundary_prec, double boundary_mean);
/* Compute_Q_sB on the device (/root/reference/R/01_utility.R:67-174,255-272): the sGP precision, block diagonal over
 * the harmonics i = 1..m (frequency i a), d = 3 (k - 2) m, from the cubic B-spline basis on `region` minus its first
 * two functions and its first / second derivatives on the grid seq(region[0], region[1], by = accuracy) with weights
 * diff(c(0, x)): one FP64 tensor-pipe Gram product of the nine function families, then the reference's block
 * formulas.  P is d x d column-major; logPdet = determinant(P)$modulus (R/02_model_fit.R:66), may be NULL. */
int bgp_sgp_precision(double a, int k, int nharm, const double* region /* 2 */, double accuracy, int device, double* P,
                      double* logPdet);
/* bgp_model_add_sgp with P and logPdet computed by bgp_sgp_precision on the model's device */
int bgp_model_add_sgp_auto(bgp_model* m, const double* x, double initial_location, double a, int k, int nharm,
                           const double* region /* 2 */, double accuracy, double u, double alpha, double boundary_prec,
                           double boundary_mean);
/* observation sharding: this process holds rows [row0, row0 + n) of a global problem of n_total rows and
 * joins an NCCL communicator of `world` ranks (nccl_unique_id: the 128-byte ncclUniqueId produced by
 * bgp_nccl_unique_id on rank 0 and broadcast by the launcher). Must precede finalize. */
int bgp_nccl_unique_id(void* id128);
int bgp_model_set_shard(bgp_model* m, int rank, int world, const void* nccl_unique_id);
/* node group: the `world` processes of this communicator hold the SAME rows (replicas, or the same observation
 * shard) and split between them the quadrature nodes of bgp_aghq_fit's grids (the K node evaluations of
 * aghq::normalize_logpost, R/02_model_fit.R:284, contiguous runs of the expand.grid order), the per-node sample
 * blocks of bgp_sample* (aghq::sample_marginal, :687-689) and the grid rows of bgp_fit_predict_*
 * (R/03_post_fit.R:235,287-296).  BFGS and the Richardson Hessian run replicated (identical on every rank).
 * Independent of bgp_model_set_shard (both may be set: a 2-D layout); may be called before or after finalize.
 * With a node group every rank must make the same fit / sample / predict / bgp_fit_get_modes calls. */
int bgp_model_set_node_group(bgp_model* m, int rank, int world, const void* nccl_unique_id);
int bgp_model_finalize(bgp_model* m);
void bgp_model_destroy(bgp_model* m);

int bgp_model_dims(const bgp_model* m, int64_t* n, int* p, int* S);

/* ------------------------------------------------------------------------------------------
 * The TMB objective and its W-derivatives at a given (W, theta) — replaces
 * objective_function<Type>::operator() (src/BayesGP.cpp:30-253) and the AD sweeps TMB runs
 * on it. Any output pointer may be NULL. H is p x p column-major (full symmetric).
 * ---------------------------------------------------------------------------------------- */
int bgp_objective(bgp_model* m, const double* W, const double* theta, double* f, double* grad /* p */,
                  double* H /* p*p */);

/* ------------------------------------------------------------------------------------------
 * TMB-style Laplace objective — replaces ff$fn / ff$gr / ff$env$last.par / ff$env$spHess
 * (call site R/02_model_fit.R:276-284; consumed inside aghq::marginal_laplace_tmb).
 *   value  = f(w_hat,theta) + 1/2 logdet H(w_hat,theta) - p/2 log(2 pi)   (NaN on inner failure)
 *   grad   = d value / d theta  (S)            — NULL to skip (costs a leverage pass)
 *   w_mode = w_hat (p)                          — ff$env$last.par[random]
 *   H      = H(w_hat, theta) (p x p)            — ff$env$spHess(last.par, random = TRUE)
 * The inner Newton warm-starts from the previous call's mode (TMB last.par.best behaviour).
 * ---------------------------------------------------------------------------------------- */
int bgp_laplace_eval(bgp_model* m, const double* theta, double* value, double* grad, double* w_mode, double* H,
                     int* newton_iters);
/* K evaluations in one call (theta is S x K column-major: node j = theta + j*S); outputs may be NULL:
 * values[K], modes p x K, Hs p x p x K. */
int bgp_laplace_eval_batch(bgp_model* m, int K, const double* theta, double* values, double* modes, double* Hs,
                           int* newton_iters_total);
/* reset / set the warm start (tmbparams W = 0, R/02_model_fit.R:249-252) */
int bgp_model_set_start(bgp_model* m, const double* W /* NULL => zeros */);
/* The state an optimiser leaves behind at its last evaluation: theta, the mode there (ff$env$last.par) and the
 * tangent d w_hat / d theta (p x S column-major; -H^-1 d2f/dW dtheta, a by-product of the last factorisation, which
 * the warm-start predictor of the next evaluations uses).  bgp_model_get_tangent reads (theta, T) of the most recent
 * evaluation; bgp_model_set_start_at restores such a state (history cleared, then this single entry; T may be NULL:
 * plain warm start as bgp_model_set_start).  bgp_aghq_fit does this implicitly: its grid phase starts from what
 * BFGS / Richardson left. */
int bgp_model_get_tangent(bgp_model* m, double* theta /* S, may be NULL */, double* T /* p x S, may be NULL */);
int bgp_model_set_start_at(bgp_model* m, const double* theta, const double* W, const double* T);
/* inner solver controls (TMB newton(): tol = grad.tol = step.tol = 1e-8, maxit = 100) */
int bgp_model_set_newton(bgp_model* m, double grad_tol, double step_tol, int maxit);

/* ------------------------------------------------------------------------------------------
 * aghq::marginal_laplace_tmb(ff, k, startingvalue)  (R/02_model_fit.R:284): BFGS (vmmin) from
 * theta0, Richardson Hessian of ff$gr (ff$he, :283), Gauss-Hermite product grid, normalisation,
 * marginals ("reuse"), per-node modes and Hessians.
 * ---------------------------------------------------------------------------------------- */
int bgp_aghq_fit(bgp_model* m, int k, const double* theta0, bgp_fit** out);
/* O-spline moment path.  A model whose only smoothing term is an IWP of order <= 4 built by bgp_model_add_iwp, with at
 * most 8 dense (boundary + fixed) columns, evaluates eta, A^T r and A^T diag(w) A from per-knot-interval moments of one
 * streaming pass over (x, y, size, dense columns) instead of the two passes over the dense design: to the right of its
 * own knot interval every O-spline column (R/01_utility.R:346-364) is a polynomial of degree order-1.  Same results to
 * rounding; on by default for eligible models.  on = 0 selects the dense (DMMA) path, on = 2 the moment path with the
 * Laplace gradient's leverages taken from the dense design (both for A/B measurements). */
int bgp_model_set_ospline(bgp_model* m, int on);
int bgp_model_get_ospline(const bgp_model* m, int* eligible, int* on);
/* Evaluation lanes.  A batch of Laplace evaluations on the moment path (bgp_laplace_eval_batch, the quadrature grids of
 * bgp_aghq_fit*) leaves most of the device idle — its dominant kernel, the p x p Cholesky, runs on 8 SMs.  With
 * lanes > 1 the nodes of a batch are dealt to that many evaluation contexts (own stream, host thread, iterate, Hessian,
 * factor, warm-start history and moment buffers; observations and all read-only arrays shared) that run concurrently.
 * Contiguous runs of the node order per lane, at least two nodes each; values differ from a one-lane run only through
 * the starting points of the inner solves.  Dense-path and observation-sharded models always use one lane. */
int bgp_model_set_lanes(bgp_model* m, int lanes);
int bgp_model_get_lanes(const bgp_model* m, int* lanes);
/* algorithmic bytes one likelihood pass of the moment path moves (u, y, eta in / out, dense columns, size): roofline numerator */
int bgp_model_ospline_bytes(const bgp_model* m, double* bytes_per_pass);
/* When numDeriv's Richardson Hessian of ff$gr (d = 1e-4) is not positive definite aghq stops in chol(); so does
 * bgp_aghq_fit (BGP_ERR_NOT_PD).  allow = 1 opts into a retry with d = 1e-3 and 1e-2 (not in the reference); the
 * number of retries a fit needed is reported by bgp_fit_get_diagnostics. */
int bgp_model_set_hessian_retry(bgp_model* m, int allow);
/* same, but with the optimisation results supplied (aghq's `optresults` argument) */
int bgp_aghq_fit_at(bgp_model* m, int k, const double* mode, const double* hessian /* S x S */, bgp_fit** out);
void bgp_fit_destroy(bgp_fit* f);
int bgp_fit_dims(const bgp_fit* f, int* S, int* K, int* p, int* k);
/* getters mirror mod$optresults / mod$normalized_posterior / mod$modesandhessians / mod$marginals.  The per-node
 * modes and Hessians live on the device(s) that evaluated them; bgp_fit_get_modes rotates them into the W order and
 * copies them out (a collective call when the model has a node group). */
int bgp_fit_get_opt(const bgp_fit* f, double* mode /* S */, double* hessian /* S*S */, int* convergence,
                    int* fn_count, int* gr_count);
int bgp_fit_get_grid(const bgp_fit* f, double* nodes /* K x S col-major */, double* weights /* K */,
                     double* logpost /* K */, double* logpost_normalized /* K */, double* lognormconst);
int bgp_fit_get_modes(const bgp_fit* f, double* modes /* p x K */, double* Hs /* p x p x K, may be NULL */);
int bgp_fit_get_marginal(const bgp_fit* f, int j, double* theta /* k */, double* logmargpost /* k */,
                         double* w /* k */);
/* the same arrays without a copy: pointers into the fit's own page-locked host mirror (modes p x K, Hs p x p x K, W
 * order), valid until bgp_fit_destroy; filled while the grid was being evaluated (collective when the model has a
 * node group: the other ranks' nodes are gathered on the first call) */
int bgp_fit_host_arrays(const bgp_fit* f, const double** modes, const double** Hs);
/* how the fit went: Richardson retries used (0 = numDeriv's default step, as the reference), inner Newton iterations
 * spent on the quadrature grids by this rank, wall clock (ms) of the BFGS + Richardson phase and of the grid phase */
int bgp_fit_get_diagnostics(const bgp_fit* f, int* hessian_fallback, int64_t* grid_newton_iters, double* opt_ms,
                            double* grid_ms);
/* node-group rank that evaluated (and holds the mode / Hessian of) each of the K nodes */
int bgp_fit_node_owner(const bgp_fit* f, int32_t* owner /* K */);

/* ------------------------------------------------------------------------------------------
 * aghq::sample_marginal(mod, M)  (R/02_model_fit.R:687-689) with the random inputs explicit:
 * samps[, m] = mode_{node_idx[m]} + chol(H_{node_idx[m]})^{-1} Z[, m];  Z is p x M standard
 * normal, node_idx 0-based. Output p x M (rows = W entries, columns = samples, as samps$samps).
 * ---------------------------------------------------------------------------------------- */
int bgp_sample(bgp_fit* f, int64_t M, const double* Z, const int32_t* node_idx, double* samps);
/* library-side draw of (node_idx, Z) with a counter-based generator, then bgp_sample; the
 * samples stay resident on the device for bgp_predict_* (pass samps = NULL to skip the copy) */
int bgp_sample_draw(bgp_fit* f, int64_t M, uint64_t seed, double* samps, int32_t* node_idx);

/* ------------------------------------------------------------------------------------------
 * Sample -> function evaluation and summary — replaces compute_post_fun_IWP
 * (R/03_post_fit.R:200-241), compute_post_fun_sGP (:261-276) and
 * extract_mean_interval_given_samps (:287-296; stats::quantile type 7).
 *   coef    (k-1) x M   spline coefficient samples      (samps$samps[random idx, ])
 *   global  (order-1) x M boundary coefficient samples  (may be NULL => zeros)
 *   icpt    M           intercept samples               (may be NULL => zeros)
 *   x       G           refined_x (already shifted by initial_location and sorted by the caller)
 * Outputs (any may be NULL): mean/plower/pupper [G]; samples G x M column-major.
 * ---------------------------------------------------------------------------------------- */
int bgp_predict_iwp(const double* coef, const double* global, const double* icpt, int64_t M, const double* knots,
                    int nknots, int order, int degree, const double* x, int64_t G, double level, int device,
                    double* mean, double* plower, double* pupper, double* samples);
int bgp_predict_sgp(const double* coef, const double* global, const double* icpt, int64_t M, double a, int k, int m,
                    const double* region /* 2 */, int boundary, const double* x, int64_t G, double level, int device,
                    double* mean, double* plower, double* pupper, double* samples);
/* The same two evaluations from the samples bgp_sample / bgp_sample_draw left on the device (no host copy of the
 * p x M matrix): coef_row0 / global_row0 / icpt_row are the first 0-based rows of the term's spline block, its
 * boundary block and the intercept in samps$samps (random_samp_indexes / boundary_samp_indexes /
 * fixed_samp_indexes$intercept, R/02_model_fit.R:627-675; -1 = block absent / not included).  With a node group the
 * rows of x are split over the ranks and the three G-vectors are returned whole on every rank. */
int bgp_fit_predict_iwp(bgp_fit* f, int coef_row0, int global_row0, int icpt_row, const double* knots, int nknots, int order,
                        int degree, const double* x, int64_t G, double level, double* mean, double* plower,
                        double* pupper);
int bgp_fit_predict_sgp(bgp_fit* f, int coef_row0, int global_row0, int icpt_row, double a, int k, int m,
                        const double* region /* 2 */, int boundary, const double* x, int64_t G, double level, double* mean,
                        double* plower, double* pupper);
/* device time (CUDA events on the call's stream, ms) of the last bgp_predict_* call on this thread: the FP64
 * tensor-pipe GEMM strips, the per-row quantile selection, and the whole device sequence */
int bgp_predict_last_timing(double* gemm_ms, double* select_ms, double* total_ms);
/* {128-row x 16-column} slices of the prediction design the GEMM of the last bgp_predict_* call executed, and how many
 * there are: the structurally empty ones (x_new is sorted: an O-spline block is a staircase, a B-spline block a
 * band) are neither copied nor multiplied.  Both zero when the map was not used (more than 1024 design columns). */
int bgp_predict_last_occupancy(double* executed_slices, double* total_slices);
/* basis evaluators on the device (R/01_utility.R:378-401, :413-419, :198-208): out is G x ncol col-major */
int bgp_basis_iwp(const double* knots, int nknots, int order, const double* x, int64_t G, int device, double* out);

/* timing of the last call, measured with CUDA events on the model's stream (milliseconds) */
int bgp_model_last_timing(const bgp_model* m, double* total_ms, double* lik_ms, double* hess_ms, double* chol_ms,
                          int64_t* lik_launches, int64_t* hess_launches, int64_t* chol_launches);

/* the leverage kernel of ff$gr (q_i = a_i^T H^-1 a_i, grad.cu): cumulative device time (ms), launches, executed flops
 * per launch (structurally non-zero {128 observations x 64 rows x 16 columns} blocks of the upper-triangular product)
 * and the n p^2 flops of the dense lower-triangular formulation */
int bgp_model_gradient_timing(const bgp_model* m, double* leverage_ms, int64_t* leverage_launches, double* leverage_flops,
                              double* dense_flops);
/* counters since creation: Laplace evaluations, accepted Newton iterations, evaluations whose log-determinant
 * came from the factor of the last Newton iteration (certified: |d logdet| <= p * max|d eta|, see below) */
int bgp_model_counters(const bgp_model* m, int64_t* laplace_evals, int64_t* newton_iters, int64_t* factor_reuses);
/* When the inner Newton converges by a full step w0 -> w1 that moved the linear predictor by delta = max |d eta| with
 * delta <= eta_tol (default 1e-7) and p * delta / 2 <= rel_tol * |value| (default 1e-10), the factor of that
 * last iteration is used for 1/2 logdet H instead of a new Hessian + Cholesky at the mode (TMB recomputes).
 * Exact bounds (newton.cu): e^-delta H(w0) <= H(w1) <= e^delta H(w0), so |d logdet| <= p delta (100x inside the 1e-8
 * tolerance of the value) and the H returned by bgp_laplace_eval* / kept by bgp_aghq_fit* — then H(w0), not H(w1) —
 * is within (e^delta - 1) <= 1e-7 relative (2-norm) of ff$env$spHess at the mode, 10x inside the 1e-6 tolerance.
 * allow = 0 restores the recomputation.  Gradients always re-form H and its factor at the mode, so a call that
 * asks for grad AND H returns H(w1). */
int bgp_model_set_factor_reuse(bgp_model* m, int allow, double eta_tol, double rel_tol);
/* flops of one Hessian launch H = A^T diag(w) A: dense (n p (p+1)) and the structurally non-zero part the
 * kernel executes after skipping empty {64-observation x 16-column} cells (roofline reporting) */
int bgp_model_hessian_flops(const bgp_model* m, double* dense, double* structural);
/* HBM bytes of one likelihood pass: dense (8 n (lda + 3)) and the part actually moved after skipping
 * structurally empty {64-observation x 64-column} groups */
int bgp_model_lik_bytes(const bgp_model* m, double* dense, double* structural);

#ifdef __cplusplus
}
#endif
#endif /* BGP_H */
