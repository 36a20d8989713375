// Internal declarations shared by the libbgp translation units.
// Host control flow is C++; every FLOP on the hot path runs in the sm_100a kernels
// declared at the bottom.  There is no CPU fallback: if CUDA is unavailable every entry
// point returns BGP_ERR_CUDA.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cmath>
#include <memory>
#include <cstdio>
#include <cstring>
#include <string>
#include <functional>
#include <vector>

#include "../../include/bgp.h"

namespace bgp {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local std::string g_last_error;

struct Status {
  int code;
  Status(int c = BGP_OK) : code(c) {}
  bool ok() const { return code == BGP_OK; }
};

#define BGP_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::bgp::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,  \
                       cudaGetErrorString(_e));                                                 \
      return BGP_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define BGP_TRY(expr)              \
  do {                             \
    int _s = (expr);               \
    if (_s != BGP_OK) return _s;   \
  } while (0)

extern std::atomic<int64_t> g_launch_count;      // models may be driven from different host threads
inline void count_launch(int64_t k = 1) { g_launch_count.fetch_add(k, std::memory_order_relaxed); }

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- device model ---------------------------------------------------------------------------
struct RandomBlock {
  int d = 0;
  int off = 0;              // first W index of U_j
  bool diag = true;
  double* P_dev = nullptr;  // d (diag) or d*d column-major
  std::vector<double> P_host;
  double logPdet = 0, u = 1, alpha = 0.5;
};

// scalars produced on the device and read by the host once per Newton iteration
struct EvalScalars {
  double f;        // objective f(W, theta)
  double ll;       // log-likelihood part
  double gmax;     // max |g|
  double quad;     // (W-mu0)^T Q (W-mu0)
  double smax;     // max |step| of the last solve
  double logdet;   // log det H of the last factorisation
  int chol_info;   // 0 ok, j+1: non-positive pivot at column j
  int nonfinite;   // 1 if eta / ll produced a non-finite value
  double sumsq;    // Gaussian: sum (y-eta)^2
  double pad;      // max |eta - previous eta| of the last likelihood pass (sharded: sum of the ranks' maxima, an upper bound)
};

struct Comm;   // NCCL communicator wrapper (comm.cpp)

// A parked host thread that runs one job at a time (the lanes of a model; newton.cu posts the jobs).
struct LaneWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, done = true, quit = false;
  LaneWorker() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(mu);
      for (;;) {
        cv.wait(lk, [this] { return has_job || quit; });
        if (quit) return;
        std::function<void()> j = std::move(job);
        has_job = false;
        lk.unlock();
        j();
        lk.lock();
        done = true;
        cv.notify_all();
      }
    });
  }
  void post(std::function<void()> j) {
    std::unique_lock<std::mutex> lk(mu);
    job = std::move(j);
    has_job = true;
    done = false;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [this] { return done; });
  }
  ~LaneWorker() {
    {
      std::unique_lock<std::mutex> lk(mu);
      quit = true;
      cv.notify_all();
    }
    if (th.joinable()) th.join();
  }
};

}  // namespace bgp

struct bgp_model {
  // ---- description -------------------------------------------------------------------------
  int64_t n = 0;        // local rows
  int family = 0;
  int device = 0;
  bool finalized = false;
  int p = 0, S = 0, J = 0;
  int lda = 0;          // row pitch of A in doubles (multiple of 16)
  // Internal column order: the dense blocks first, [X_1..X_J | Xf_0..Xf_F | B_1..B_J], so that the always
  // non-zero columns share column box 0 and the spline blocks keep their zero structure box-aligned.  The
  // external (ABI) order is the W layout of src/BayesGP.cpp:76-127, [B | X | Xf]; vectors and matrices are
  // rotated at the API boundary (io.cu).
  int nD = 0;           // number of dense (boundary + fixed) columns
  std::vector<double> qfix_host;   // theta-independent diagonal of Q, internal order (p)
  double* xbuf = nullptr;          // device scratch for boundary rotations (max(lda, p*p) doubles)
  std::vector<bgp::RandomBlock> rnd;
  std::vector<int> bnd_dim, fix_dim;
  std::vector<double> bnd_prec, bnd_mean, fix_prec, fix_mean;
  std::vector<double> theta_u, theta_alpha;   // length S (noise last)
  double noise_u = 1.0, noise_alpha = 0.5;
  // staging of host blocks before finalize (column-major device copies)
  struct Staged { int ncol; double* dev; };
  std::vector<Staged> st_rnd, st_bnd, st_fix;
  // ---- device state --------------------------------------------------------------------------
  cudaStream_t stream = nullptr;
  double* A = nullptr;          // n x lda row-major (observation-major)
  double* y = nullptr;
  double* size = nullptr;
  double* eta = nullptr;
  double* wobs = nullptr;       // -d2 ll / d eta2, padded to a multiple of 64 with zeros
  double* c3 = nullptr;         // d w / d eta (only filled on request)
  double* qfix = nullptr;       // theta-independent diagonal of Q (p)
  double* mu0 = nullptr;        // prior mean (p)
  double* W = nullptr;          // current iterate (lda)
  double* Wtrial = nullptr;     // trial point (lda)
  double* Wmode = nullptr;      // warm start / last mode (lda)
  double* g = nullptr;          // gradient (lda)
  double* step = nullptr;       // Newton step (lda)
  double* Tan = nullptr;        // S x lda tangent d w_hat / d theta at the last mode (warm-start predictor)
  std::vector<double> theta_last;
  bool tan_valid = false;
  // recent evaluations (theta, mode, tangent) for the warm-start predictor: cubic Hermite through the two
  // nearest entries collinear with the new theta (grid rows, line searches), else first order from the nearest
  static constexpr int NHIST = 8;
  struct Hist {
    std::vector<double> theta;
    double* W = nullptr;      // lda
    double* T = nullptr;      // S x lda
    uint64_t stamp = 0;       // 0 = empty
  };
  Hist hist[NHIST];
  uint64_t hist_clock = 0;
  bool use_predictor = true;
  bool use_hermite = true;
  double* H = nullptr;          // p x ldh column-major (full symmetric after reduce)
  double* L = nullptr;          // Cholesky factor (lower, column-major p x ldh)
  double* Ldinv = nullptr;      // 1 / L_jj (ldh)
  double* Linv = nullptr;       // L^-1, row-major p x ldl (lower; allocated on first gradient call)
  int ldl = 0;
  double* zobs = nullptr;       // c3 * leverage per observation
  void* grad_plan = nullptr;    // opaque (grad.cu)
  int ldh = 0;
  double* theta_dev = nullptr;  // S (+ exp(theta))
  double* part_g = nullptr;     // [lik_blocks][lda]
  double* part_s = nullptr;     // [lik_blocks][4]  (ll, sumsq, nonfinite, -)
  int lik_blocks = 0;
  double* part_H = nullptr;     // split-K partial tiles
  size_t part_H_bytes = 0;
  bgp::EvalScalars* sc_dev = nullptr;    // [2]: [1] = snapshot of the scalars of the starting point (speculative Newton)
  bgp::EvalScalars* sc_host = nullptr;   // pinned, [2]
  // The first Newton iteration is enqueued without waiting for the scalars of the starting point (a predicted start
  // is practically never converged or non-finite): one host round trip per evaluation instead of two.
  bool speculate = true;
  // batch evaluations: modes / Hessians leave through two pinned slots; the copy into the caller's arrays is
  // deferred to the moment the next evaluation's kernels are in flight (host_hook runs there, once)
  double* pin_out[2] = {nullptr, nullptr};
  size_t pin_out_elems = 0;
  cudaEvent_t pin_ev[2] = {nullptr, nullptr};
  std::function<void()> host_hook;
  // page-locked destinations: the rotated mode / Hessian of a node leave through two device staging buffers and a
  // second stream, so the copy over PCIe never holds up the next evaluation's kernels
  cudaStream_t out_stream = nullptr;
  double* out_stage[2] = {nullptr, nullptr};     // [H (p x p, external order) | mode (p)]
  cudaEvent_t out_ready[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
  bool out_used[2] = {false, false};
  unsigned out_count = 0;
  double ll_const = 0.0;        // theta- and W-independent part of the log-likelihood
  // O-spline moment path (ospline.cu): a model whose only smoothing term is an IWP evaluates eta, g_lik and H_lik
  // from per-knot-interval moments instead of the dense design
  struct IwpTerm {
    double* x_dev = nullptr;     // covariate in the caller's row order (kept until finalize)
    double x0 = 0.0;
    int order = 0;
    std::vector<double> kneg, kpos;
  };
  std::vector<IwpTerm> iwp_terms;
  void* osp_plan = nullptr;     // opaque (ospline.cu)
  bool osp_on = false;          // the likelihood pass / Hessian go through the moment path
  bool osp_dense_grad = false;  // ... but the Laplace gradient takes its leverages from the dense design (A/B switch)
  void* syrk_plan = nullptr;    // opaque (syrk.cu)
  void* lik_plan = nullptr;     // opaque (lik.cu)
  // {64-observation chunk} x {16-column box} occupancy (rowsort.cu): bit b of occ[c] set iff chunk c has a
  // non-zero in columns 16b .. 16b+15
  uint64_t* occ_dev = nullptr;
  std::vector<uint64_t> occ_host;
  int64_t nchunks = 0;
  double hess_useful_flops = 0.0;   // structurally non-zero flops of one Hessian launch (syrk.cu)
  // ---- solver controls ---------------------------------------------------------------------
  double grad_tol = 1e-8, step_tol = 1e-8;
  int maxit = 100;
  // Certified reuse of the last Newton factorisation for the log-determinant (newton.cu): allowed when the
  // accepted full step moved the linear predictor by delta = max |d eta| with delta <= reuse_eta_tol and
  // p * delta / 2 <= reuse_rel_tol * |value|  (|logdet H(w1) - logdet H(w0)| <= p * delta).
  bool allow_reuse = true;
  double reuse_eta_tol = 1e-7, reuse_rel_tol = 1e-10;
  bool factor_is_exact = true;   // H in memory was formed at the mode itself (and L is its factor unless L_is_reversed)
  bool obs_at_mode = false;      // eta / wobs / c3 / sc_dev on the device belong to the last mode (Wmode)
  bool L_holds_H = false;        // m->L already holds a copy of m->H (written by the moment path's Hessian kernel)
  bool L_is_reversed = false;    // the gradient left the factor of H in reversed order in L (grad.cu)
  int64_t n_evals = 0, n_newton = 0, n_reuse = 0;
  int64_t n_refactor = 0;        // second attempts of the factorisation at the mode (newton.cu)
  // ---- sharding ------------------------------------------------------------------------------
  int rank = 0, world = 1;
  int64_t n_total = 0;
  bgp::Comm* comm = nullptr;
  double* red_buf = nullptr;    // [g_lik | ll | sumsq | flag | max d eta] for the all-reduce after a likelihood pass
  double* hpack = nullptr;      // packed lower triangle of the likelihood Hessian for its all-reduce
  // node group (bgp_model_set_node_group): ranks that hold the same rows and split quadrature nodes, sample blocks
  // and prediction rows between them; independent of the observation shards (a 2-D layout is allowed)
  int node_rank = 0, node_world = 1;
  bgp::Comm* node_comm = nullptr;
  bool hessian_retry = false;   // Richardson retry with larger steps when the theta Hessian is not PD (fit.cu)
  // blocks lent to fits (per-node Hessians on the device, their pinned host mirror) come back here when the fit is
  // destroyed: a fit per quadrature grid must not pay a cudaMalloc / cudaHostAlloc / cudaFree each time
  struct PoolBlock { size_t bytes; void* ptr; };
  std::vector<PoolBlock> dev_pool, pin_pool;
  std::shared_ptr<int> alive = std::make_shared<int>(1);   // fits outliving the model see 0 here
  // ---- lanes -------------------------------------------------------------------------------------
  // A batch of evaluations on the moment path leaves most of the device idle (its dominant kernel, the Cholesky, runs
  // on 8 SMs): the nodes of a batch are dealt to n_lanes evaluation contexts that run concurrently, each on its own
  // stream and host thread with its own iterate / Hessian / factor / history / moment buffers; the observations and
  // every other read-only array are shared with the parent.  Lane 0 is the model itself.
  int n_lanes = 1;
  std::vector<bgp_model*> lanes;   // lanes 1 .. n_lanes - 1 (owned)
  bool is_lane = false;
  bgp::LaneWorker* worker = nullptr;   // a lane's host thread (parked between batches)
  // ---- timing ----------------------------------------------------------------------------------
  cudaEvent_t ev[8] = {nullptr};
  double t_total = 0, t_lik = 0, t_hess = 0, t_chol = 0, t_lev = 0;
  int64_t n_lik = 0, n_hess = 0, n_chol = 0, n_lev = 0;
  double lev_flops = 0.0;          // executed flops of one leverage launch (structurally non-zero slices, grad.cu)
  // per-phase device timing: marks are (event, phase starting here); harvested after each sync
  std::vector<cudaEvent_t> ev_pool, sealed_pool;
  std::vector<int> marks, sealed_marks;
};

struct bgp_fit {
  bgp_model* model = nullptr;
  int S = 0, K = 0, p = 0, k = 0;
  std::vector<double> mode, hessian;   // S, S*S (column-major)
  int convergence = 0, fn_count = 0, gr_count = 0;
  int hessian_fallback = 0;            // number of Richardson retries with a larger step (0 = numDeriv default)
  std::vector<double> nodes;           // K x S column-major
  std::vector<double> weights, logpost, logpost_norm;
  double lognormconst = 0.0;
  // Per-node modes and Hessians (modesandhessians) stay on the device, internal column order, on the node-group
  // rank that evaluated the node: sampling reads them in place; bgp_fit_get_modes rotates and gathers on request.
  std::vector<int> owner;              // K: node-group rank holding node j
  std::vector<int> slot;               // K: slot of node j in modes_dev / Hs_dev on its owner, -1 elsewhere
  int n_local = 0;
  double* modes_dev = nullptr;         // n_local x lda
  double* Hs_dev = nullptr;            // n_local x (p x ldh)
  size_t modes_dev_bytes = 0, Hs_dev_bytes = 0, mirror_bytes = 0;
  // pinned host mirror in the caller's layout (external order): modes p x K, then Hs p x p x K.  Filled while the
  // grid is being evaluated (asynchronous copies behind each node); nodes of other node-group ranks on request.
  double* mirror = nullptr;
  std::vector<unsigned char> mirrored;  // K
  std::shared_ptr<int> model_alive;
  int64_t grid_newton_iters = 0;       // inner Newton iterations spent on the quadrature grids (this rank)
  double grid_ms = 0.0, opt_ms = 0.0;  // wall clock of the grid phase / the BFGS + Richardson phase
  std::vector<std::vector<double>> marg_theta, marg_lmp, marg_w;
  // device residents for sampling / prediction (sample.cu)
  double* samps_dev = nullptr;         // p x M column-major
  int64_t samps_M = 0;
  double* Linv_dev = nullptr;          // p x ldl scratch (L^-1 and its transpose)
  double* LinvT_dev = nullptr;
  double* mode_dev = nullptr;
};

namespace bgp {
enum { PH_OTHER = 0, PH_LIK = 1, PH_HESS = 2, PH_CHOL = 3, PH_LEV = 4 };
void phase_mark(bgp_model* m, int phase);
void phase_harvest(bgp_model* m);   // call only right after a stream synchronize: seals the marks of the finished segment
void phase_collect(bgp_model* m);   // elapsed times of the sealed segment into the timers (any time later)
}  // namespace bgp

namespace bgp {

// ---- kernels (each in its own .cu) ------------------------------------------------------------
// lik.cu: eta = A W ; per-observation likelihood ; partial g = A^T r ; block partials
// rvec != NULL: only part_g = A^T rvec is produced (W_dev ignored)
int launch_lik(bgp_model* m, const double* W_dev, bool want_c3, double tau, const double* rvec = nullptr);
int lik_max_lda();
int lik_plan_create(bgp_model* m);
void lik_plan_destroy(bgp_model* m);
// ospline.cu: the same two steps from knot-interval moments (models with one IWP term); the pass leaves
// [g_lik | ll | sumsq | flag | max d eta] in red_buf, the Hessian step H_lik in m->H
int osp_plan_create(bgp_model* m);
void osp_plan_destroy(bgp_model* m);
int osp_launch_lik(bgp_model* m, const double* W_dev, double tau, const double* theta = nullptr);
int osp_launch_hessian(bgp_model* m, const double* theta);
// A^T (c3 q) for the Laplace gradient from the moments: V is the upper factor of H^-1 (grad.cu); result in red_buf[0 .. lda)
int osp_launch_leverage(bgp_model* m, const double* V, int ldl);
// finish.cu: reduce partials (+ allreduce when sharded), add prior terms -> f / g / gmax in sc_dev
int launch_finish(bgp_model* m, const double* W_dev, const double* theta, double tau);
double theta_constant(const bgp_model* m, const double* theta);
// rowsort.cu: sort observations by zero pattern, build the occupancy map
int build_row_order(bgp_model* m);
// syrk.cu: H = A^T diag(w) A (+ allreduce when sharded) + Q(theta)
int syrk_plan_create(bgp_model* m);
void syrk_plan_destroy(bgp_model* m);
int launch_hessian(bgp_model* m, const double* theta);
// chol.cu: L = chol(H), logdet, optionally step = -H^-1 g and max|step|
// theta_tan / W_tan != NULL: also form the tangents d w_hat / d theta (at W_tan) into m->Tan on the cluster's idle ranks
// write_trial (with solve): Wtrial = W + step as well (the full Newton step, no separate axpy launch)
int launch_chol_solve(bgp_model* m, bool solve, const double* theta_tan = nullptr, const double* W_tan = nullptr,
                      bool write_trial = false);
constexpr int CHOL_TANGENT_MAX_S = 7;
int launch_tangent(bgp_model* m, const double* theta);
// basis.cu
int launch_iwp_block(bgp_model* m, const double* x_dev, int64_t n, double x0, const double* kneg, int nneg,
                     const double* kpos, int npos, int order, double* dstB, int ldB, double* dstX, int ldX,
                     bool col_major, cudaStream_t st);
// io.cu: external (ABI) <-> internal column order at the API boundary
inline int ext2int(const bgp_model* m, int e) { return e < m->p - m->nD ? e + m->nD : e - (m->p - m->nD); }
int copy_vec_in(bgp_model* m, const double* host_ext, double* dev_int);     // pads dev_int to lda with zeros
int copy_vec_out(bgp_model* m, const double* dev_int, double* host_ext);
int copy_H_out(bgp_model* m, double* host_ext);                             // from m->H (p x ldh) to p x p
int rot_vec_dev(bgp_model* m, const double* dev_int, double* dev_ext);      // p doubles, external order, on the device
int rot_H_dev(bgp_model* m, const double* H_int, double* dev_ext, int lde); // p x ldh internal -> p x lde external
int launch_sgp_block(const double* x_dev, int64_t n, double x0, double a, int k, int m, double lo, double hi, double* dstB,
                     double* dstX, cudaStream_t st);
// model.cu: evaluation lanes (see bgp_model::n_lanes)
int lanes_ensure(bgp_model* m, int count);
void lanes_destroy(bgp_model* m);
// ospline.cu: a lane's private copy of the per-evaluation buffers of the moment path
int osp_plan_clone_for_lane(const bgp_model* parent, bgp_model* lane);
void osp_plan_destroy_lane(bgp_model* lane);
// model.cu: size-keyed free lists of device / pinned blocks (see bgp_model::dev_pool)
void* pool_take(bgp_model* m, bool pinned, size_t bytes, size_t* got_bytes);
void pool_give(bgp_model* m, bool pinned, void* ptr, size_t bytes);
// basis.cu: Compute_Q_sB per harmonic into the block-diagonal d x d precision (column-major, device)
int launch_sgp_precision(double a, int k, int m, double lo, double hi, double accuracy, double* P_dev, cudaStream_t st);
// newton.cu
int eval_fg_async(bgp_model* m, const double* W_dev, const double* theta, bool want_c3);
int laplace_inner(bgp_model* m, const double* theta, double* value, int* iters);
// where a batch of evaluations leaves its modes / Hessians: caller arrays in the external order (modes p x K,
// Hs p x p x K) and / or device arrays in the internal order (modes K x lda, Hs K x (p x ldh); slot dev_slot[j])
struct BatchSink {
  double* modes_host = nullptr;
  double* Hs_host = nullptr;
  bool host_pinned = false;       // the host arrays are page-locked: copy straight into them, no staging slots
  double* modes_dev = nullptr;
  double* Hs_dev = nullptr;
  const int* dev_slot = nullptr;
};
int laplace_batch(bgp_model* m, int K, const double* theta, const unsigned char* mine, double* values,
                  const BatchSink& sink, int* iters_total, int* first_failed);
// grad.cu: d/dtheta of the Laplace objective at the mode left on the device by laplace_inner
int laplace_gradient(bgp_model* m, const double* theta, double* grad_host);
void grad_plan_destroy(bgp_model* m);

// kgemm.cu: C[m][n] = bias[m] + sum_k A[m][k] B[n][k]  (both operands K-major, row pitch multiple of 2 doubles)
//   out_col_major: out[n' * ldc + m] with n' = col_perm ? col_perm[n] : n ; else out[m * ldc + n]
int launch_kgemm(const double* A, int64_t M, int64_t lda, const double* B, int64_t N, int64_t ldb, int K,
                 const double* bias, double* out, int64_t ldc, bool out_col_major, const int32_t* col_perm,
                 cudaStream_t st, const unsigned long long* a_occ = nullptr);
//   a_occ: optional map from launch_kgemm_occ (one word per 128-row tile of A, bit s = k-slice s not all zero, K <= 1024):
//   empty slices are neither copied nor multiplied
int launch_kgemm_occ(const double* A, int64_t M, int64_t lda, int K, unsigned long long* occ, cudaStream_t st);
// grad.cu: L^-1 (row-major, lower) and optionally its transpose (row-major, upper) from m->L
int launch_trtri(bgp_model* m, double* Linv, int ldl, double* LinvT);
// sample.cu
void fit_release_device(bgp_fit* f);

// comm.cpp
int comm_unique_id(void* id128);
int comm_create(bgp_model* m, const void* id128);
void comm_destroy(bgp_model* m);
int comm_allreduce_sum(bgp_model* m, double* buf, size_t count);
int comm_open(Comm** out, int rank, int world, const void* id128);
void comm_close(Comm** c);
int node_allreduce_sum(bgp_model* m, double* buf, size_t count);
// owner of piece j of K when K pieces are dealt to `world` ranks in contiguous, balanced runs
inline int piece_owner(int64_t j, int64_t K, int world) {
  const int64_t base = K / world, rem = K % world, cut = rem * (base + 1);
  return (int)(j < cut ? j / (base + 1) : rem + (j - cut) / (base > 0 ? base : 1));
}
inline void piece_bounds(int64_t K, int rank, int world, int64_t* lo, int64_t* hi) {
  const int64_t base = K / world, rem = K % world;
  *lo = rank * base + (rank < rem ? rank : rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}

}  // namespace bgp
