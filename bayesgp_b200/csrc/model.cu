// model.cu — device-resident model: replaces the `tmbdat` list handed to TMB::MakeADFun
// (/root/reference/R/02_model_fit.R:152-183) and the DATA_* unpacking of
// /root/reference/src/BayesGP.cpp:34-73.
//
// HBM layout: the design matrix A = [B_1..B_J | X_1..X_J | Xf_0..Xf_F] is
// stored ONCE, observation-major (n x lda doubles, lda = round_up(p, 16), zero padded), so that
//   * the likelihood pass streams whole rows with 16-byte coalesced loads, and
//   * the Hessian kernel TMA-loads {16 column x 16 row} boxes whose 128-byte lines take the
//     hardware swizzle.
// R hands the blocks over column-major; they are transposed into place on the device.
#include <algorithm>

#include <cstdlib>

#include "bgp_internal.h"

namespace bgp {

// src: n x d column-major  ->  dst[i * lda + off + c]
__global__ void transpose_in_kernel(const double* __restrict__ src, int64_t n, int d, double* __restrict__ dst, int lda,
                                    int off) {
  __shared__ double tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
    const int64_t i = i0 + threadIdx.x;
    const int c = c0 + cc;
    tile[cc][threadIdx.x] = (i < n && c < d) ? src[(size_t)c * n + i] : 0.0;
  }
  __syncthreads();
  for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
    const int64_t i = i0 + ii;
    const int c = c0 + threadIdx.x;
    if (i < n && c < d) dst[(size_t)i * lda + off + c] = tile[threadIdx.x][ii];
  }
}

void* pool_take(bgp_model* m, bool pinned, size_t bytes, size_t* got_bytes) {
  auto& pool = pinned ? m->pin_pool : m->dev_pool;
  int best = -1;
  for (int i = 0; i < (int)pool.size(); ++i)
    if (pool[i].bytes >= bytes && (best < 0 || pool[i].bytes < pool[best].bytes)) best = i;
  if (best >= 0 && pool[best].bytes <= 2 * bytes + (1u << 20)) {
    void* p = pool[best].ptr;
    *got_bytes = pool[best].bytes;
    pool.erase(pool.begin() + best);
    return p;
  }
  void* p = nullptr;
  const cudaError_t e = pinned ? cudaHostAlloc(&p, bytes, cudaHostAllocDefault) : cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    set_error("%s of %zu bytes failed: %s", pinned ? "cudaHostAlloc" : "cudaMalloc", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  *got_bytes = bytes;
  return p;
}

void pool_give(bgp_model* m, bool pinned, void* ptr, size_t bytes) {
  if (!ptr) return;
  auto& pool = pinned ? m->pin_pool : m->dev_pool;
  if (pool.size() >= 8) {                     // bounded: drop the smallest block
    int small = 0;
    for (int i = 1; i < (int)pool.size(); ++i)
      if (pool[i].bytes < pool[small].bytes) small = i;
    if (pool[small].bytes < bytes) {
      if (pinned) cudaFreeHost(pool[small].ptr);
      else cudaFree(pool[small].ptr);
      pool[small] = {bytes, ptr};
    } else {
      if (pinned) cudaFreeHost(ptr);
      else cudaFree(ptr);
    }
    return;
  }
  pool.push_back({bytes, ptr});
}

static int stage_block(bgp_model* m, std::vector<bgp_model::Staged>& dst, int ncol, const double* host) {
  bgp_model::Staged s;
  s.ncol = ncol;
  s.dev = nullptr;
  if (ncol > 0) {
    if (!host) {
      set_error("NULL matrix for a block with %d columns", ncol);
      return BGP_ERR_ARG;
    }
    const size_t bytes = (size_t)m->n * ncol * sizeof(double);
    BGP_CUDA(cudaMalloc(&s.dev, bytes));
    BGP_CUDA(cudaMemcpyAsync(s.dev, host, bytes, cudaMemcpyHostToDevice, m->stream));
    BGP_CUDA(cudaStreamSynchronize(m->stream));
  }
  dst.push_back(s);
  return BGP_OK;
}

static double lgamma_sum_poisson(const double* y, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += std::lgamma(y[i] + 1.0);
  return -s;
}

static double lchoose_sum(const double* y, const double* size, int64_t n) {
  // TMB dbinom_robust adds lgamma(size+1) - lgamma(k+1) - lgamma(size-k+1) only when size > 1
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double sz = size ? size[i] : 1.0;
    if (sz > 1.0) s += std::lgamma(sz + 1.0) - std::lgamma(y[i] + 1.0) - std::lgamma(sz - y[i] + 1.0);
  }
  return s;
}

}  // namespace bgp

namespace bgp {

static void lane_free(bgp_model* l) {
  delete l->worker;             // joins the lane's host thread
  l->worker = nullptr;
  if (l->stream) cudaStreamSynchronize(l->stream);
  if (l->out_stream) cudaStreamSynchronize(l->out_stream);
  osp_plan_destroy_lane(l);
  for (double* ptr : {l->W, l->Wtrial, l->Wmode, l->g, l->step, l->Tan, l->xbuf, l->H, l->L, l->Ldinv, l->red_buf})
    if (ptr) cudaFree(ptr);
  for (auto& h : l->hist) {
    if (h.W) cudaFree(h.W);
    if (h.T) cudaFree(h.T);
  }
  if (l->sc_dev) cudaFree(l->sc_dev);
  if (l->sc_host) cudaFreeHost(l->sc_host);
  for (int i = 0; i < 2; ++i) {
    if (l->out_stage[i]) cudaFree(l->out_stage[i]);
    if (l->out_ready[i]) cudaEventDestroy(l->out_ready[i]);
    if (l->out_done[i]) cudaEventDestroy(l->out_done[i]);
    if (l->pin_out[i]) cudaFreeHost(l->pin_out[i]);
    if (l->pin_ev[i]) cudaEventDestroy(l->pin_ev[i]);
  }
  if (l->out_stream) cudaStreamDestroy(l->out_stream);
  for (int i = 0; i < 8; ++i)
    if (l->ev[i]) cudaEventDestroy(l->ev[i]);
  for (cudaEvent_t e : l->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : l->sealed_pool) cudaEventDestroy(e);
  if (l->stream) cudaStreamDestroy(l->stream);
  delete l;
}

// Lanes 1 .. count - 1 of a finalized single-device model on the moment path: a copy of the model's description that
// shares every read-only device array with it and owns the state one evaluation writes.
int lanes_ensure(bgp_model* m, int count) {
  while ((int)m->lanes.size() < count - 1) {
    bgp_model* l = new bgp_model(*m);
    l->is_lane = true;
    l->n_lanes = 1;
    l->lanes.clear();
    l->worker = nullptr;
    // nothing below may be shared with the parent: reset, then allocate
    l->stream = nullptr;
    l->W = l->Wtrial = l->Wmode = l->g = l->step = l->Tan = l->xbuf = l->H = l->L = l->Ldinv = l->red_buf = nullptr;
    for (auto& h : l->hist) {
      h.W = h.T = nullptr;
      h.stamp = 0;
    }
    l->sc_dev = nullptr;
    l->sc_host = nullptr;
    l->theta_dev = nullptr;
    for (int i = 0; i < 2; ++i) {
      l->pin_out[i] = nullptr;
      l->pin_ev[i] = nullptr;
      l->out_stage[i] = nullptr;
      l->out_ready[i] = l->out_done[i] = nullptr;
      l->out_used[i] = false;
    }
    l->pin_out_elems = 0;
    l->out_stream = nullptr;
    l->out_count = 0;
    l->host_hook = nullptr;
    for (int i = 0; i < 8; ++i) l->ev[i] = nullptr;
    l->ev_pool.clear();
    l->sealed_pool.clear();
    l->marks.clear();
    l->sealed_marks.clear();
    l->grad_plan = nullptr;       // lanes never take gradients
    l->Linv = l->zobs = nullptr;
    l->hpack = nullptr;
    l->dev_pool.clear();
    l->pin_pool.clear();
    l->alive = std::make_shared<int>(1);
    l->osp_plan = nullptr;
    l->t_total = l->t_lik = l->t_hess = l->t_chol = l->t_lev = 0.0;
    l->n_lik = l->n_hess = l->n_chol = l->n_lev = 0;
    l->n_evals = l->n_newton = l->n_reuse = 0;
    m->lanes.push_back(l);        // owned from here on: freed with the model whatever happens below
    BGP_CUDA(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
    const size_t vb = (size_t)m->lda * sizeof(double);
    auto dalloc = [&](double** ptr, size_t bytes) -> int {
      BGP_CUDA(cudaMalloc(ptr, bytes));
      BGP_CUDA(cudaMemsetAsync(*ptr, 0, bytes, l->stream));
      return BGP_OK;
    };
    for (double** ptr : {&l->W, &l->Wtrial, &l->Wmode, &l->g, &l->step}) BGP_TRY(dalloc(ptr, vb));
    BGP_TRY(dalloc(&l->Tan, (size_t)std::max(1, m->S) * vb));
    for (auto& h : l->hist) {
      BGP_TRY(dalloc(&h.T, (size_t)std::max(1, m->S) * vb));
      BGP_TRY(dalloc(&h.W, vb));
    }
    const size_t hb = (size_t)m->ldh * m->p * sizeof(double);
    BGP_TRY(dalloc(&l->H, hb));
    BGP_TRY(dalloc(&l->L, hb));
    BGP_TRY(dalloc(&l->Ldinv, (size_t)m->ldh * sizeof(double)));
    BGP_TRY(dalloc(&l->xbuf, std::max((size_t)m->lda, (size_t)m->p * m->p) * sizeof(double)));
    BGP_TRY(dalloc(&l->red_buf, ((size_t)m->lda + 8) * sizeof(double)));
    BGP_CUDA(cudaMalloc(&l->sc_dev, 2 * sizeof(EvalScalars)));
    BGP_CUDA(cudaMemsetAsync(l->sc_dev, 0, 2 * sizeof(EvalScalars), l->stream));
    BGP_CUDA(cudaMallocHost(&l->sc_host, 2 * sizeof(EvalScalars)));
    for (int i = 0; i < 8; ++i) BGP_CUDA(cudaEventCreate(&l->ev[i]));
    BGP_TRY(osp_plan_clone_for_lane(m, l));
    BGP_CUDA(cudaStreamSynchronize(l->stream));
    l->worker = new LaneWorker();
  }
  return BGP_OK;
}

void lanes_destroy(bgp_model* m) {
  for (bgp_model* l : m->lanes) lane_free(l);
  m->lanes.clear();
}

}  // namespace bgp

using namespace bgp;

extern "C" {

int bgp_model_new(int64_t n, int family, const double* y, const double* size, int device, bgp_model** out) {
  if (!out || !y || n <= 0) {
    set_error("bgp_model_new: bad arguments");
    return BGP_ERR_ARG;
  }
  if (family != BGP_FAMILY_GAUSSIAN && family != BGP_FAMILY_POISSON && family != BGP_FAMILY_BINOMIAL &&
      family != BGP_FAMILY_NONE) {
    set_error("bgp_model_new: unsupported family code %d (Coxph / case-crossover are out of scope)", family);
    return BGP_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available: libbgp has no CPU fallback");
    return BGP_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error("device %d out of range (%d devices)", device, ndev);
    return BGP_ERR_ARG;
  }
  BGP_CUDA(cudaSetDevice(device));
  bgp_model* m = new bgp_model();
  m->n = n;
  m->n_total = n;
  m->family = family;
  m->device = device;
  if (const char* e = getenv("BGP_NO_PREDICTOR")) m->use_predictor = !(e[0] == '1');   // diagnostics only
  if (const char* e = getenv("BGP_NO_HERMITE")) m->use_hermite = !(e[0] == '1');
  int st = [&]() -> int {
    BGP_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    const size_t nb = (size_t)n * sizeof(double);
    const size_t nb_pad = (size_t)(round_up64(n, 64) + 64) * sizeof(double);   // bulk copies read whole 8-row stages
    BGP_CUDA(cudaMalloc(&m->y, nb_pad));
    BGP_CUDA(cudaMemsetAsync(m->y, 0, nb_pad, m->stream));
    BGP_CUDA(cudaMemcpyAsync(m->y, y, nb, cudaMemcpyHostToDevice, m->stream));
    if (family == BGP_FAMILY_BINOMIAL) {
      BGP_CUDA(cudaMalloc(&m->size, nb_pad));
      BGP_CUDA(cudaMemsetAsync(m->size, 0, nb_pad, m->stream));
      if (size) {
        BGP_CUDA(cudaMemcpyAsync(m->size, size, nb, cudaMemcpyHostToDevice, m->stream));
      } else {
        std::vector<double> ones((size_t)n, 1.0);   // R/02_model_fit.R:176-183
        BGP_CUDA(cudaMemcpyAsync(m->size, ones.data(), nb, cudaMemcpyHostToDevice, m->stream));
        BGP_CUDA(cudaStreamSynchronize(m->stream));
      }
    }
    BGP_CUDA(cudaStreamSynchronize(m->stream));
    return BGP_OK;
  }();
  if (st != BGP_OK) {
    bgp_model_destroy(m);
    return st;
  }
  if (family == BGP_FAMILY_POISSON) m->ll_const = lgamma_sum_poisson(y, n);
  else if (family == BGP_FAMILY_BINOMIAL) m->ll_const = lchoose_sum(y, size, n);
  else if (family == BGP_FAMILY_GAUSSIAN) m->ll_const = -0.5 * (double)n * std::log(2.0 * M_PI);
  *out = m;
  return BGP_OK;
}

#define BGP_CHECK_BUILDING(m)                                       \
  do {                                                              \
    if (!(m)) {                                                     \
      set_error("NULL model handle");                               \
      return BGP_ERR_ARG;                                           \
    }                                                               \
    if ((m)->finalized) {                                           \
      set_error("model already finalized");                         \
      return BGP_ERR_STATE;                                         \
    }                                                               \
    BGP_CUDA(cudaSetDevice((m)->device));                           \
  } while (0)

int bgp_model_add_random(bgp_model* m, int d, const double* B, const double* P, int p_is_diag, double logPdet, double u,
                         double alpha) {
  BGP_CHECK_BUILDING(m);
  if (d <= 0 || !B || !P) {
    set_error("bgp_model_add_random: bad arguments");
    return BGP_ERR_ARG;
  }
  if ((int)m->rnd.size() >= 16) {
    set_error("at most 16 smoothing terms are supported");
    return BGP_ERR_ARG;
  }
  BGP_TRY(stage_block(m, m->st_rnd, d, B));
  RandomBlock rb;
  rb.d = d;
  rb.diag = p_is_diag != 0;
  rb.logPdet = logPdet;
  rb.u = u;
  rb.alpha = alpha;
  const size_t pb = (rb.diag ? (size_t)d : (size_t)d * d) * sizeof(double);
  BGP_CUDA(cudaMalloc(&rb.P_dev, pb));
  BGP_CUDA(cudaMemcpy(rb.P_dev, P, pb, cudaMemcpyHostToDevice));
  rb.P_host.assign(P, P + pb / sizeof(double));
  m->rnd.push_back(rb);
  return BGP_OK;
}

int bgp_model_add_boundary(bgp_model* m, int ncol, const double* X, double prec, double mean) {
  BGP_CHECK_BUILDING(m);
  if (ncol < 0) {
    set_error("bgp_model_add_boundary: negative column count");
    return BGP_ERR_ARG;
  }
  BGP_TRY(stage_block(m, m->st_bnd, ncol, X));
  m->bnd_dim.push_back(ncol);
  m->bnd_prec.push_back(prec);
  m->bnd_mean.push_back(mean);
  return BGP_OK;
}

int bgp_model_add_fixed(bgp_model* m, int ncol, const double* Xf, double prec, double mean) {
  BGP_CHECK_BUILDING(m);
  if (ncol <= 0) {
    set_error("bgp_model_add_fixed: column count must be positive");
    return BGP_ERR_ARG;
  }
  BGP_TRY(stage_block(m, m->st_fix, ncol, Xf));
  m->fix_dim.push_back(ncol);
  m->fix_prec.push_back(prec);
  m->fix_mean.push_back(mean);
  return BGP_OK;
}

int bgp_model_set_noise_prior(bgp_model* m, double u, double alpha) {
  BGP_CHECK_BUILDING(m);
  m->noise_u = u;
  m->noise_alpha = alpha;
  return BGP_OK;
}

int bgp_model_add_iwp(bgp_model* m, const double* x, double initial_location, const double* knots, int nknots, int order,
                      double u, double alpha, double boundary_prec, double boundary_mean) {
  BGP_CHECK_BUILDING(m);
  if (!x || !knots || nknots < 2 || order < 1 || order > 8) {
    set_error("bgp_model_add_iwp: bad arguments (order must be 1..8)");
    return BGP_ERR_ARG;
  }
  if ((int)m->rnd.size() >= 16) {
    set_error("at most 16 smoothing terms are supported");
    return BGP_ERR_ARG;
  }
  // knot split of local_poly_helper / compute_weights_precision (R/01_utility.R:325-344,378-401)
  std::vector<double> kneg, kpos;
  double kmin = knots[0], kmax = knots[0];
  for (int i = 1; i < nknots; ++i) {
    kmin = std::min(kmin, knots[i]);
    kmax = std::max(kmax, knots[i]);
  }
  auto uniq_sorted = [](std::vector<double>& v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  };
  if (kmin >= 0) {
    kpos.assign(knots, knots + nknots);
  } else {
    for (int i = 0; i < nknots; ++i) kneg.push_back(knots[i] < 0 ? -knots[i] : 0.0);
    uniq_sorted(kneg);
    if (kmax > 0) {
      for (int i = 0; i < nknots; ++i) kpos.push_back(knots[i] > 0 ? knots[i] : 0.0);
      uniq_sorted(kpos);
    }
  }
  const int nneg = kneg.empty() ? 0 : (int)kneg.size() - 1;
  const int npos = kpos.empty() ? 0 : (int)kpos.size() - 1;
  const int d = nneg + npos;
  if (d <= 0) {
    set_error("bgp_model_add_iwp: knots define no basis function");
    return BGP_ERR_ARG;
  }
  std::vector<double> Pdiag;
  for (int i = 0; i < nneg; ++i) Pdiag.push_back(kneg[i + 1] - kneg[i]);
  for (int i = 0; i < npos; ++i) Pdiag.push_back(kpos[i + 1] - kpos[i]);
  double logPdet = 0.0;
  for (double v : Pdiag) logPdet += std::log(v);

  double *x_dev = nullptr, *kn_dev = nullptr, *kp_dev = nullptr;
  const size_t nb = (size_t)m->n * sizeof(double);
  BGP_CUDA(cudaMalloc(&x_dev, nb));
  BGP_CUDA(cudaMemcpyAsync(x_dev, x, nb, cudaMemcpyHostToDevice, m->stream));
  if (!kneg.empty()) {
    BGP_CUDA(cudaMalloc(&kn_dev, kneg.size() * sizeof(double)));
    BGP_CUDA(cudaMemcpyAsync(kn_dev, kneg.data(), kneg.size() * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  }
  if (!kpos.empty()) {
    BGP_CUDA(cudaMalloc(&kp_dev, kpos.size() * sizeof(double)));
    BGP_CUDA(cudaMemcpyAsync(kp_dev, kpos.data(), kpos.size() * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  }
  bgp_model::Staged sB, sX;
  sB.ncol = d;
  sX.ncol = order - 1;
  sB.dev = sX.dev = nullptr;
  BGP_CUDA(cudaMalloc(&sB.dev, (size_t)m->n * d * sizeof(double)));
  if (sX.ncol > 0) BGP_CUDA(cudaMalloc(&sX.dev, (size_t)m->n * sX.ncol * sizeof(double)));
  BGP_TRY(launch_iwp_block(m, x_dev, m->n, initial_location, kn_dev, (int)kneg.size(), kp_dev, (int)kpos.size(), order,
                           sB.dev, (int)m->n, sX.dev, (int)m->n, true, m->stream));
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  {
    // the covariate stays until finalize: a model with this single term takes the O-spline moment path (ospline.cu)
    bgp_model::IwpTerm it;
    it.x_dev = x_dev;
    it.x0 = initial_location;
    it.order = order;
    it.kneg = kneg;
    it.kpos = kpos;
    m->iwp_terms.push_back(it);
  }
  if (kn_dev) cudaFree(kn_dev);
  if (kp_dev) cudaFree(kp_dev);
  m->st_rnd.push_back(sB);
  m->st_bnd.push_back(sX);
  RandomBlock rb;
  rb.d = d;
  rb.diag = true;
  rb.logPdet = logPdet;
  rb.u = u;
  rb.alpha = alpha;
  BGP_CUDA(cudaMalloc(&rb.P_dev, (size_t)d * sizeof(double)));
  BGP_CUDA(cudaMemcpy(rb.P_dev, Pdiag.data(), (size_t)d * sizeof(double), cudaMemcpyHostToDevice));
  rb.P_host = Pdiag;
  m->rnd.push_back(rb);
  m->bnd_dim.push_back(order - 1);
  m->bnd_prec.push_back(boundary_prec);
  m->bnd_mean.push_back(boundary_mean);
  return BGP_OK;
}

int bgp_model_add_sgp(bgp_model* m, const double* x, double initial_location, double a, int k, int nharm, const double* region,
                      const double* P, double logPdet, double u, double alpha, double boundary_prec, double boundary_mean) {
  BGP_CHECK_BUILDING(m);
  if (k < 3) {      // R/02_model_fit.R:511-514
    set_error("Error: parameter <k> in the random effect part should be >= 3.");
    return BGP_ERR_ARG;
  }
  if (k < 4) {      // R accepts k = 3, then fda::create.bspline.basis(nbasis = 3, norder = 4) stops (R/01_utility.R:179-183)
    set_error("sGP with k = 3: fda::create.bspline.basis needs nbasis >= norder = 4");
    return BGP_ERR_ARG;
  }
  if (!x || !region || !P || nharm < 1 || !(region[1] > region[0])) {
    set_error("bgp_model_add_sgp: bad arguments (m >= 1, region increasing)");
    return BGP_ERR_ARG;
  }
  if ((int)m->rnd.size() >= 16) {
    set_error("at most 16 smoothing terms are supported");
    return BGP_ERR_ARG;
  }
  const int d = 3 * (k - 2) * nharm;
  double* x_dev = nullptr;
  const size_t nb = (size_t)m->n * sizeof(double);
  BGP_CUDA(cudaMalloc(&x_dev, nb));
  BGP_CUDA(cudaMemcpyAsync(x_dev, x, nb, cudaMemcpyHostToDevice, m->stream));
  bgp_model::Staged sB, sX;
  sB.ncol = d;
  sX.ncol = 2 * nharm;
  sB.dev = sX.dev = nullptr;
  BGP_CUDA(cudaMalloc(&sB.dev, (size_t)m->n * d * sizeof(double)));
  BGP_CUDA(cudaMalloc(&sX.dev, (size_t)m->n * sX.ncol * sizeof(double)));
  BGP_TRY(launch_sgp_block(x_dev, m->n, initial_location, a, k, nharm, region[0], region[1], sB.dev, sX.dev, m->stream));
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  cudaFree(x_dev);
  m->st_rnd.push_back(sB);
  m->st_bnd.push_back(sX);
  RandomBlock rb;
  rb.d = d;
  rb.diag = false;
  rb.logPdet = logPdet;
  rb.u = u;
  rb.alpha = alpha;
  const size_t pb = (size_t)d * d * sizeof(double);
  BGP_CUDA(cudaMalloc(&rb.P_dev, pb));
  BGP_CUDA(cudaMemcpy(rb.P_dev, P, pb, cudaMemcpyHostToDevice));
  rb.P_host.assign(P, P + (size_t)d * d);
  m->rnd.push_back(rb);
  m->bnd_dim.push_back(2 * nharm);
  m->bnd_prec.push_back(boundary_prec);
  m->bnd_mean.push_back(boundary_mean);
  return BGP_OK;
}

// determinant(P)$modulus (R/02_model_fit.R:66): log |det P| by LU with partial pivoting, P d x d column-major
static double log_abs_det(std::vector<double> A, int d) {
  double s = 0.0;
  for (int c = 0; c < d; ++c) {
    int piv = c;
    for (int r = c + 1; r < d; ++r)
      if (std::fabs(A[(size_t)c * d + r]) > std::fabs(A[(size_t)c * d + piv])) piv = r;
    const double pv = A[(size_t)c * d + piv];
    if (pv == 0.0) return -INFINITY;
    if (piv != c)
      for (int j = c; j < d; ++j) std::swap(A[(size_t)j * d + c], A[(size_t)j * d + piv]);
    s += std::log(std::fabs(pv));
    for (int r = c + 1; r < d; ++r) A[(size_t)c * d + r] /= pv;
    for (int j = c + 1; j < d; ++j) {
      const double f = A[(size_t)j * d + c];
      if (f == 0.0) continue;
      double* col = &A[(size_t)j * d];
      const double* l = &A[(size_t)c * d];
      for (int r = c + 1; r < d; ++r) col[r] -= l[r] * f;
    }
  }
  return s;
}

int bgp_sgp_precision(double a, int k, int nharm, const double* region, double accuracy, int device, double* P,
                      double* logPdet) {
  if (k < 4 || nharm < 1 || !region || !(region[1] > region[0]) || !(accuracy > 0.0) || !P) {
    set_error("bgp_sgp_precision: bad arguments (k >= 4, m >= 1, region increasing, accuracy > 0)");
    return BGP_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("no CUDA device %d available: libbgp has no CPU fallback", device);
    return BGP_ERR_CUDA;
  }
  BGP_CUDA(cudaSetDevice(device));
  const int d = 3 * (k - 2) * nharm;
  double* P_dev = nullptr;
  BGP_CUDA(cudaMalloc(&P_dev, (size_t)d * d * sizeof(double)));
  int rc = launch_sgp_precision(a, k, nharm, region[0], region[1], accuracy, P_dev, 0);
  if (rc == BGP_OK && cudaMemcpy(P, P_dev, (size_t)d * d * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("cudaMemcpy of the sGP precision failed");
    rc = BGP_ERR_CUDA;
  }
  cudaFree(P_dev);
  if (rc == BGP_OK && logPdet) *logPdet = log_abs_det(std::vector<double>(P, P + (size_t)d * d), d);
  return rc;
}

int bgp_model_add_sgp_auto(bgp_model* m, const double* x, double initial_location, double a, int k, int nharm,
                           const double* region, double accuracy, double u, double alpha, double boundary_prec,
                           double boundary_mean) {
  BGP_CHECK_BUILDING(m);
  if (k < 4 || nharm < 1 || !region) {
    set_error("bgp_model_add_sgp_auto: bad arguments");
    return BGP_ERR_ARG;
  }
  const int d = 3 * (k - 2) * nharm;
  std::vector<double> P((size_t)d * d);
  double logPdet = 0.0;
  BGP_TRY(bgp_sgp_precision(a, k, nharm, region, accuracy, m->device, P.data(), &logPdet));
  return bgp_model_add_sgp(m, x, initial_location, a, k, nharm, region, P.data(), logPdet, u, alpha, boundary_prec,
                           boundary_mean);
}

int bgp_nccl_unique_id(void* id128) { return comm_unique_id(id128); }

int bgp_model_set_shard(bgp_model* m, int rank, int world, const void* nccl_unique_id) {
  BGP_CHECK_BUILDING(m);
  if (world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_unique_id)) {
    set_error("bgp_model_set_shard: bad arguments");
    return BGP_ERR_ARG;
  }
  m->rank = rank;
  m->world = world;
  if (world > 1) BGP_TRY(comm_create(m, nccl_unique_id));
  return BGP_OK;
}

int bgp_model_set_node_group(bgp_model* m, int rank, int world, const void* nccl_unique_id) {
  if (!m || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_unique_id)) {
    set_error("bgp_model_set_node_group: bad arguments");
    return BGP_ERR_ARG;
  }
  BGP_CUDA(cudaSetDevice(m->device));
  comm_close(&m->node_comm);
  m->node_rank = rank;
  m->node_world = world;
  if (world > 1) {
    BGP_TRY(comm_open(&m->node_comm, rank, world, nccl_unique_id));
    // NCCL sets its channels up on the first collective (about a second): pay for it here, not inside the first fit
    double* tmp = nullptr;
    BGP_CUDA(cudaMalloc(&tmp, 8 * sizeof(double)));
    BGP_CUDA(cudaMemsetAsync(tmp, 0, 8 * sizeof(double), m->stream));
    const int st = node_allreduce_sum(m, tmp, 8);
    cudaStreamSynchronize(m->stream);
    cudaFree(tmp);
    BGP_TRY(st);
  }
  return BGP_OK;
}

int bgp_model_set_ospline(bgp_model* m, int on) {
  if (!m || !m->finalized) {
    set_error("bgp_model_set_ospline: model is not finalized");
    return BGP_ERR_STATE;
  }
  if (on && !m->osp_plan) {
    set_error("bgp_model_set_ospline: the model is not eligible (one IWP term of order <= 4 built by bgp_model_add_iwp, at most 8 dense columns)");
    return BGP_ERR_ARG;
  }
  m->osp_on = on != 0;
  m->osp_dense_grad = on == 2;
  m->obs_at_mode = false;
  m->L_holds_H = false;
  return BGP_OK;
}

int bgp_model_get_ospline(const bgp_model* m, int* eligible, int* on) {
  if (!m) return BGP_ERR_ARG;
  if (eligible) *eligible = m->osp_plan ? 1 : 0;
  if (on) *on = m->osp_on ? 1 : 0;
  return BGP_OK;
}

int bgp_model_set_lanes(bgp_model* m, int lanes) {
  if (!m || !m->finalized) {
    set_error("bgp_model_set_lanes: model is not finalized");
    return BGP_ERR_STATE;
  }
  if (lanes < 1 || lanes > 16) {
    set_error("bgp_model_set_lanes: 1 .. 16 lanes");
    return BGP_ERR_ARG;
  }
  m->n_lanes = lanes;
  return BGP_OK;
}

int bgp_model_get_lanes(const bgp_model* m, int* lanes) {
  if (!m || !lanes) return BGP_ERR_ARG;
  *lanes = m->n_lanes;
  return BGP_OK;
}

int bgp_model_ospline_bytes(const bgp_model* m, double* bytes_per_pass) {
  if (!m || !m->osp_plan) {
    set_error("bgp_model_ospline_bytes: the model has no O-spline moment path");
    return BGP_ERR_ARG;
  }
  // per observation: u, y, previous eta in, eta out, the dense columns (+ size for the Binomial family)
  const double per_obs = 8.0 * (4 + m->nD + (m->family == BGP_FAMILY_BINOMIAL ? 1 : 0));
  if (bytes_per_pass) *bytes_per_pass = per_obs * (double)m->n;
  return BGP_OK;
}

int bgp_model_set_hessian_retry(bgp_model* m, int allow) {
  if (!m) return BGP_ERR_ARG;
  m->hessian_retry = allow != 0;
  return BGP_OK;
}

int bgp_model_finalize(bgp_model* m) {
  BGP_CHECK_BUILDING(m);
  if (m->st_fix.empty() && m->st_bnd.empty() && m->st_rnd.empty()) {
    set_error("model has no design columns");
    return BGP_ERR_ARG;
  }
  m->J = (int)m->rnd.size();
  m->S = m->J + (m->family == BGP_FAMILY_GAUSSIAN ? 1 : 0);
  int p = 0;
  for (int d : m->bnd_dim) p += d;
  for (int d : m->fix_dim) p += d;
  m->nD = p;                       // internal order: dense blocks first (bgp_internal.h)
  for (auto& rb : m->rnd) {
    rb.off = p;
    p += rb.d;
  }
  m->p = p;
  m->lda = round_up(p, 16);
  m->ldh = round_up(p, 8);
  if (m->lda > lik_max_lda()) {
    set_error("latent dimension p = %d exceeds the supported maximum %d", p, lik_max_lda());
    return BGP_ERR_ARG;
  }
  for (int j = 0; j < m->J; ++j) {
    m->theta_u.push_back(m->rnd[j].u);
    m->theta_alpha.push_back(m->rnd[j].alpha);
  }
  if (m->family == BGP_FAMILY_GAUSSIAN) {
    m->theta_u.push_back(m->noise_u);
    m->theta_alpha.push_back(m->noise_alpha);
  }
  const int64_t n = m->n;
  const size_t abytes = (size_t)n * m->lda * sizeof(double);
  BGP_CUDA(cudaMalloc(&m->A, abytes));
  BGP_CUDA(cudaMemsetAsync(m->A, 0, abytes, m->stream));
  int off = 0;
  auto place = [&](std::vector<bgp_model::Staged>& v) -> int {
    for (auto& s : v) {
      if (s.ncol > 0) {
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((s.ncol + 31) / 32)), block(32, 8);
        transpose_in_kernel<<<grid, block, 0, m->stream>>>(s.dev, n, s.ncol, m->A, m->lda, off);
        count_launch();
        BGP_CUDA(cudaGetLastError());
      }
      off += s.ncol;
    }
    return BGP_OK;
  };
  BGP_TRY(place(m->st_bnd));
  BGP_TRY(place(m->st_fix));
  BGP_TRY(place(m->st_rnd));
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  BGP_TRY(osp_plan_create(m));     // reads the staged dense columns, y and size in the caller's row order
  for (auto& it : m->iwp_terms)
    if (it.x_dev) {
      cudaFree(it.x_dev);
      it.x_dev = nullptr;
    }
  for (auto* v : {&m->st_rnd, &m->st_bnd, &m->st_fix}) {
    for (auto& s : *v)
      if (s.dev) cudaFree(s.dev);
    v->clear();
  }
  // prior mean and the theta-independent diagonal of Q (src/BayesGP.cpp:222-238)
  std::vector<double> mu0((size_t)m->lda, 0.0), qfix((size_t)m->lda, 0.0);
  int o = 0;
  for (size_t b = 0; b < m->bnd_dim.size(); ++b)
    for (int c = 0; c < m->bnd_dim[b]; ++c, ++o) {
      mu0[o] = m->bnd_mean[b];
      qfix[o] = m->bnd_prec[b];
    }
  for (size_t b = 0; b < m->fix_dim.size(); ++b)
    for (int c = 0; c < m->fix_dim[b]; ++c, ++o) {
      mu0[o] = m->fix_mean[b];
      qfix[o] = m->fix_prec[b];
    }
  const size_t vb = (size_t)m->lda * sizeof(double);
  auto dalloc = [&](double** ptr, size_t bytes) -> int {
    BGP_CUDA(cudaMalloc(ptr, bytes));
    BGP_CUDA(cudaMemsetAsync(*ptr, 0, bytes, m->stream));     // in stream order (the stream is non-blocking)
    return BGP_OK;
  };
  m->qfix_host.assign(qfix.begin(), qfix.begin() + m->p);
  BGP_TRY(dalloc(&m->xbuf, std::max((size_t)m->lda, (size_t)m->p * m->p) * sizeof(double)));
  BGP_TRY(dalloc(&m->mu0, vb));
  BGP_TRY(dalloc(&m->qfix, vb));
  BGP_CUDA(cudaMemcpyAsync(m->mu0, mu0.data(), vb, cudaMemcpyHostToDevice, m->stream));
  BGP_CUDA(cudaMemcpyAsync(m->qfix, qfix.data(), vb, cudaMemcpyHostToDevice, m->stream));
  for (double** ptr : {&m->W, &m->Wtrial, &m->Wmode, &m->g, &m->step}) BGP_TRY(dalloc(ptr, vb));
  const size_t nob = (size_t)(round_up64(n, 64) + 64) * sizeof(double);
  BGP_TRY(dalloc(&m->eta, nob));
  BGP_TRY(dalloc(&m->wobs, nob));
  BGP_TRY(dalloc(&m->c3, nob));
  const size_t hb = (size_t)m->ldh * m->p * sizeof(double);
  BGP_TRY(dalloc(&m->H, hb));
  BGP_TRY(dalloc(&m->L, hb));
  BGP_TRY(dalloc(&m->Ldinv, (size_t)m->ldh * sizeof(double)));
  BGP_TRY(dalloc(&m->theta_dev, 64 * sizeof(double)));
  BGP_TRY(dalloc(&m->Tan, (size_t)std::max(1, m->S) * m->lda * sizeof(double)));
  for (auto& h : m->hist) {
    BGP_TRY(dalloc(&h.T, (size_t)std::max(1, m->S) * m->lda * sizeof(double)));
    BGP_TRY(dalloc(&h.W, vb));
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  m->lik_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(sms, (n + 7) / 8));   // one persistent CTA per SM
  BGP_TRY(dalloc(&m->part_g, (size_t)m->lik_blocks * m->lda * sizeof(double)));
  BGP_TRY(dalloc(&m->part_s, (size_t)m->lik_blocks * 4 * sizeof(double)));
  BGP_TRY(dalloc(&m->red_buf, ((size_t)m->lda + 8) * sizeof(double)));
  BGP_CUDA(cudaMalloc(&m->sc_dev, 2 * sizeof(EvalScalars)));
  BGP_CUDA(cudaMemsetAsync(m->sc_dev, 0, 2 * sizeof(EvalScalars), m->stream));
  BGP_CUDA(cudaMallocHost(&m->sc_host, 2 * sizeof(EvalScalars)));
  BGP_CUDA(cudaStreamSynchronize(m->stream));     // the zero fills and the two copies above
  if (const char* e = getenv("BGP_NO_SPECULATION")) m->speculate = !(e[0] == '1');   // diagnostics only
  if (const char* e = getenv("BGP_LANES")) m->n_lanes = std::max(1, std::min(16, atoi(e)));
  for (int i = 0; i < 8; ++i) BGP_CUDA(cudaEventCreate(&m->ev[i]));
  BGP_TRY(build_row_order(m));
  BGP_TRY(syrk_plan_create(m));
  BGP_TRY(lik_plan_create(m));
  if (m->world > 1) {
    // the likelihood constant and n are global quantities
    double buf[2] = {m->ll_const, (double)m->n};
    BGP_CUDA(cudaMemcpy(m->red_buf, buf, sizeof(buf), cudaMemcpyHostToDevice));
    BGP_TRY(comm_allreduce_sum(m, m->red_buf, 2));
    BGP_CUDA(cudaStreamSynchronize(m->stream));
    BGP_CUDA(cudaMemcpy(buf, m->red_buf, sizeof(buf), cudaMemcpyDeviceToHost));
    m->ll_const = buf[0];
    m->n_total = (int64_t)std::llround(buf[1]);
  }
  m->finalized = true;
  return BGP_OK;
}

void bgp_model_destroy(bgp_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  lanes_destroy(m);
  syrk_plan_destroy(m);
  lik_plan_destroy(m);
  grad_plan_destroy(m);
  osp_plan_destroy(m);
  for (auto& it : m->iwp_terms)
    if (it.x_dev) cudaFree(it.x_dev);
  comm_destroy(m);
  for (auto* v : {&m->st_rnd, &m->st_bnd, &m->st_fix})
    for (auto& s : *v)
      if (s.dev) cudaFree(s.dev);
  for (auto& rb : m->rnd)
    if (rb.P_dev) cudaFree(rb.P_dev);
  for (double* ptr : {m->A, m->y, m->size, m->eta, m->wobs, m->c3, m->qfix, m->mu0, m->W, m->Wtrial, m->Wmode, m->g,
                      m->step, m->Tan, m->xbuf, m->H, m->L, m->Ldinv, m->theta_dev, m->part_g, m->part_s, m->part_H, m->red_buf,
                      m->hpack})
    if (ptr) cudaFree(ptr);
  for (auto& h : m->hist) {
    if (h.W) cudaFree(h.W);
    if (h.T) cudaFree(h.T);
  }
  *m->alive = 0;
  for (auto& b : m->dev_pool) cudaFree(b.ptr);
  for (auto& b : m->pin_pool) cudaFreeHost(b.ptr);
  if (m->occ_dev) cudaFree(m->occ_dev);
  if (m->sc_dev) cudaFree(m->sc_dev);
  if (m->sc_host) cudaFreeHost(m->sc_host);
  for (int i = 0; i < 2; ++i) {
    if (m->out_stage[i]) cudaFree(m->out_stage[i]);
    if (m->out_ready[i]) cudaEventDestroy(m->out_ready[i]);
    if (m->out_done[i]) cudaEventDestroy(m->out_done[i]);
  }
  if (m->out_stream) cudaStreamDestroy(m->out_stream);
  for (int i = 0; i < 2; ++i) {
    if (m->pin_out[i]) cudaFreeHost(m->pin_out[i]);
    if (m->pin_ev[i]) cudaEventDestroy(m->pin_ev[i]);
  }
  for (int i = 0; i < 8; ++i)
    if (m->ev[i]) cudaEventDestroy(m->ev[i]);
  for (cudaEvent_t e : m->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : m->sealed_pool) cudaEventDestroy(e);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int bgp_model_dims(const bgp_model* m, int64_t* n, int* p, int* S) {
  if (!m || !m->finalized) {
    set_error("model not finalized");
    return BGP_ERR_STATE;
  }
  if (n) *n = m->n;
  if (p) *p = m->p;
  if (S) *S = m->S;
  return BGP_OK;
}

}  // extern "C"
