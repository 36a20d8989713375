// newton.cu — the inner Newton / Laplace driver: host control flow around the device kernels.
//
// Replaces TMB's `ff$fn(theta)` for MakeADFun(random = "W") (call site
// /root/reference/R/02_model_fit.R:276-284; algorithm SURVEY.md Appendix A.1): minimise
// f(., theta) over W by damped Newton from the previous mode (TMB warm start), then
//   value = f(w_hat, theta) + 1/2 logdet H(w_hat, theta) - p/2 log(2 pi).
// Convergence mirrors TMB newton(): max|g| < grad.tol or max|step| < step.tol (both 1e-8),
// maxit 100; a step is accepted when the objective is finite and either it or max|g| decreased.
// One device->host read of 80 bytes of scalars per Newton iteration is the only synchronisation.
#include <algorithm>
#include <cstring>

#include <time.h>

#include <thread>

#include "bgp_internal.h"

namespace bgp {

__global__ void axpy_trial_kernel(const double* __restrict__ W, const double* __restrict__ step, double t, int p, int lda,
                                  double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < lda) out[i] = i < p ? fma(t, step[i], W[i]) : 0.0;
}

struct PredictArgs {
  const double* Wmode;
  const double* Tan;
  int S, lda;
  double dtheta[17];
  double* out;
};
// first-order predictor of the next mode: W0 = w_hat(theta_last) + sum_k T_k (theta_k - theta_last_k)
__global__ void predict_start_kernel(const PredictArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.lda) return;
  double v = a.Wmode[i];
  for (int k = 0; k < a.S; ++k) v = fma(a.Tan[(size_t)k * a.lda + i], a.dtheta[k], v);
  a.out[i] = v;
}

struct HermiteArgs {
  const double *Wa, *Ta, *Wb, *Tb;   // a: the evaluation before the last, b: the last one
  int S, lda;
  double u[17];                      // theta_b - theta_a
  double h00, h10, h01, h11;         // cubic Hermite basis at s = 1 + tau
  double* out;
};
// cubic Hermite extrapolation of the mode along the line theta_a -> theta_b -> theta_c (values and
// directional tangents at a and b): O(h^4) start instead of the O(h^2) of the tangent predictor
__global__ void hermite_start_kernel(const HermiteArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.lda) return;
  double ma = 0.0, mb = 0.0;
  for (int k = 0; k < a.S; ++k) {
    ma = fma(a.Ta[(size_t)k * a.lda + i], a.u[k], ma);
    mb = fma(a.Tb[(size_t)k * a.lda + i], a.u[k], mb);
  }
  a.out[i] = a.h00 * a.Wa[i] + a.h10 * ma + a.h01 * a.Wb[i] + a.h11 * mb;
}

static inline double tau_of(const bgp_model* m, const double* theta) {
  return m->family == BGP_FAMILY_GAUSSIAN ? std::exp(theta[m->S - 1]) : 1.0;
}

// BGP_HOST_TRACE=1: where the host spends an evaluation (enqueue / waiting for the device / bookkeeping), printed per batch
struct HostTrace {
  bool on = getenv("BGP_HOST_TRACE") != nullptr;
  double t_sync = 0, t_inner = 0, t_batch = 0;
  int n_sync = 0, n_inner = 0;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
  }
};
static HostTrace g_trace;

static int read_scalars(bgp_model* m, EvalScalars* out) {
  BGP_CUDA(cudaMemcpyAsync(m->sc_host, m->sc_dev, sizeof(EvalScalars), cudaMemcpyDeviceToHost, m->stream));
  phase_mark(m, PH_OTHER);
  if (m->host_hook) {                       // deferred host work of the previous evaluation: the GPU is busy now
    m->host_hook();
    m->host_hook = nullptr;
  }
  phase_collect(m);                         // likewise the phase timers of the previous segment
  const double t0 = g_trace.on ? HostTrace::now() : 0.0;
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  if (g_trace.on) {
    g_trace.t_sync += HostTrace::now() - t0;
    g_trace.n_sync++;
  }
  phase_harvest(m);
  *out = *m->sc_host;
  return BGP_OK;
}

// the scalars of the speculative first iteration: [1] = starting point (snapshot), [0] = Cholesky + trial point
static int read_scalars2(bgp_model* m, EvalScalars* start, EvalScalars* out) {
  BGP_CUDA(cudaMemcpyAsync(m->sc_host, m->sc_dev, 2 * sizeof(EvalScalars), cudaMemcpyDeviceToHost, m->stream));
  phase_mark(m, PH_OTHER);
  if (m->host_hook) {                       // deferred host work of the previous evaluation: the GPU is busy now
    m->host_hook();
    m->host_hook = nullptr;
  }
  phase_collect(m);                         // likewise the phase timers of the previous segment
  const double t0 = g_trace.on ? HostTrace::now() : 0.0;
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  const double t1 = g_trace.on ? HostTrace::now() : 0.0;
  phase_harvest(m);
  if (g_trace.on) {
    g_trace.t_sync += t1 - t0;
    g_trace.n_sync++;
    g_trace.t_batch += HostTrace::now() - t1;      // harvest share (reported separately below)
  }
  *out = m->sc_host[0];
  *start = m->sc_host[1];
  return BGP_OK;
}

// f, g (device), gmax at W_dev; leaves eta / wobs (/ c3) of that point on the device
int eval_fg_async(bgp_model* m, const double* W_dev, const double* theta, bool want_c3) {
  const double tau = tau_of(m, theta);
  m->obs_at_mode = false;                 // callers that evaluate at the mode set it again
  phase_mark(m, PH_LIK);
  if (m->osp_on) {
    // moment path (ospline.cu): on a single device the prior completion rides in its last kernel
    BGP_TRY(osp_launch_lik(m, W_dev, tau, m->world == 1 ? theta : nullptr));
    m->n_lik++;
    if (m->world > 1) BGP_TRY(launch_finish(m, W_dev, theta, tau));
    phase_mark(m, PH_OTHER);
    return BGP_OK;
  }
  BGP_TRY(launch_lik(m, W_dev, want_c3, tau));
  m->n_lik++;
  BGP_TRY(launch_finish(m, W_dev, theta, tau));
  phase_mark(m, PH_OTHER);
  return BGP_OK;
}

__global__ void commit_mode_kernel(const double* __restrict__ W, const double* __restrict__ Tan, int lda, int S,
                                   double* __restrict__ Wmode, double* __restrict__ hist_W, double* __restrict__ hist_T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < lda) {
    const double v = W[i];
    Wmode[i] = v;
    if (hist_W) hist_W[i] = v;
  } else if (i < lda * (1 + S)) {
    hist_T[i - lda] = Tan[i - lda];
  }
}

// One node's results into the fit's device slots (internal order) and, rotated to the external order of the ABI
// (io.cu), into a staging buffer [H (p x p) | mode (p)] for the copy to the host: one launch.
struct SinkArgs {
  const double *Wmode, *H;
  int p, nD, lda, ldh;
  double *modes_slot, *Hs_slot, *stage;
};
__global__ void sink_node_kernel(const SinkArgs a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  const int nU = a.p - a.nD;
  if (c == a.p) {                                  // the extra row of CTAs: the mode
    if (r < a.lda && a.modes_slot) a.modes_slot[r] = a.Wmode[r];
    if (r < a.p && a.stage) a.stage[(size_t)a.p * a.p + r] = a.Wmode[r < nU ? r + a.nD : r - nU];
    return;
  }
  if (r < a.ldh && a.Hs_slot) a.Hs_slot[(size_t)c * a.ldh + r] = a.H[(size_t)c * a.ldh + r];
  if (r < a.p && a.stage) {
    const int ir = r < nU ? r + a.nD : r - nU, ic = c < nU ? c + a.nD : c - nU;
    a.stage[(size_t)c * a.p + r] = a.H[(size_t)ic * a.ldh + ir];
  }
}

__global__ void sink_modes_kernel(const SinkArgs a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int nU = a.p - a.nD;
  if (r < a.lda && a.modes_slot) a.modes_slot[r] = a.Wmode[r];
  if (r < a.p && a.stage) a.stage[(size_t)a.p * a.p + r] = a.Wmode[r < nU ? r + a.nD : r - nU];
}

int laplace_inner(bgp_model* m, const double* theta, double* value, int* iters_out) {
  EvalScalars sc;
  // c3 = d w / d eta rides along with every pass (one more 8-byte store per observation): a gradient call at this
  // theta then finds w, c3 and the scalars of the mode on the device and needs no pass of its own
  const bool c3w = m->family == BGP_FAMILY_POISSON || m->family == BGP_FAMILY_BINOMIAL;
  bool pass_at_W = false;                 // the last likelihood pass was evaluated at m->W
  const int threads = 256, blocks = (m->lda + threads - 1) / threads;
  // start from the previous mode (TMB last.par.best), moved along the tangent d w_hat / d theta when the
  // step in theta is moderate; fall back to the plain warm start, then to W = 0, if that point is non-finite
  bool predicted = false;
  double start_dist = INFINITY;          // max-norm distance in theta to the history entry the start came from
  if (m->use_predictor && m->S <= 17) {
    // nearest history entry (max-norm in theta)
    int e1 = -1;
    double d1 = 0.0;
    for (int i = 0; i < bgp_model::NHIST; ++i) {
      const auto& h = m->hist[i];
      if (!h.stamp) continue;
      double d = 0.0;
      for (int k = 0; k < m->S; ++k) d = std::max(d, std::fabs(theta[k] - h.theta[k]));
      if (e1 < 0 || d < d1 || (d == d1 && h.stamp > m->hist[e1].stamp)) {
        e1 = i;
        d1 = d;
      }
    }
    if (e1 >= 0 && d1 <= 2.0) {
      start_dist = d1;
      const auto& hb = m->hist[e1];
      // second entry: collinear with (theta, e1), the closest such to e1; theta = theta_b + tau (theta_b - theta_a)
      int e2 = -1;
      double tau2 = 0.0, u2 = 0.0;
      if (m->use_hermite && d1 > 0.0) {
        for (int i = 0; i < bgp_model::NHIST; ++i) {
          const auto& ha = m->hist[i];
          if (!ha.stamp || i == e1) continue;
          double uu = 0.0, ud = 0.0, umax = 0.0;
          for (int k = 0; k < m->S; ++k) {
            const double u = hb.theta[k] - ha.theta[k];
            uu += u * u;
            ud += u * (theta[k] - hb.theta[k]);
            umax = std::max(umax, std::fabs(u));
          }
          if (!(uu > 0.0)) continue;
          const double tau = ud / uu;
          double dev = 0.0;
          for (int k = 0; k < m->S; ++k)
            dev = std::max(dev, std::fabs((theta[k] - hb.theta[k]) - tau * (hb.theta[k] - ha.theta[k])));
          // inside [a, b] (tau in (-1, 0)) or at most two intervals beyond b
          if (dev <= 1e-12 * std::max(1.0, d1) && tau > -1.0 && tau <= 2.0 && (e2 < 0 || umax < u2)) {
            e2 = i;
            tau2 = tau;
            u2 = umax;
          }
        }
      }
      if (e2 >= 0) {
        const auto& hA = m->hist[e2];
        HermiteArgs ha;
        ha.Wa = hA.W;
        ha.Ta = hA.T;
        ha.Wb = hb.W;
        ha.Tb = hb.T;
        ha.S = m->S;
        ha.lda = m->lda;
        for (int k = 0; k < m->S; ++k) ha.u[k] = hb.theta[k] - hA.theta[k];
        const double s = 1.0 + tau2, s2 = s * s, s3 = s2 * s;
        ha.h00 = 2.0 * s3 - 3.0 * s2 + 1.0;
        ha.h10 = s3 - 2.0 * s2 + s;
        ha.h01 = -2.0 * s3 + 3.0 * s2;
        ha.h11 = s3 - s2;
        ha.out = m->W;
        hermite_start_kernel<<<blocks, threads, 0, m->stream>>>(ha);
        count_launch();
        predicted = true;
      } else {
        PredictArgs pa;
        for (int k = 0; k < m->S; ++k) pa.dtheta[k] = theta[k] - hb.theta[k];
        pa.Wmode = hb.W;
        pa.Tan = hb.T;
        pa.S = m->S;
        pa.lda = m->lda;
        pa.out = m->W;
        predict_start_kernel<<<blocks, threads, 0, m->stream>>>(pa);
        count_launch();
        predicted = true;
      }
    }
  }
  if (!predicted)
    BGP_CUDA(cudaMemcpyAsync(m->W, m->Wmode, (size_t)m->lda * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
  BGP_TRY(eval_fg_async(m, m->W, theta, c3w));
  pass_at_W = true;
  // Speculation: a start predicted from a different theta is practically never converged or non-finite, so the
  // first Newton iteration is enqueued behind it at once and the starting point's scalars are read together with
  // the iteration's (one host round trip less; a wrong guess costs one likelihood pass or is redone below).
  bool spec = m->speculate && predicted && start_dist > 0.0;
  if (spec) {
    BGP_CUDA(cudaMemcpyAsync(m->sc_dev + 1, m->sc_dev, sizeof(EvalScalars), cudaMemcpyDeviceToDevice, m->stream));
    sc.nonfinite = 0;
    sc.f = sc.gmax = NAN;
  } else {
    BGP_TRY(read_scalars(m, &sc));
  }
  if (sc.nonfinite && predicted) {
    BGP_CUDA(cudaMemcpyAsync(m->W, m->Wmode, (size_t)m->lda * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
    BGP_TRY(eval_fg_async(m, m->W, theta, c3w));
    pass_at_W = true;
    BGP_TRY(read_scalars(m, &sc));
  }
  if (sc.nonfinite) {
    BGP_CUDA(cudaMemsetAsync(m->W, 0, (size_t)m->lda * sizeof(double), m->stream));
    BGP_TRY(eval_fg_async(m, m->W, theta, c3w));
    pass_at_W = true;
    BGP_TRY(read_scalars(m, &sc));
    if (sc.nonfinite) {
      set_error("objective is not finite at the starting point");
      *value = NAN;
      return BGP_ERR_NONFINITE;
    }
  }
  double f = sc.f, gmax = sc.gmax;
  int iters = 0;
  bool converged = false, have_factor = false, reused = false;
  double logdet = NAN;
  // Stagnation: TMB's newton() runs to maxit = 100 and answers NaN when max|g| never reaches the tolerance (a theta
  // far outside the posterior, e.g. exp(theta) = 1e-29: the gradient's rounding noise sits above 1e-8).  Twelve
  // consecutive iterations that improve neither the best max|g| by 10 % nor f end the same way ~85 Hessians earlier.
  double best_gmax = INFINITY, best_f = INFINITY;
  int since_progress = 0;
  for (int it = 0; it < m->maxit; ++it) {
    const bool spec_now = spec && it == 0;
    if (!spec_now && gmax < m->grad_tol) {
      converged = true;
      break;
    }
    if (!spec_now) {
      if (gmax < 0.9 * best_gmax || f < best_f - 1e-12 * std::fabs(best_f)) {
        best_gmax = std::min(best_gmax, gmax);
        best_f = std::min(best_f, f);
        since_progress = 0;
      } else if (++since_progress >= 12) {
        break;                                   // not converged: NaN, as after maxit iterations
      }
    }
    phase_mark(m, PH_HESS);
    BGP_TRY(launch_hessian(m, theta));
    m->n_hess++;
    phase_mark(m, PH_CHOL);
    const bool tan_in_chol = m->use_predictor && m->S <= CHOL_TANGENT_MAX_S;
    // speculative full step: the solve writes the trial point W + step itself; it is evaluated before anything is read back
    BGP_TRY(launch_chol_solve(m, true, tan_in_chol ? theta : nullptr, m->W, true));
    m->n_chol++;
    phase_mark(m, PH_OTHER);
    // the Cholesky scalars must be captured before the trial evaluation overwrites f / gmax:
    // they live in different fields of EvalScalars, so one read after the trial eval suffices.
    BGP_TRY(eval_fg_async(m, m->Wtrial, theta, c3w));
    pass_at_W = false;
    if (spec_now) {
      EvalScalars sc0;
      BGP_TRY(read_scalars2(m, &sc0, &sc));
      spec = false;
      if (sc0.nonfinite) {
        // the predicted start was not finite: what was enqueued behind it is void; redo from the plain warm start
        BGP_CUDA(cudaMemcpyAsync(m->W, m->Wmode, (size_t)m->lda * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
        BGP_TRY(eval_fg_async(m, m->W, theta, c3w));
        pass_at_W = true;
        BGP_TRY(read_scalars(m, &sc));
        if (sc.nonfinite) {
          BGP_CUDA(cudaMemsetAsync(m->W, 0, (size_t)m->lda * sizeof(double), m->stream));
          BGP_TRY(eval_fg_async(m, m->W, theta, c3w));
          pass_at_W = true;
          BGP_TRY(read_scalars(m, &sc));
          if (sc.nonfinite) {
            set_error("objective is not finite at the starting point");
            *value = NAN;
            return BGP_ERR_NONFINITE;
          }
        }
        f = sc.f;
        gmax = sc.gmax;
        it = -1;
        continue;
      }
      f = sc0.f;
      gmax = sc0.gmax;
      if (gmax < m->grad_tol && sc.chol_info == 0) {
        // already converged at the start: H, L and logdet were formed there (the trial evaluation was not needed;
        // eta / wobs on the device are those of W + step, closer to the mode still)
        converged = true;
        have_factor = true;
        logdet = sc.logdet;
        break;
      }
    } else {
      BGP_TRY(read_scalars(m, &sc));
    }
    if (sc.chol_info != 0) {
      set_error("Hessian not positive definite (pivot %d) at Newton iteration %d", sc.chol_info, it);
      *value = NAN;
      *iters_out = iters;
      return BGP_ERR_NOT_PD;
    }
    if (sc.smax < m->step_tol) {
      // W is at the precision limit: keep W (H, L, logdet were computed at W)
      converged = true;
      have_factor = true;
      logdet = sc.logdet;
      // the trial evaluation overwrote eta/wobs/g with those of W + step; they differ from W's by
      // less than step_tol and are not used again for this theta.
      break;
    }
    const double deta_full = sc.pad;      // max |eta(W + step) - eta(W)|: W is where H was formed
    static const bool newton_debug = getenv("BGP_NEWTON_DEBUG") != nullptr;     // env: diagnostics
    if (newton_debug)
      fprintf(stderr, "[newton] it %d  gmax %.3e -> %.3e  max|step| %.3e  max|d eta| %.3e  f %.15g -> %.15g (%.3e)\n", it,
              gmax, sc.gmax, sc.smax, deta_full, f, sc.f, sc.f - f);
    double t = 1.0;
    bool accepted = false, full_step = true;
    for (int h = 0; h < 40; ++h) {
      if (!sc.nonfinite && (sc.f <= f || sc.gmax < gmax)) {
        accepted = true;
        break;
      }
      t *= 0.5;
      full_step = false;
      axpy_trial_kernel<<<blocks, threads, 0, m->stream>>>(m->W, m->step, t, m->p, m->lda, m->Wtrial);
      count_launch();
      BGP_TRY(eval_fg_async(m, m->Wtrial, theta, c3w));
      pass_at_W = false;
      BGP_TRY(read_scalars(m, &sc));
    }
    if (!accepted) break;
    std::swap(m->W, m->Wtrial);
    pass_at_W = true;
    f = sc.f;
    gmax = sc.gmax;
    ++iters;
    if (gmax < m->grad_tol && full_step && m->allow_reuse && std::isfinite(sc.logdet)) {
      // Converged by a full Newton step from the point w0 where H was factored.  With delta = max_i |eta_i(w1) -
      // eta_i(w0)| the observation weights obey w_i(w1) = w_i(w0) e^{e_i}, |e_i| <= delta (Poisson: log w = eta;
      // Binomial: d log w / d eta = 1 - 2 pi in (-1, 1); Gaussian: w constant), hence in the Loewner order
      //     e^{-delta} H(w0)  <=  H(w1)  <=  e^{delta} H(w0)          (Q >= 0 only helps)
      // and EXACTLY (no linearisation)  |logdet H(w1) - logdet H(w0)| <= p * delta,
      //                                 ||H(w1) - H(w0)||_2 <= (e^{delta} - 1) ||H(w0)||_2.
      // Sharded models carry the SUM over ranks of the local maxima, an upper bound every rank sees identically.
      // When p * delta / 2 is far inside the 1e-8 tolerance of the value (and delta <= 1e-7 far inside the 1e-6
      // tolerance of the Hessian) the factor of the last iteration serves as the factor at the mode; the H left
      // on the device — what the caller receives as spHess — is then H(w0), within (e^{delta} - 1) of H(w1).
      const double val = f + 0.5 * sc.logdet - 0.5 * (double)m->p * std::log(2.0 * M_PI);
      const bool tiny = m->family == BGP_FAMILY_GAUSSIAN ||
                        (deta_full <= m->reuse_eta_tol && 0.5 * m->p * deta_full <= m->reuse_rel_tol * std::fabs(val));
      if (tiny) {
        converged = true;
        have_factor = true;
        reused = true;
        logdet = sc.logdet;
        break;
      }
    }
  }
  *iters_out = iters;
  if (!converged) {
    set_error("inner Newton did not converge (%d iterations, max|g| = %.3e)", iters, gmax);
    *value = NAN;
    return BGP_ERR_NO_CONVERGENCE;
  }
  if (!have_factor) {
    // Hessian and factor at the mode (eta / wobs on the device belong to W)
    phase_mark(m, PH_HESS);
    BGP_TRY(launch_hessian(m, theta));
    m->n_hess++;
    phase_mark(m, PH_CHOL);
    const bool tan_in_chol = m->use_predictor && m->S <= CHOL_TANGENT_MAX_S;
    // Rank 0 of the cluster carries a right-hand side here too (the gradient at the mode; its solution is not used).
    // The variant without one, next to ranks that do carry tangents, reported a non-positive pivot on a positive
    // definite matrix in ~1 of 4 runs of the whole GPU suite (Gaussian, p = 601, 16-wide panels, four tangents; never
    // in isolation): with BGP_PIVOT_DEBUG the inputs were verified sane, a second attempt of the same variant failed
    // the same way and the solve variant — the one every Newton iteration runs — succeeded on the re-formed matrix
    // (DESIGN.md section 8).  The kernel-side cause was not found; the variant is no longer used with tangents.
    BGP_TRY(launch_chol_solve(m, tan_in_chol, tan_in_chol ? theta : nullptr, m->W));
    m->n_chol++;
    BGP_TRY(read_scalars(m, &sc));
    if (sc.chol_info != 0) {
      // a non-positive pivot at a point reached through factorisations of (numerically) the same matrix is suspect:
      // form H and factor once more before answering NaN; counted and reported on stderr so it cannot hide
      const int first_pivot = sc.chol_info;
      if (getenv("BGP_PIVOT_DEBUG")) {
        // diagnostics: what went into the factorisation
        std::vector<double> hd((size_t)m->p), wv((size_t)m->n), hcol((size_t)m->p), wm((size_t)m->lda);
        cudaMemcpy2D(hd.data(), sizeof(double), m->H, (size_t)(m->ldh + 1) * sizeof(double), sizeof(double), (size_t)m->p,
                     cudaMemcpyDeviceToHost);
        cudaMemcpy(wv.data(), m->wobs, (size_t)m->n * sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(hcol.data(), m->H + (size_t)(first_pivot - 1) * m->ldh, (size_t)m->p * sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(wm.data(), m->W, (size_t)m->lda * sizeof(double), cudaMemcpyDeviceToHost);
        double wmin = 1e300, wmax = -1e300, dmin = 1e300, cmax = 0, wabs = 0;
        int wbad = 0, dbad = 0, cbad = 0;
        for (double v : wv) {
          if (!std::isfinite(v)) ++wbad;
          wmin = std::min(wmin, v);
          wmax = std::max(wmax, v);
        }
        for (double v : hd) {
          if (!std::isfinite(v)) ++dbad;
          dmin = std::min(dmin, v);
        }
        for (double v : hcol) {
          if (!std::isfinite(v)) ++cbad;
          cmax = std::max(cmax, std::fabs(v));
        }
        for (int i = 0; i < m->p; ++i) wabs = std::max(wabs, std::fabs(wm[(size_t)i]));
        fprintf(stderr, "[bgp] pivot debug: wobs min %.6g max %.6g nonfinite %d | diag(H) min %.6g nonfinite %d H[%d][%d] %.6g | column max %.6g nonfinite %d | max|W| %.6g | theta",
                wmin, wmax, wbad, dmin, dbad, first_pivot - 1, first_pivot - 1, hd[(size_t)first_pivot - 1], cmax, cbad, wabs);
        for (int k = 0; k < m->S; ++k) fprintf(stderr, " %.6g", theta[k]);
        fprintf(stderr, "\n");
      }
      phase_mark(m, PH_HESS);
      BGP_TRY(launch_hessian(m, theta));
      m->n_hess++;
      phase_mark(m, PH_CHOL);
      BGP_TRY(launch_chol_solve(m, true, tan_in_chol ? theta : nullptr, m->W));      // the variant the Newton iterations run
      m->n_chol++;
      BGP_TRY(read_scalars(m, &sc));
      ++m->n_refactor;
      fprintf(stderr, "[bgp] factorisation at the mode reported pivot %d; second attempt: %s (p = %d, attempt %lld of this model)\n",
              first_pivot, sc.chol_info == 0 ? "positive definite" : "same failure", m->p, (long long)m->n_refactor);
    }
    if (sc.chol_info != 0) {
      set_error("Hessian not positive definite at the mode (pivot %d)", sc.chol_info);
      *value = NAN;
      return BGP_ERR_NOT_PD;
    }
    logdet = sc.logdet;
  }
  double *hist_W = nullptr, *hist_T = nullptr;
  if (m->use_predictor && m->S <= 17) {
    // the tangent came out of the last Cholesky launch (idle cluster ranks) unless there are too many thetas
    if (m->S > CHOL_TANGENT_MAX_S) BGP_TRY(launch_tangent(m, theta));
    m->theta_last.assign(theta, theta + m->S);
    m->tan_valid = true;
    // record (theta, mode, tangent): same theta => overwrite, else replace the oldest entry
    int slot = -1;
    for (int i = 0; i < bgp_model::NHIST && slot < 0; ++i)
      if (m->hist[i].stamp && m->hist[i].theta == m->theta_last) slot = i;
    if (slot < 0) {
      slot = 0;
      for (int i = 1; i < bgp_model::NHIST; ++i)
        if (m->hist[i].stamp < m->hist[slot].stamp) slot = i;
    }
    auto& h = m->hist[slot];
    h.theta = m->theta_last;
    h.stamp = ++m->hist_clock;
    hist_W = h.W;
    hist_T = h.T;
  }
  // the mode into Wmode and, with its tangent, into the history slot: one launch
  commit_mode_kernel<<<(m->lda * (1 + (hist_T ? m->S : 0)) + 255) / 256, 256, 0, m->stream>>>(m->W, m->Tan, m->lda, hist_T ? m->S : 0,
                                                                                            m->Wmode, hist_W, hist_T);
  count_launch();
  m->factor_is_exact = !reused;
  m->obs_at_mode = pass_at_W;
  ++m->n_evals;
  m->n_newton += iters;
  m->n_reuse += reused ? 1 : 0;
  *value = f + 0.5 * logdet - 0.5 * (double)m->p * std::log(2.0 * M_PI);
  return BGP_OK;
}

// K Laplace evaluations (theta: S x K, node j at theta + j*S) of which this call runs those with mine[j] != 0
// (NULL: all).  Evaluation order (results go back to slot j): start next to what is already known — the nearest
// history entry, or the node closest to the centroid when the history is empty (the warm start is the mode at the
// grid centre in aghq's flow) — then always the node nearest to an evaluated one, so every inner solve starts from a
// close, usually collinear, pair of neighbours.  Modes / Hessians leave either through two pinned slots (host sink:
// the device -> pinned copy is asynchronous, the pinned -> caller copy runs while the next evaluation's kernels are
// in flight) or by device-to-device copies (device sink: nothing crosses PCIe).  Nodes that fail get NaN and the
// worst status is returned after all nodes were tried; values of nodes that are not mine are left untouched.
static int laplace_batch_serial(bgp_model* m, int K, const double* theta, const unsigned char* mine, double* values,
                                const BatchSink& sink, int* iters_total, int* first_failed) {
  int total = 0, worst = BGP_OK;
  if (first_failed) *first_failed = -1;
  // the deferred copies point into the caller's arrays: none may outlive this call, whichever way it returns
  struct HookGuard {
    bgp_model* m;
    ~HookGuard() {
      m->host_hook = nullptr;
      if (m->out_stream) cudaStreamSynchronize(m->out_stream);   // copies into the caller's arrays: none in flight
    }
  } hook_guard{m};
  m->host_hook = nullptr;
  const int S = m->S;
  std::vector<int> order;
  {
    std::vector<char> done((size_t)K, 0);
    int todo = 0;
    for (int j = 0; j < K; ++j) {
      if (mine && !mine[j]) done[j] = 1;
      else ++todo;
    }
    std::vector<double> known;                          // thetas with a mode on record
    for (const auto& h : m->hist)
      if (h.stamp && (int)h.theta.size() == S) known.insert(known.end(), h.theta.begin(), h.theta.end());
    if (known.empty()) {
      std::vector<double> c((size_t)S, 0.0);
      for (int j = 0; j < K; ++j)
        if (!done[j])
          for (int k = 0; k < S; ++k) c[k] += theta[(size_t)j * S + k] / todo;
      known = c;
    }
    // distance of every open node to the nearest known theta, updated as nodes are taken (O(K^2 S) in all)
    std::vector<double> dist((size_t)K, INFINITY);
    auto relax = [&](const double* t) {
      for (int j = 0; j < K; ++j) {
        if (done[j]) continue;
        double d = 0.0;
        for (int k = 0; k < S; ++k) {
          const double u = theta[(size_t)j * S + k] - t[k];
          d += u * u;
        }
        dist[j] = std::min(dist[j], d);
      }
    };
    for (size_t e = 0; e + S <= known.size(); e += S) relax(&known[e]);
    for (int it = 0; it < todo; ++it) {
      int best = -1;
      for (int j = 0; j < K; ++j)
        if (!done[j] && (best < 0 || dist[j] < dist[best])) best = j;
      done[best] = 1;
      order.push_back(best);
      relax(theta + (size_t)best * S);
    }
  }
  const size_t pp = (size_t)m->p * m->p;
  for (size_t oi = 0; oi < order.size(); ++oi) {
    const int j = order[oi];
    int iters = 0;
    double v = NAN;
    const double tt0 = g_trace.on ? HostTrace::now() : 0.0;
    int st = laplace_inner(m, theta + (size_t)j * m->S, &v, &iters);
    if (g_trace.on) {
      g_trace.t_inner += HostTrace::now() - tt0;
      g_trace.n_inner++;
    }
    total += iters;
    values[j] = st == BGP_OK ? v : NAN;
    if (st == BGP_ERR_CUDA || st == BGP_ERR_NCCL) return st;
    if (st != BGP_OK) {
      if (worst == BGP_OK && first_failed) *first_failed = j;
      worst = st;
      continue;
    }
    const bool to_pinned = (sink.modes_host || sink.Hs_host) && sink.host_pinned;
    if (sink.modes_dev || sink.Hs_dev || to_pinned) {
      const size_t slot = sink.dev_slot ? (size_t)sink.dev_slot[j] : (size_t)j;
      SinkArgs sa;
      sa.Wmode = m->Wmode;
      sa.H = m->H;
      sa.p = m->p;
      sa.nD = m->nD;
      sa.lda = m->lda;
      sa.ldh = m->ldh;
      sa.modes_slot = sink.modes_dev ? sink.modes_dev + slot * m->lda : nullptr;
      sa.Hs_slot = sink.Hs_dev ? sink.Hs_dev + slot * (size_t)m->p * m->ldh : nullptr;
      sa.stage = nullptr;
      int b = 0;
      if (to_pinned) {
        // page-locked destination: rotate into one of two staging buffers here, copy to slot j on the output stream
        if (!m->out_stream) {
          BGP_CUDA(cudaStreamCreateWithFlags(&m->out_stream, cudaStreamNonBlocking));
          for (int i = 0; i < 2; ++i) {
            BGP_CUDA(cudaMalloc(&m->out_stage[i], (pp + (size_t)m->p) * sizeof(double)));
            BGP_CUDA(cudaEventCreateWithFlags(&m->out_ready[i], cudaEventDisableTiming));
            BGP_CUDA(cudaEventCreateWithFlags(&m->out_done[i], cudaEventDisableTiming));
          }
        }
        b = (int)(m->out_count++ & 1u);
        if (m->out_used[b]) BGP_CUDA(cudaStreamWaitEvent(m->stream, m->out_done[b], 0));   // its previous tenant has left
        sa.stage = m->out_stage[b];
      }
      const bool want_H = sa.Hs_slot || (to_pinned && sink.Hs_host);
      dim3 grid((std::max(m->lda, m->ldh) + 255) / 256, m->p + 1);
      if (!want_H) {                               // modes only: just the extra row
        sa.H = nullptr;
        grid = dim3((m->lda + 255) / 256, 1);
        sink_modes_kernel<<<grid, 256, 0, m->stream>>>(sa);
      } else {
        sink_node_kernel<<<grid, 256, 0, m->stream>>>(sa);
      }
      count_launch();
      BGP_CUDA(cudaGetLastError());
      if (to_pinned) {
        BGP_CUDA(cudaEventRecord(m->out_ready[b], m->stream));
        BGP_CUDA(cudaStreamWaitEvent(m->out_stream, m->out_ready[b], 0));
        if (sink.modes_host)
          BGP_CUDA(cudaMemcpyAsync(sink.modes_host + (size_t)j * m->p, m->out_stage[b] + pp, (size_t)m->p * sizeof(double),
                                   cudaMemcpyDeviceToHost, m->out_stream));
        if (sink.Hs_host)
          BGP_CUDA(cudaMemcpyAsync(sink.Hs_host + (size_t)j * pp, m->out_stage[b], pp * sizeof(double), cudaMemcpyDeviceToHost,
                                   m->out_stream));
        BGP_CUDA(cudaEventRecord(m->out_done[b], m->out_stream));
        m->out_used[b] = true;
      }
    }
    if (to_pinned) {
      // nothing more: the copies are in flight on the output stream
    } else if (sink.modes_host || sink.Hs_host) {
      const size_t need = pp + (size_t)m->p;
      if (m->pin_out_elems < need) {
        for (int i = 0; i < 2; ++i) {
          if (m->pin_out[i]) cudaFreeHost(m->pin_out[i]);
          m->pin_out[i] = nullptr;
          BGP_CUDA(cudaMallocHost(&m->pin_out[i], need * sizeof(double)));
          if (!m->pin_ev[i]) BGP_CUDA(cudaEventCreateWithFlags(&m->pin_ev[i], cudaEventDisableTiming));
        }
        m->pin_out_elems = need;
      }
      if (m->host_hook) {                   // the slot's previous tenant (two evaluations ago at the latest)
        m->host_hook();
        m->host_hook = nullptr;
      }
      const int slot = (int)(oi & 1);
      double* pin = m->pin_out[slot];
      if (sink.modes_host) BGP_TRY(copy_vec_out(m, m->Wmode, pin + pp));
      if (sink.Hs_host) BGP_TRY(copy_H_out(m, pin));
      BGP_CUDA(cudaEventRecord(m->pin_ev[slot], m->stream));
      double* mode_dst = sink.modes_host ? sink.modes_host + (size_t)j * m->p : nullptr;
      double* H_dst = sink.Hs_host ? sink.Hs_host + (size_t)j * pp : nullptr;
      cudaEvent_t ev = m->pin_ev[slot];
      const size_t pn = (size_t)m->p;
      m->host_hook = [pin, pp, pn, mode_dst, H_dst, ev]() {
        cudaEventSynchronize(ev);
        if (mode_dst) memcpy(mode_dst, pin + pp, pn * sizeof(double));
        if (H_dst) memcpy(H_dst, pin, pp * sizeof(double));
      };
    }
  }
  if (m->host_hook) {
    m->host_hook();
    m->host_hook = nullptr;
  }
  if (m->out_stream) BGP_CUDA(cudaStreamSynchronize(m->out_stream));
  if (g_trace.on && g_trace.n_inner > 0) {
    fprintf(stderr, "[host] per evaluation: laplace_inner %.1f us of which waiting for the device %.1f us (%d syncs), event harvest %.1f us\n",
            g_trace.t_inner / g_trace.n_inner, g_trace.t_sync / g_trace.n_inner, g_trace.n_sync, g_trace.t_batch / g_trace.n_inner);
    g_trace = HostTrace();
  }
  if (iters_total) *iters_total = total;
  return worst;
}

// The nodes of a batch dealt to the model's evaluation lanes (bgp_model::n_lanes): contiguous runs of the caller's node
// order (neighbours share a lane, which keeps the warm starts close — the split of a node group, one level down), one
// host thread per lane.  Every lane starts from a copy of the model's warm-start history and evaluates its run with the
// serial driver above; results go to disjoint slots of the same sink.  Values differ from a one-lane run only through
// the starting points (well inside the tolerances the inner solve converges to).
int laplace_batch(bgp_model* m, int K, const double* theta, const unsigned char* mine, double* values,
                  const BatchSink& sink, int* iters_total, int* first_failed) {
  int count = 0;
  for (int j = 0; j < K; ++j) count += (!mine || mine[j]) ? 1 : 0;
  int L = m->n_lanes;
  if (!m->osp_on || m->world > 1 || m->is_lane) L = 1;
  L = std::min(L, count / 2);                       // a lane's first node costs two Newton iterations: at least two nodes each
  if (L <= 1) return laplace_batch_serial(m, K, theta, mine, values, sink, iters_total, first_failed);
  BGP_TRY(lanes_ensure(m, L));
  std::vector<bgp_model*> lane((size_t)L);
  lane[0] = m;
  for (int l = 1; l < L; ++l) lane[(size_t)l] = m->lanes[(size_t)l - 1];
  // the model's pending work (history entries, the mode of the last evaluation) must be complete before it is copied
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  for (int l = 1; l < L; ++l) {
    bgp_model* ln = lane[(size_t)l];
    ln->osp_on = true;
    ln->osp_dense_grad = m->osp_dense_grad;
    ln->use_predictor = m->use_predictor;
    ln->use_hermite = m->use_hermite;
    ln->speculate = m->speculate;
    ln->allow_reuse = m->allow_reuse;
    ln->reuse_eta_tol = m->reuse_eta_tol;
    ln->reuse_rel_tol = m->reuse_rel_tol;
    ln->grad_tol = m->grad_tol;
    ln->step_tol = m->step_tol;
    ln->maxit = m->maxit;
    ln->theta_last = m->theta_last;
    ln->tan_valid = m->tan_valid;
    ln->hist_clock = m->hist_clock;
    ln->factor_is_exact = false;
    ln->obs_at_mode = false;
    ln->L_holds_H = false;
    const size_t vb = (size_t)m->lda * sizeof(double);
    BGP_CUDA(cudaMemcpyAsync(ln->Wmode, m->Wmode, vb, cudaMemcpyDeviceToDevice, ln->stream));
    BGP_CUDA(cudaMemcpyAsync(ln->Tan, m->Tan, (size_t)std::max(1, m->S) * vb, cudaMemcpyDeviceToDevice, ln->stream));
    for (int i = 0; i < bgp_model::NHIST; ++i) {
      ln->hist[i].theta = m->hist[i].theta;
      ln->hist[i].stamp = m->hist[i].stamp;
      if (!m->hist[i].stamp) continue;
      BGP_CUDA(cudaMemcpyAsync(ln->hist[i].W, m->hist[i].W, vb, cudaMemcpyDeviceToDevice, ln->stream));
      BGP_CUDA(cudaMemcpyAsync(ln->hist[i].T, m->hist[i].T, (size_t)std::max(1, m->S) * vb, cudaMemcpyDeviceToDevice, ln->stream));
    }
  }
  for (int l = 1; l < L; ++l) BGP_CUDA(cudaStreamSynchronize(lane[(size_t)l]->stream));
  // contiguous, count-balanced runs of the nodes that are mine
  std::vector<std::vector<unsigned char>> mask((size_t)L, std::vector<unsigned char>((size_t)K, 0));
  {
    int idx = 0;
    for (int j = 0; j < K; ++j)
      if (!mine || mine[j]) mask[(size_t)piece_owner(idx++, count, L)][(size_t)j] = 1;
  }
  std::vector<int> rc((size_t)L, BGP_OK), its((size_t)L, 0), ff((size_t)L, -1);
  std::vector<std::string> err((size_t)L);
  auto run = [&](int l) {
    bgp_model* ln = lane[(size_t)l];
    if (cudaSetDevice(ln->device) != cudaSuccess) {
      rc[(size_t)l] = BGP_ERR_CUDA;
      err[(size_t)l] = "cudaSetDevice failed in an evaluation lane";
      return;
    }
    rc[(size_t)l] = laplace_batch_serial(ln, K, theta, mask[(size_t)l].data(), values, sink, &its[(size_t)l], &ff[(size_t)l]);
    if (rc[(size_t)l] != BGP_OK) err[(size_t)l] = g_last_error;
    cudaStreamSynchronize(ln->stream);
  };
  for (int l = 1; l < L; ++l) lane[(size_t)l]->worker->post([&run, l] { run(l); });
  run(0);
  for (int l = 1; l < L; ++l) lane[(size_t)l]->worker->wait();
  int total = 0, worst = BGP_OK, bad = -1;
  for (int l = 0; l < L; ++l) {
    total += its[(size_t)l];
    if (rc[(size_t)l] != BGP_OK && (worst == BGP_OK || rc[(size_t)l] == BGP_ERR_CUDA)) {
      worst = rc[(size_t)l];
      g_last_error = err[(size_t)l];
    }
    if (ff[(size_t)l] >= 0 && (bad < 0 || ff[(size_t)l] < bad)) bad = ff[(size_t)l];
    if (l > 0) {
      // the lanes' counters and phase timers belong to the model
      bgp_model* ln = lane[(size_t)l];
      phase_collect(ln);
      phase_harvest(ln);
      phase_collect(ln);
      m->t_lik += ln->t_lik;
      m->t_hess += ln->t_hess;
      m->t_chol += ln->t_chol;
      m->n_lik += ln->n_lik;
      m->n_hess += ln->n_hess;
      m->n_chol += ln->n_chol;
      m->n_evals += ln->n_evals;
      m->n_newton += ln->n_newton;
      m->n_reuse += ln->n_reuse;
      ln->t_lik = ln->t_hess = ln->t_chol = 0.0;
      ln->n_lik = ln->n_hess = ln->n_chol = 0;
      ln->n_evals = ln->n_newton = ln->n_reuse = 0;
    }
  }
  if (iters_total) *iters_total = total;
  if (first_failed) *first_failed = bad;
  return worst;
}

}  // namespace bgp

using namespace bgp;

#define BGP_CHECK_READY(m)                    \
  do {                                        \
    if (!(m) || !(m)->finalized) {            \
      set_error("model not finalized");       \
      return BGP_ERR_STATE;                   \
    }                                         \
    BGP_CUDA(cudaSetDevice((m)->device));     \
  } while (0)

extern "C" {

int bgp_objective(bgp_model* m, const double* W, const double* theta, double* f, double* grad, double* H) {
  BGP_CHECK_READY(m);
  if (!W || !theta) {
    set_error("bgp_objective: NULL W / theta");
    return BGP_ERR_ARG;
  }
  BGP_TRY(copy_vec_in(m, W, m->Wtrial));
  BGP_TRY(eval_fg_async(m, m->Wtrial, theta, false));
  EvalScalars sc;
  BGP_TRY(read_scalars(m, &sc));
  if (f) *f = sc.nonfinite ? NAN : sc.f;
  if (grad) BGP_TRY(copy_vec_out(m, m->g, grad));
  if (H) {
    BGP_TRY(launch_hessian(m, theta));
    BGP_TRY(copy_H_out(m, H));
  }
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  return BGP_OK;
}

int bgp_model_set_start(bgp_model* m, const double* W) {
  BGP_CHECK_READY(m);
  BGP_CUDA(cudaMemsetAsync(m->Wmode, 0, (size_t)m->lda * sizeof(double), m->stream));
  if (W) BGP_TRY(copy_vec_in(m, W, m->Wmode));
  m->tan_valid = false;
  for (auto& h : m->hist) h.stamp = 0;
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  return BGP_OK;
}

int bgp_model_get_tangent(bgp_model* m, double* theta, double* T) {
  BGP_CHECK_READY(m);
  if (!m->tan_valid || (int)m->theta_last.size() != m->S) {
    set_error("bgp_model_get_tangent: no Laplace evaluation on record (or the predictor is switched off)");
    return BGP_ERR_STATE;
  }
  if (theta) std::copy(m->theta_last.begin(), m->theta_last.end(), theta);
  if (T)
    for (int k = 0; k < m->S; ++k) BGP_TRY(copy_vec_out(m, m->Tan + (size_t)k * m->lda, T + (size_t)k * m->p));
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  return BGP_OK;
}

int bgp_model_set_start_at(bgp_model* m, const double* theta, const double* W, const double* T) {
  BGP_CHECK_READY(m);
  if (!theta || !W) {
    set_error("bgp_model_set_start_at: NULL theta / W");
    return BGP_ERR_ARG;
  }
  BGP_TRY(bgp_model_set_start(m, W));                    // clears the history, W becomes the warm start
  if (!T || !m->use_predictor || m->S > 17) return BGP_OK;
  auto& h = m->hist[0];
  h.theta.assign(theta, theta + m->S);
  h.stamp = ++m->hist_clock;
  BGP_CUDA(cudaMemcpyAsync(h.W, m->Wmode, (size_t)m->lda * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
  for (int k = 0; k < m->S; ++k) {
    BGP_TRY(copy_vec_in(m, T + (size_t)k * m->p, h.T + (size_t)k * m->lda));
    BGP_CUDA(cudaStreamSynchronize(m->stream));          // copy_vec_in stages through one buffer
  }
  BGP_CUDA(cudaMemcpyAsync(m->Tan, h.T, (size_t)m->S * m->lda * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
  m->theta_last = h.theta;
  m->tan_valid = true;
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  return BGP_OK;
}

int bgp_model_set_newton(bgp_model* m, double grad_tol, double step_tol, int maxit) {
  if (!m) return BGP_ERR_ARG;
  if (grad_tol > 0) m->grad_tol = grad_tol;
  if (step_tol > 0) m->step_tol = step_tol;
  if (maxit > 0) m->maxit = maxit;
  return BGP_OK;
}

int bgp_laplace_eval(bgp_model* m, const double* theta, double* value, double* grad, double* w_mode, double* H,
                     int* newton_iters) {
  BGP_CHECK_READY(m);
  if (!theta || !value) {
    set_error("bgp_laplace_eval: NULL theta / value");
    return BGP_ERR_ARG;
  }
  int iters = 0;
  cudaEventRecord(m->ev[0], m->stream);
  int st = laplace_inner(m, theta, value, &iters);
  if (newton_iters) *newton_iters = iters;
  if (st != BGP_OK) {
    // inner failure: NaN value so a vmmin-style caller can backtrack (R_FINITE test)
    *value = NAN;
    if (st == BGP_ERR_NOT_PD || st == BGP_ERR_NO_CONVERGENCE || st == BGP_ERR_NONFINITE) {
      cudaStreamSynchronize(m->stream);
      return st;
    }
    return st;
  }
  if (grad) {
    BGP_TRY(laplace_gradient(m, theta, grad));
  }
  if (w_mode) BGP_TRY(copy_vec_out(m, m->Wmode, w_mode));
  if (H) BGP_TRY(copy_H_out(m, H));
  cudaEventRecord(m->ev[1], m->stream);
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]);
  m->t_total = ms;
  return BGP_OK;
}

int bgp_laplace_eval_batch(bgp_model* m, int K, const double* theta, double* values, double* modes, double* Hs,
                           int* newton_iters_total) {
  BGP_CHECK_READY(m);
  if (K <= 0 || !theta || !values) {
    set_error("bgp_laplace_eval_batch: bad arguments");
    return BGP_ERR_ARG;
  }
  BatchSink sink;
  sink.modes_host = modes;
  sink.Hs_host = Hs;
  cudaEventRecord(m->ev[0], m->stream);
  int st = laplace_batch(m, K, theta, nullptr, values, sink, newton_iters_total, nullptr);
  cudaEventRecord(m->ev[1], m->stream);
  BGP_CUDA(cudaStreamSynchronize(m->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]);
  m->t_total = ms;
  return st;
}

int bgp_model_last_timing(const bgp_model* m, double* total_ms, double* lik_ms, double* hess_ms, double* chol_ms,
                          int64_t* lik_launches, int64_t* hess_launches, int64_t* chol_launches) {
  if (!m) return BGP_ERR_ARG;
  if (total_ms) *total_ms = m->t_total;
  phase_collect(const_cast<bgp_model*>(m));     // timers are a cache of the recorded events
  if (lik_ms) *lik_ms = m->t_lik;
  if (hess_ms) *hess_ms = m->t_hess;
  if (chol_ms) *chol_ms = m->t_chol;
  if (lik_launches) *lik_launches = m->n_lik;
  if (hess_launches) *hess_launches = m->n_hess;
  if (chol_launches) *chol_launches = m->n_chol;
  return BGP_OK;
}

int bgp_model_gradient_timing(const bgp_model* m, double* leverage_ms, int64_t* leverage_launches, double* leverage_flops,
                              double* dense_flops) {
  if (!m) return BGP_ERR_ARG;
  phase_collect(const_cast<bgp_model*>(m));
  if (leverage_ms) *leverage_ms = m->t_lev;
  if (leverage_launches) *leverage_launches = m->n_lev;
  if (leverage_flops) *leverage_flops = m->lev_flops;
  if (dense_flops) *dense_flops = (double)m->n * m->p * m->p;     // n p^2: the lower-triangular factor on dense rows
  return BGP_OK;
}

int bgp_model_counters(const bgp_model* m, int64_t* laplace_evals, int64_t* newton_iters, int64_t* factor_reuses) {
  if (!m) return BGP_ERR_ARG;
  if (laplace_evals) *laplace_evals = m->n_evals;
  if (newton_iters) *newton_iters = m->n_newton;
  if (factor_reuses) *factor_reuses = m->n_reuse;
  return BGP_OK;
}

int bgp_model_set_factor_reuse(bgp_model* m, int allow, double eta_tol, double rel_tol) {
  if (!m) return BGP_ERR_ARG;
  m->allow_reuse = allow != 0;
  if (eta_tol > 0) m->reuse_eta_tol = eta_tol;
  if (rel_tol > 0) m->reuse_rel_tol = rel_tol;
  return BGP_OK;
}

int bgp_model_hessian_flops(const bgp_model* m, double* dense, double* structural) {
  if (!m || !m->finalized) {
    set_error("model not finalized");
    return BGP_ERR_STATE;
  }
  if (dense) *dense = (double)m->n * m->p * (m->p + 1.0);
  if (structural) *structural = m->hess_useful_flops;
  return BGP_OK;
}

int bgp_model_lik_bytes(const bgp_model* m, double* dense, double* structural) {
  if (!m || !m->finalized) {
    set_error("model not finalized");
    return BGP_ERR_STATE;
  }
  const double vec = 24.0 * (double)m->n;            // y in, eta and w out
  if (dense) *dense = 8.0 * (double)m->n * m->lda + vec;
  if (structural) {
    const int nj = (m->lda + 63) / 64;
    double groups = 0.0;
    for (uint64_t o : m->occ_host)
      for (int j = 0; j < nj; ++j) groups += ((o >> (4 * j)) & 0xfull) ? 1.0 : 0.0;
    *structural = groups * 64.0 * 512.0 + vec;
  }
  return BGP_OK;
}

}  // extern "C"
