// fit.cu — the outer AGHQ inference: replaces aghq::marginal_laplace_tmb(ff, k, startingvalue)
// (call site /root/reference/R/02_model_fit.R:284) and the objects BayesGP reads from its result
// (SURVEY.md section 8b).  Host C++ drives the device Laplace objective:
//   optimize_theta  -> stats::optim(method = "BFGS") == vmmin()                   (Appendix A.2)
//   ff$he           -> numDeriv::jacobian(ff$gr, .), Richardson, R/02_model_fit.R:283 (A.3)
//   normalize_logpost -> mvQuad "GHe" product grid rescaled by chol(H^-1)          (A.4)
//   marginals ("reuse"), per-node modes and Hessians                               (A.5)
#include <algorithm>
#include <chrono>
#include <functional>
#include <limits>

#include <cstdlib>

#include "bgp_internal.h"

namespace bgp {

void fit_release_device(bgp_fit* f);

// ---- ff$fn / ff$gr with TMB's "same theta => reuse the inner optimum" behaviour ---------------------
struct FF {
  bgp_model* m;
  int n_fn = 0, n_gr = 0;
  std::vector<double> last_theta;
  bool have = false;
  double last_value = NAN;

  int ensure(const double* theta, double* value) {
    const int S = m->S;
    if (have && (int)last_theta.size() == S && std::equal(theta, theta + S, last_theta.begin())) {
      *value = last_value;
      return BGP_OK;
    }
    int iters = 0;
    const int64_t h0 = m->n_hess, l0 = m->n_lik;
    int st = laplace_inner(m, theta, value, &iters);
    static const bool dbg = getenv("BGP_FIT_DEBUG") != nullptr;     // env: diagnostics
    if (dbg) {
      fprintf(stderr, "[fit] ff(");
      for (int i = 0; i < S; ++i) fprintf(stderr, "%s%.6f", i ? ", " : "", theta[i]);
      fprintf(stderr, ") = %.10g  status %d  newton %d  hessians %lld  passes %lld\n", *value, st, iters,
              (long long)(m->n_hess - h0), (long long)(m->n_lik - l0));
    }
    if (st == BGP_ERR_CUDA || st == BGP_ERR_NCCL) return st;
    if (st != BGP_OK) {
      *value = NAN;          // inner failure => NaN, the optimiser backtracks (R_FINITE test)
      have = false;
      return BGP_OK;
    }
    last_theta.assign(theta, theta + S);
    last_value = *value;
    have = true;
    return BGP_OK;
  }
  int fn(const double* theta, double* value) {
    ++n_fn;
    return ensure(theta, value);
  }
  int gr(const double* theta, double* g) {
    ++n_gr;
    double v;
    BGP_TRY(ensure(theta, &v));
    if (!std::isfinite(v)) {
      for (int i = 0; i < m->S; ++i) g[i] = NAN;
      return BGP_OK;
    }
    return laplace_gradient(m, theta, g);
  }
};

// ---- stats::optim(method = "BFGS"): vmmin() of R's optim.c (A.2) ---------------------------------------
static int vmmin(FF& ff, std::vector<double>& b, double* Fmin_out, int* fail, int maxit = 100) {
  const double stepredn = 0.2, acctol = 1e-4, reltest = 10.0, abstol = -std::numeric_limits<double>::infinity();
  const double reltol = std::sqrt(std::numeric_limits<double>::epsilon());
  const int n = (int)b.size();
  std::vector<double> g(n), t(n), X(n), c(n), B((size_t)n * n, 0.0);
  double f;
  BGP_TRY(ff.fn(b.data(), &f));
  if (!std::isfinite(f)) {
    set_error("initial value in 'vmmin' is not finite");
    return BGP_ERR_NONFINITE;
  }
  double Fmin = f;
  int funcount = 1, gradcount = 1, iter = 1, count = 0;
  BGP_TRY(ff.gr(b.data(), g.data()));
  int ilast = gradcount;
  do {
    if (ilast == gradcount)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) B[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int i = 0; i < n; ++i) {
      X[i] = b[i];
      c[i] = g[i];
    }
    double gradproj = 0.0;
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s -= B[(size_t)i * n + j] * c[j];
      t[i] = s;
      gradproj += s * c[i];
    }
    if (gradproj < 0.0) {
      double steplength = 1.0;
      bool accpoint = false;
      do {
        count = 0;
        for (int i = 0; i < n; ++i) {
          b[i] = X[i] + steplength * t[i];
          if (reltest + X[i] == reltest + b[i]) ++count;
        }
        if (count < n) {
          BGP_TRY(ff.fn(b.data(), &f));
          ++funcount;
          accpoint = std::isfinite(f) && (f <= Fmin + gradproj * steplength * acctol);
          if (!accpoint) steplength *= stepredn;
        }
      } while (!(count == n || accpoint));
      const bool enough = (f > abstol) && std::fabs(f - Fmin) > reltol * (std::fabs(Fmin) + reltol);
      if (!enough) {
        count = n;
        Fmin = f;
      }
      if (count < n) {
        Fmin = f;
        BGP_TRY(ff.gr(b.data(), g.data()));
        ++gradcount;
        ++iter;
        double D1 = 0.0;
        for (int i = 0; i < n; ++i) {
          t[i] = steplength * t[i];
          c[i] = g[i] - c[i];
          D1 += t[i] * c[i];
        }
        if (D1 > 0) {
          double D2 = 0.0;
          for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s += B[(size_t)i * n + j] * c[j];
            X[i] = s;
            D2 += s * c[i];
          }
          D2 = 1.0 + D2 / D1;
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
              B[(size_t)i * n + j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1;
        } else {
          ilast = gradcount;
        }
      } else {
        if (ilast < gradcount) {
          count = 0;
          ilast = gradcount;
        }
      }
    } else {
      count = 0;
      if (ilast == gradcount) count = n;
      else ilast = gradcount;
    }
    if (iter >= maxit) break;
    if (gradcount - ilast > 2 * n) ilast = gradcount;
  } while (count != n || ilast != gradcount);
  *Fmin_out = Fmin;
  *fail = iter < maxit ? 0 : 1;
  return BGP_OK;
}

// ---- numDeriv::jacobian(ff$gr, theta), method "Richardson" (A.3) ------------------------------------------
static int richardson_jacobian(FF& ff, const std::vector<double>& x, std::vector<double>& J /* S x S col-major */,
                               double d = 1e-4) {
  const int n = (int)x.size(), r = 4;
  const double eps = 1e-4, v = 2.0;
  const double zero_tol = std::sqrt(std::numeric_limits<double>::epsilon() / 7e-7);
  std::vector<double> f0(n), h(n), gp(n), gm(n), xx(x);
  BGP_TRY(ff.gr(x.data(), f0.data()));
  for (int i = 0; i < n; ++i) h[i] = std::fabs(d * x[i]) + eps * (std::fabs(x[i]) < zero_tol ? 1.0 : 0.0);
  // a[k][row][col]
  std::vector<double> a((size_t)r * n * n, 0.0);
  for (int k = 0; k < r; ++k) {
    for (int i = 0; i < n; ++i) {
      xx = x;
      xx[i] = x[i] + h[i];
      BGP_TRY(ff.gr(xx.data(), gp.data()));
      xx[i] = x[i] - h[i];
      BGP_TRY(ff.gr(xx.data(), gm.data()));
      for (int row = 0; row < n; ++row) a[((size_t)k * n + row) * n + i] = (gp[row] - gm[row]) / (2.0 * h[i]);
    }
    for (int i = 0; i < n; ++i) h[i] /= v;
  }
  int rows = r;
  for (int mm = 1; mm < r; ++mm) {
    const double f4 = std::pow(4.0, mm);
    for (int k = 0; k + 1 < rows; ++k)
      for (int e = 0; e < n * n; ++e)
        a[(size_t)k * n * n + e] = (a[(size_t)(k + 1) * n * n + e] * f4 - a[(size_t)k * n * n + e]) / (f4 - 1.0);
    --rows;
  }
  J.assign((size_t)n * n, 0.0);
  for (int row = 0; row < n; ++row)
    for (int col = 0; col < n; ++col) J[(size_t)col * n + row] = a[(size_t)row * n + col];
  return BGP_OK;
}

// ---- Gauss-Hermite ("GHe": probabilists' nodes, weights that integrate g(z) dz) (A.4) ----------------------
static void gh_rule(int k, std::vector<double>& z, std::vector<double>& w) {
  // Golub-Welsch on the Jacobi matrix of He_k (off-diagonals sqrt(i)), Jacobi rotations, then Newton polish
  std::vector<double> A((size_t)k * k, 0.0);
  for (int i = 0; i + 1 < k; ++i) A[(size_t)i * k + i + 1] = A[(size_t)(i + 1) * k + i] = std::sqrt((double)(i + 1));
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < k; ++i)
      for (int j = i + 1; j < k; ++j) off += A[(size_t)i * k + j] * A[(size_t)i * k + j];
    if (off < 1e-30) break;
    for (int pI = 0; pI < k; ++pI)
      for (int q = pI + 1; q < k; ++q) {
        const double apq = A[(size_t)pI * k + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double th = (A[(size_t)q * k + q] - A[(size_t)pI * k + pI]) / (2.0 * apq);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < k; ++r) {
          const double arp = A[(size_t)r * k + pI], arq = A[(size_t)r * k + q];
          A[(size_t)r * k + pI] = c * arp - s * arq;
          A[(size_t)r * k + q] = s * arp + c * arq;
        }
        for (int r = 0; r < k; ++r) {
          const double apr = A[(size_t)pI * k + r], aqr = A[(size_t)q * k + r];
          A[(size_t)pI * k + r] = c * apr - s * aqr;
          A[(size_t)q * k + r] = s * apr + c * aqr;
        }
      }
  }
  z.resize(k);
  for (int i = 0; i < k; ++i) z[i] = A[(size_t)i * k + i];
  std::sort(z.begin(), z.end());
  auto he = [&](double x, double& hk, double& hkm1) {   // He_k(x), He_{k-1}(x)
    double p0 = 1.0, p1 = x;
    if (k == 0) {
      hk = 1.0;
      hkm1 = 0.0;
      return;
    }
    for (int j = 1; j < k; ++j) {
      const double p2 = x * p1 - j * p0;
      p0 = p1;
      p1 = p2;
    }
    hk = p1;
    hkm1 = p0;
  };
  for (int i = 0; i < k; ++i)
    for (int itn = 0; itn < 3; ++itn) {
      double hk, hkm1;
      he(z[i], hk, hkm1);
      z[i] -= hk / (k * hkm1);     // He_k' = k He_{k-1}
    }
  for (int i = 0; i < k; ++i) {   // exact symmetry
    const double a = 0.5 * (z[i] - z[k - 1 - i]);
    z[i] = a;
  }
  for (int i = 0; i < k / 2; ++i) z[k - 1 - i] = -z[i];
  if (k % 2 == 1) z[k / 2] = 0.0;
  double kfact = 1.0;
  for (int j = 2; j <= k; ++j) kfact *= j;
  w.resize(k);
  for (int i = 0; i < k; ++i) {
    double hk, hkm1;
    he(z[i], hk, hkm1);
    const double wn = kfact * std::sqrt(2.0 * M_PI) / ((double)k * k * hkm1 * hkm1);   // weight for exp(-z^2/2)
    w[i] = wn * std::exp(0.5 * z[i] * z[i]);
  }
}

static bool chol_lower(std::vector<double>& A, int n) {   // column-major in place, lower
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)k * n + j] * A[(size_t)k * n + j];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)j * n + i];
      for (int k = 0; k < j; ++k) s -= A[(size_t)k * n + i] * A[(size_t)k * n + j];
      A[(size_t)j * n + i] = s / d;
    }
    for (int i = 0; i < j; ++i) A[(size_t)j * n + i] = 0.0;
  }
  return true;
}

static bool invert_general(const std::vector<double>& Ain, int n, std::vector<double>& inv) {  // Gauss-Jordan, col-major
  std::vector<double> A(Ain);
  inv.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[(size_t)c * n + r]) > std::fabs(A[(size_t)c * n + piv])) piv = r;
    if (std::fabs(A[(size_t)c * n + piv]) == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        std::swap(A[(size_t)j * n + c], A[(size_t)j * n + piv]);
        std::swap(inv[(size_t)j * n + c], inv[(size_t)j * n + piv]);
      }
    const double d = A[(size_t)c * n + c];
    for (int j = 0; j < n; ++j) {
      A[(size_t)j * n + c] /= d;
      inv[(size_t)j * n + c] /= d;
    }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = A[(size_t)c * n + r];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) {
        A[(size_t)j * n + r] -= f * A[(size_t)j * n + c];
        inv[(size_t)j * n + r] -= f * inv[(size_t)j * n + c];
      }
    }
  }
  return true;
}

static double logsumexp(const std::vector<double>& v) {
  double mx = -std::numeric_limits<double>::infinity();
  for (double x : v) mx = std::max(mx, x);
  if (!std::isfinite(mx)) return mx;
  double s = 0.0;
  for (double x : v) s += std::exp(x - mx);
  return mx + std::log(s);
}

// mvQuad::rescale(grid, m = mode, C = forceSymmetric(solve(H)), dec.type = 2) for the coordinate order `ord`
static int rescaled_grid(const std::vector<double>& mode, const std::vector<double>& hess, int S, int k,
                         const std::vector<int>& ord, std::vector<double>& nodes /* K x S col-major */,
                         std::vector<double>& weights, double* L00) {
  std::vector<double> C;
  if (!invert_general(hess, S, C)) {
    set_error("theta Hessian is singular");
    return BGP_ERR_NOT_PD;
  }
  for (int i = 0; i < S; ++i)           // forceSymmetric keeps the upper triangle
    for (int j = i + 1; j < S; ++j) C[(size_t)i * S + j] = C[(size_t)j * S + i];
  std::vector<double> Cp((size_t)S * S);
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) Cp[(size_t)j * S + i] = C[(size_t)ord[j] * S + ord[i]];
  if (!chol_lower(Cp, S)) {
    set_error("inverse theta Hessian is not positive definite (quadrature cannot be centred)");
    return BGP_ERR_NOT_PD;
  }
  std::vector<double> z, w;
  gh_rule(k, z, w);
  int K = 1;
  for (int i = 0; i < S; ++i) K *= k;
  nodes.assign((size_t)K * S, 0.0);
  weights.assign(K, 0.0);
  double detL = 1.0;
  for (int i = 0; i < S; ++i) detL *= Cp[(size_t)i * S + i];
  std::vector<int> idx(S);
  std::vector<double> zz(S);
  for (int j = 0; j < K; ++j) {
    int r = j;
    double wprod = 1.0;
    for (int d = 0; d < S; ++d) {       // first coordinate fastest (expand.grid)
      idx[d] = r % k;
      r /= k;
      zz[d] = z[idx[d]];
      wprod *= w[idx[d]];
    }
    for (int a = 0; a < S; ++a) {
      double s = mode[ord[a]];
      for (int b = 0; b <= a; ++b) s += Cp[(size_t)b * S + a] * zz[b];
      nodes[(size_t)ord[a] * K + j] = s;
    }
    weights[j] = wprod * detL;
  }
  *L00 = Cp[0];
  return BGP_OK;
}

static int fit_core(bgp_model* m, int k, const double* theta0, const double* mode_in, const double* hess_in,
                    bgp_fit** out) {
  const int S = m->S, p = m->p;
  if (S < 1) {
    set_error("For model with no hyper-parameter, the method cannot be aghq");   // R/02_model_fit.R:253-255
    return BGP_ERR_ARG;
  }
  if (k < 1 || k > 64) {
    set_error("aghq_k must be in 1..64");
    return BGP_ERR_ARG;
  }
  bgp_fit* f = new bgp_fit();
  f->model = m;
  f->S = S;
  f->p = p;
  f->k = k;
  FF ff{m};
  auto fail = [&](int code) {
    fit_release_device(f);
    delete f;
    return code;
  };
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_opt0 = now();
  cudaEventRecord(m->ev[0], m->stream);     // device time of the whole call -> bgp_model_last_timing(total_ms)
  f->mode.assign(S, 0.0);
  if (mode_in) {
    f->mode.assign(mode_in, mode_in + S);
  } else {
    for (int i = 0; i < S; ++i) f->mode[i] = theta0 ? theta0[i] : 0.0;
    double Fmin;
    int st = vmmin(ff, f->mode, &Fmin, &f->convergence);
    if (st != BGP_OK) return fail(st);
  }
  if (hess_in) {
    f->hessian.assign(hess_in, hess_in + (size_t)S * S);
  } else {
    int st = richardson_jacobian(ff, f->mode, f->hessian);
    if (st != BGP_OK) return fail(st);
    auto is_pd = [&](const std::vector<double>& H) {
      std::vector<double> C;
      if (!invert_general(H, S, C)) return false;
      for (int i = 0; i < S; ++i)
        for (int j = i + 1; j < S; ++j) C[(size_t)i * S + j] = C[(size_t)j * S + i];
      return chol_lower(C, S);
    };
    const bool dbg = getenv("BGP_FIT_DEBUG") != nullptr;
    auto dump = [&](const char* tag) {
      if (!dbg) return;
      fprintf(stderr, "[fit] %s mode:", tag);
      for (int i = 0; i < S; ++i) fprintf(stderr, " %.8f", f->mode[i]);
      fprintf(stderr, " hessian:");
      for (int i = 0; i < S * S; ++i) fprintf(stderr, " %.6g", f->hessian[i]);
      fprintf(stderr, "\n");
    };
    dump("richardson d=1e-4");
    // numDeriv's default step (d = 1e-4) differentiates a gradient that carries ~cond(H) * eps * |L| of rounding
    // noise; at large n the result can come out indefinite, where aghq stops with a chol() error — and so does
    // this library (rescaled_grid below) unless the caller opted into the retry with 10x / 100x larger steps
    // (bgp_model_set_hessian_retry); the number of retries is reported by bgp_fit_get_diagnostics.
    if (m->hessian_retry)
      for (double d = 1e-3; !is_pd(f->hessian) && d <= 1.001e-2; d *= 10.0) {
        st = richardson_jacobian(ff, f->mode, f->hessian, d);
        if (st != BGP_OK) return fail(st);
        ++f->hessian_fallback;
        dump("richardson retry");
      }
  }
  f->opt_ms = now() - t_opt0;
  const double t_grid0 = now();
  std::vector<int> ord(S);
  for (int i = 0; i < S; ++i) ord[i] = i;
  double L00 = 1.0;
  int st = rescaled_grid(f->mode, f->hessian, S, k, ord, f->nodes, f->weights, &L00);
  if (st != BGP_OK) return fail(st);
  const int K = (int)f->weights.size();
  f->K = K;
  f->logpost.assign(K, 0.0);
  // Node shards: the K nodes of a grid are dealt to the ranks of the node group in contiguous runs of
  // the expand.grid order (neighbours along the first coordinate share a rank: good warm starts).  Each rank keeps
  // the modes / Hessians of its nodes on its device; only the K values are exchanged.
  const int nw = m->node_world, nr = m->node_rank;
  f->owner.resize(K);
  f->slot.assign(K, -1);
  std::vector<unsigned char> mine(K, 0);
  // Contiguous, count-balanced runs of the expand.grid order.  Measured on C3 (scripts/partition_probe.py,
  // profiles/r02_partition_probe.txt): from the optimiser's end state the FIRST node of any run costs two Newton
  // iterations (4.0 ms) wherever it lies — the first-order predictor is not accurate enough for one — and every
  // further node ~2.2 ms, so runs of equal length are the best split; weighting far runs lighter made it worse.
  for (int j = 0; j < K; ++j) f->owner[j] = piece_owner(j, K, nw);
  for (int j = 0; j < K; ++j)
    if (f->owner[j] == nr) {
      mine[j] = 1;
      f->slot[j] = f->n_local++;
    }
  f->model_alive = m->alive;
  if (f->n_local > 0) {
    f->modes_dev = (double*)pool_take(m, false, (size_t)f->n_local * m->lda * sizeof(double), &f->modes_dev_bytes);
    f->Hs_dev = (double*)pool_take(m, false, (size_t)f->n_local * p * m->ldh * sizeof(double), &f->Hs_dev_bytes);
    if (!f->modes_dev || !f->Hs_dev) return fail(BGP_ERR_CUDA);
  }
  // host mirror of modesandhessians (external order), page-locked: each node's mode / Hessian follows its
  // evaluation out over PCIe while the next node is being evaluated.  Skipped when it would not fit comfortably
  // in host memory (then bgp_fit_get_modes rotates and copies on request).
  const size_t mirror_need = ((size_t)K * p + (size_t)K * p * p) * sizeof(double);
  if (mirror_need <= ((size_t)2 << 30)) {
    f->mirror = (double*)pool_take(m, true, mirror_need, &f->mirror_bytes);
    if (!f->mirror) return fail(BGP_ERR_CUDA);
    f->mirrored.assign(K, 0);
  }
  double* vals_dev = nullptr;
  if (nw > 1 && cudaMalloc(&vals_dev, (size_t)K * sizeof(double)) != cudaSuccess) {
    set_error("cudaMalloc failed");
    return fail(BGP_ERR_CUDA);
  }
  struct Free {
    double* p;
    ~Free() { if (p) cudaFree(p); }
  } free_vals{vals_dev};
  std::vector<double> th((size_t)S * K), vals(K);
  // one grid: this rank's nodes through the batch path (centre-out order, nothing but scalars crosses PCIe), then
  // the values of all ranks.  aghq stops when a node's log posterior is not finite; so does this (every rank sees
  // the NaN through the all-reduce and returns the same error).
  auto eval_grid = [&](const std::vector<double>& nodes, std::vector<double>& lp, bool keep) -> int {
    for (int j = 0; j < K; ++j)
      for (int a = 0; a < S; ++a) th[(size_t)j * S + a] = nodes[(size_t)a * K + j];
    std::fill(vals.begin(), vals.end(), 0.0);
    BatchSink sink;
    if (keep) {
      sink.modes_dev = f->modes_dev;
      sink.Hs_dev = f->Hs_dev;
      sink.dev_slot = f->slot.data();
      if (f->mirror) {
        sink.modes_host = f->mirror;
        sink.Hs_host = f->mirror + (size_t)K * p;
        sink.host_pinned = true;
        for (int j = 0; j < K; ++j) f->mirrored[j] = (nw > 1 ? mine[j] : 1);
      }
    }
    int iters = 0, bad = -1;
    int rc = laplace_batch(m, K, th.data(), nw > 1 ? mine.data() : nullptr, vals.data(), sink, &iters, &bad);
    if (rc == BGP_ERR_CUDA || rc == BGP_ERR_NCCL) return rc;
    std::string local_msg = rc != BGP_OK ? g_last_error : std::string();
    f->grid_newton_iters += iters;
    ff.n_fn += K;
    if (nw > 1) {
      BGP_CUDA(cudaMemcpyAsync(vals_dev, vals.data(), (size_t)K * sizeof(double), cudaMemcpyHostToDevice, m->stream));
      BGP_TRY(node_allreduce_sum(m, vals_dev, (size_t)K));
      BGP_CUDA(cudaMemcpyAsync(vals.data(), vals_dev, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    }
    BGP_CUDA(cudaStreamSynchronize(m->stream));
    for (int j = 0; j < K; ++j) {
      if (!std::isfinite(vals[j])) {
        std::string at;
        for (int a = 0; a < S; ++a) at += (a ? ", " : "") + std::to_string(th[(size_t)j * S + a]);
        set_error("log posterior is not finite at quadrature node %d (theta = %s)%s%s", j, at.c_str(),
                  local_msg.empty() ? "" : ": ", local_msg.c_str());
        return rc != BGP_OK ? rc : BGP_ERR_NONFINITE;
      }
      lp[j] = -vals[j];
    }
    return BGP_OK;
  };
  st = eval_grid(f->nodes, f->logpost, true);
  if (st != BGP_OK) return fail(st);
  {
    std::vector<double> t(K);
    for (int j = 0; j < K; ++j) t[j] = f->logpost[j] + std::log(f->weights[j]);
    f->lognormconst = logsumexp(t);
  }
  f->logpost_norm.resize(K);
  for (int j = 0; j < K; ++j) f->logpost_norm[j] = f->logpost[j] - f->lognormconst;
  // marginals, method "reuse" (A.5)
  std::vector<double> z1, w1;
  gh_rule(k, z1, w1);
  f->marg_theta.resize(S);
  f->marg_lmp.resize(S);
  f->marg_w.resize(S);
  for (int j = 0; j < S; ++j) {
    std::vector<double> nodes_j, weights_j, lpn_j(K);
    double l00 = L00;
    if (j == 0) {
      weights_j = f->weights;
      lpn_j = f->logpost_norm;
    } else {
      std::vector<int> o2;
      o2.push_back(j);
      for (int i = 0; i < S; ++i)
        if (i != j) o2.push_back(i);
      st = rescaled_grid(f->mode, f->hessian, S, k, o2, nodes_j, weights_j, &l00);
      if (st != BGP_OK) return fail(st);
      std::vector<double> lp(K);
      st = eval_grid(nodes_j, lp, false);
      if (st != BGP_OK) return fail(st);
      std::vector<double> t(K);
      for (int q = 0; q < K; ++q) t[q] = lp[q] + std::log(weights_j[q]);
      const double lnc = logsumexp(t);
      for (int q = 0; q < K; ++q) lpn_j[q] = lp[q] - lnc;
    }
    f->marg_theta[j].resize(k);
    f->marg_lmp[j].resize(k);
    f->marg_w[j].resize(k);
    for (int q = 0; q < k; ++q) {
      std::vector<double> t;
      for (int e = q; e < K; e += k) t.push_back(lpn_j[e] + std::log(weights_j[e]));   // first coordinate == q
      f->marg_theta[j][q] = f->mode[j] + l00 * z1[q];
      f->marg_w[j][q] = w1[q] * l00;
      f->marg_lmp[j][q] = logsumexp(t) - std::log(f->marg_w[j][q]);
    }
  }
  cudaEventRecord(m->ev[1], m->stream);
  if (cudaStreamSynchronize(m->stream) != cudaSuccess) {
    set_error("CUDA error at the end of the fit");
    return fail(BGP_ERR_CUDA);
  }
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]);
    m->t_total = ms;
  }
  f->grid_ms = now() - t_grid0;
  f->fn_count = ff.n_fn;
  f->gr_count = ff.n_gr;
  *out = f;
  return BGP_OK;
}

// modesandhessians in the caller's layout (external order).  Nodes evaluated by this rank are already in the pinned
// mirror; nodes held by other ranks of the node group are rotated on their owner's device, gathered over the group
// (zeros from everyone else) and copied in, a few nodes at a time.  Without a mirror everything takes that route.
static int complete_mirror(const bgp_fit* f, double* modes, double* Hs) {
  bgp_model* m = f->model;
  const int p = f->p, K = f->K;
  const size_t pp = (size_t)p * p;
  double* dst_modes = f->mirror ? f->mirror : modes;
  double* dst_Hs = f->mirror ? f->mirror + (size_t)K * p : Hs;
  std::vector<int> todo;
  for (int j = 0; j < K; ++j)
    if (!f->mirror || !f->mirrored[j]) todo.push_back(j);
  // every rank of a node group walks the same list (its own nodes are "mirrored", the others' are not, so the
  // lists differ): the collective is therefore over ALL nodes that are missing on ANY rank = all K when sharded
  if (m->node_world > 1) {
    todo.clear();
    for (int j = 0; j < K; ++j) todo.push_back(j);
  }
  if (todo.empty()) return BGP_OK;
  const bool want_H = f->mirror || Hs;
  const size_t per = (want_H ? pp : 0) + (size_t)p;
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(todo.size(), ((size_t)256 << 20) / (per * sizeof(double))));
  size_t stage_bytes = 0;
  double* stage = (double*)pool_take(m, false, (size_t)chunk * per * sizeof(double), &stage_bytes);
  if (!stage) return BGP_ERR_CUDA;
  std::vector<double> host;
  int rc = [&]() -> int {
    for (size_t t0 = 0; t0 < todo.size(); t0 += chunk) {
      const int nj = (int)std::min<size_t>(chunk, todo.size() - t0);
      if (m->node_world > 1) BGP_CUDA(cudaMemsetAsync(stage, 0, (size_t)nj * per * sizeof(double), m->stream));
      for (int q = 0; q < nj; ++q) {
        const int j = todo[t0 + q];
        if (f->slot[j] < 0) continue;
        double* dst = stage + (size_t)q * per;
        BGP_TRY(rot_vec_dev(m, f->modes_dev + (size_t)f->slot[j] * m->lda, dst));
        if (want_H) BGP_TRY(rot_H_dev(m, f->Hs_dev + (size_t)f->slot[j] * p * m->ldh, dst + p, p));
      }
      BGP_TRY(node_allreduce_sum(m, stage, (size_t)nj * per));
      if (f->mirror) {
        for (int q = 0; q < nj; ++q) {
          const int j = todo[t0 + q];
          if (f->mirrored[j]) continue;
          BGP_CUDA(cudaMemcpyAsync(dst_modes + (size_t)j * p, stage + (size_t)q * per, (size_t)p * sizeof(double),
                                   cudaMemcpyDeviceToHost, m->stream));
          BGP_CUDA(cudaMemcpyAsync(dst_Hs + (size_t)j * pp, stage + (size_t)q * per + p, pp * sizeof(double),
                                   cudaMemcpyDeviceToHost, m->stream));
        }
        BGP_CUDA(cudaStreamSynchronize(m->stream));
      } else {
        host.resize((size_t)nj * per);
        BGP_CUDA(cudaMemcpyAsync(host.data(), stage, (size_t)nj * per * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        BGP_CUDA(cudaStreamSynchronize(m->stream));
        for (int q = 0; q < nj; ++q) {
          const int j = todo[t0 + q];
          const double* src = host.data() + (size_t)q * per;
          if (dst_modes) std::copy(src, src + p, dst_modes + (size_t)j * p);
          if (dst_Hs) std::copy(src + p, src + p + pp, dst_Hs + (size_t)j * pp);
        }
      }
    }
    return BGP_OK;
  }();
  pool_give(m, false, stage, stage_bytes);
  if (rc == BGP_OK && f->mirror)
    for (int j = 0; j < K; ++j) const_cast<bgp_fit*>(f)->mirrored[j] = 1;
  return rc;
}

}  // namespace bgp

using namespace bgp;

extern "C" {

int bgp_aghq_fit(bgp_model* m, int k, const double* theta0, bgp_fit** out) {
  if (!m || !m->finalized || !out) {
    set_error("bgp_aghq_fit: model not finalized / NULL output");
    return BGP_ERR_STATE;
  }
  BGP_CUDA(cudaSetDevice(m->device));
  return fit_core(m, k, theta0, nullptr, nullptr, out);
}

int bgp_aghq_fit_at(bgp_model* m, int k, const double* mode, const double* hessian, bgp_fit** out) {
  if (!m || !m->finalized || !out || !mode || !hessian) {
    set_error("bgp_aghq_fit_at: bad arguments");
    return BGP_ERR_STATE;
  }
  BGP_CUDA(cudaSetDevice(m->device));
  return fit_core(m, k, nullptr, mode, hessian, out);
}

void bgp_fit_destroy(bgp_fit* f) {
  if (!f) return;
  fit_release_device(f);
  delete f;
}

int bgp_fit_dims(const bgp_fit* f, int* S, int* K, int* p, int* k) {
  if (!f) return BGP_ERR_ARG;
  if (S) *S = f->S;
  if (K) *K = f->K;
  if (p) *p = f->p;
  if (k) *k = f->k;
  return BGP_OK;
}

int bgp_fit_get_opt(const bgp_fit* f, double* mode, double* hessian, int* convergence, int* fn_count, int* gr_count) {
  if (!f) return BGP_ERR_ARG;
  if (mode) std::copy(f->mode.begin(), f->mode.end(), mode);
  if (hessian) std::copy(f->hessian.begin(), f->hessian.end(), hessian);
  if (convergence) *convergence = f->convergence;
  if (fn_count) *fn_count = f->fn_count;
  if (gr_count) *gr_count = f->gr_count;
  return BGP_OK;
}

int bgp_fit_get_grid(const bgp_fit* f, double* nodes, double* weights, double* logpost, double* logpost_normalized,
                     double* lognormconst) {
  if (!f) return BGP_ERR_ARG;
  if (nodes) std::copy(f->nodes.begin(), f->nodes.end(), nodes);
  if (weights) std::copy(f->weights.begin(), f->weights.end(), weights);
  if (logpost) std::copy(f->logpost.begin(), f->logpost.end(), logpost);
  if (logpost_normalized) std::copy(f->logpost_norm.begin(), f->logpost_norm.end(), logpost_normalized);
  if (lognormconst) *lognormconst = f->lognormconst;
  return BGP_OK;
}

int bgp_fit_get_modes(const bgp_fit* f, double* modes, double* Hs) {
  if (!f || !f->model) return BGP_ERR_ARG;
  if (!modes && !Hs) return BGP_OK;
  BGP_CUDA(cudaSetDevice(f->model->device));
  BGP_TRY(complete_mirror(f, modes, Hs));
  if (f->mirror) {
    const size_t nm = (size_t)f->K * f->p;
    if (modes) memcpy(modes, f->mirror, nm * sizeof(double));
    if (Hs) memcpy(Hs, f->mirror + nm, nm * f->p * sizeof(double));
  }
  return BGP_OK;
}

int bgp_fit_host_arrays(const bgp_fit* f, const double** modes, const double** Hs) {
  if (!f || !f->model) return BGP_ERR_ARG;
  if (!f->mirror) {
    set_error("this fit keeps no host mirror of its modes / Hessians (more than 2 GiB): use bgp_fit_get_modes");
    return BGP_ERR_STATE;
  }
  BGP_CUDA(cudaSetDevice(f->model->device));
  BGP_TRY(complete_mirror(f, nullptr, nullptr));
  if (modes) *modes = f->mirror;
  if (Hs) *Hs = f->mirror + (size_t)f->K * f->p;
  return BGP_OK;
}

int bgp_fit_get_diagnostics(const bgp_fit* f, int* hessian_fallback, int64_t* grid_newton_iters, double* opt_ms,
                            double* grid_ms) {
  if (!f) return BGP_ERR_ARG;
  if (hessian_fallback) *hessian_fallback = f->hessian_fallback;
  if (grid_newton_iters) *grid_newton_iters = f->grid_newton_iters;
  if (opt_ms) *opt_ms = f->opt_ms;
  if (grid_ms) *grid_ms = f->grid_ms;
  return BGP_OK;
}

int bgp_fit_node_owner(const bgp_fit* f, int32_t* owner) {
  if (!f || !owner) return BGP_ERR_ARG;
  for (int j = 0; j < f->K; ++j) owner[j] = f->owner[j];
  return BGP_OK;
}

int bgp_fit_get_marginal(const bgp_fit* f, int j, double* theta, double* logmargpost, double* w) {
  if (!f || j < 0 || j >= f->S) {
    set_error("bgp_fit_get_marginal: index out of range");
    return BGP_ERR_ARG;
  }
  if (theta) std::copy(f->marg_theta[j].begin(), f->marg_theta[j].end(), theta);
  if (logmargpost) std::copy(f->marg_lmp[j].begin(), f->marg_lmp[j].end(), logmargpost);
  if (w) std::copy(f->marg_w[j].begin(), f->marg_w[j].end(), w);
  return BGP_OK;
}

}  // extern "C"
