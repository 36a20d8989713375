"""``ff`` — the TMB-style Laplace objective, backed by libbgp on a B200.

Host-side mirror of what ``TMB::MakeADFun(data = tmbdat, parameters = tmbparams,
random = "W", DLL = "BayesGP")`` returns at ``/root/reference/R/02_model_fit.R:276-283``:
the same names (``par``, ``fn``, ``gr``, ``he``, ``env.last_par``, ``env.spHess``), the
same argument meaning and the same failure behaviour (NaN value on inner-Newton failure).
All arithmetic runs in the CUDA library; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import BgpError, check, dptr, fmat, fvec

FAMILY_CODES = {"Gaussian": 0, "Poisson": 1, "Binomial": 2, "none": -2}   # R/02_model_fit.R:9-28


class TMBData:
    """The ``tmbdat`` list of R/02_model_fit.R:152-183 (dense column-major blocks)."""

    def __init__(self, y, family, B, P, logPdet, u, alpha, X, betaprec, betamean, Xf, beta_fixed_prec,
                 beta_fixed_mean, size=None):
        self.y = fvec(y)
        self.family = FAMILY_CODES[family] if isinstance(family, str) else int(family)
        self.B = [fmat(b) for b in B]
        self.P = [np.asarray(p, dtype=np.float64) for p in P]       # 1-D => diagonal
        self.logPdet = [float(v) for v in logPdet]
        self.u = [float(v) for v in u]
        self.alpha = [float(v) for v in alpha]
        self.X = [fmat(np.asarray(x, dtype=np.float64).reshape(len(self.y), -1)) for x in X]
        self.betaprec = [float(v) for v in betaprec]
        self.betamean = [float(v) for v in betamean]
        self.Xf = [fmat(np.asarray(x, dtype=np.float64).reshape(len(self.y), -1)) for x in Xf]
        self.beta_fixed_prec = [float(v) for v in beta_fixed_prec]
        self.beta_fixed_mean = [float(v) for v in beta_fixed_mean]
        self.size = None if size is None else fvec(size)


class _Env:
    """``ff$env``: last.par (mode of the most recent evaluation) and spHess."""

    def __init__(self, owner):
        self._o = owner
        self.last_par = None
        self.last_theta = None
        self.last_H = None

    def spHess(self, par=None, random=True):
        return self.last_H


class LaplaceObjective:
    """ff <- MakeADFun(..., random = "W").  ``fn``/``gr`` take theta (length S)."""

    def __init__(self, data: Optional[TMBData] = None, device: int = 0, *, y=None, family=None, size=None):
        """Either pass a complete ``TMBData`` (dense blocks, as R builds them) or ``y``/``family``
        and then add blocks with ``add_random`` / ``add_boundary`` / ``add_fixed`` / ``add_iwp``
        (same order as the W layout of src/BayesGP.cpp:76-127) before ``finalize()``."""
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.device = device
        self._finalized = False
        self.env = _Env(self)
        self.n_fn = 0
        self.n_gr = 0
        self.newton_iters = 0
        self.node_rank, self.node_world = 0, 1
        if data is not None:
            y, family, size = data.y, data.family, data.size
        fam = FAMILY_CODES[family] if isinstance(family, str) else int(family)
        self.family = fam
        yv = fvec(y)
        sv = None if size is None else fvec(size)
        check(self._lib.bgp_model_new(len(yv), fam, dptr(yv), dptr(sv), device, C.byref(self._h)))
        if data is not None:
            d = data
            try:
                for j in range(len(d.B)):
                    self.add_random(d.B[j], d.P[j], d.logPdet[j], d.u[j], d.alpha[j])
                for j in range(len(d.X)):
                    self.add_boundary(d.X[j], d.betaprec[j], d.betamean[j])
                for j in range(len(d.Xf)):
                    self.add_fixed(d.Xf[j], d.beta_fixed_prec[j], d.beta_fixed_mean[j])
                if fam == 0:
                    self.set_noise_prior(d.u[-1], d.alpha[-1])
            except Exception:
                self.close()
                raise

    # -- builder (tmbdat lists, R/02_model_fit.R:50-71,123-150) -------------------------------------
    def add_random(self, B, P, logPdet, u=1.0, alpha=0.5):
        B = fmat(B)
        P = np.asarray(P, dtype=np.float64)
        diag = P.ndim == 1
        Pm = fvec(P) if diag else fmat(P)
        check(self._lib.bgp_model_add_random(self._h, B.shape[1], dptr(B), dptr(Pm), int(diag), float(logPdet),
                                             float(u), float(alpha)))

    def add_boundary(self, X, prec=0.01, mean=0.0):
        X = fmat(X)
        ncol = X.shape[1] if X.ndim == 2 else 0
        check(self._lib.bgp_model_add_boundary(self._h, ncol, dptr(X) if ncol else None, float(prec), float(mean)))

    def add_fixed(self, Xf, prec=0.01, mean=0.0):
        Xf = fmat(np.asarray(Xf, dtype=np.float64).reshape(-1, 1) if np.ndim(Xf) == 1 else Xf)
        check(self._lib.bgp_model_add_fixed(self._h, Xf.shape[1], dptr(Xf), float(prec), float(mean)))

    def add_iwp(self, x, initial_location, knots, order, u=1.0, alpha=0.5, boundary_prec=0.01, boundary_mean=0.0):
        """Device-side construction of an IWP term from the covariate (no n x k host matrix)."""
        x = fvec(x)
        knots = fvec(knots)
        check(self._lib.bgp_model_add_iwp(self._h, dptr(x), float(initial_location), dptr(knots), len(knots),
                                          int(order), float(u), float(alpha), float(boundary_prec),
                                          float(boundary_mean)))

    def add_sgp(self, x, initial_location, a, k, m, region, P, logPdet, u=1.0, alpha=0.5, boundary_prec=0.01,
                boundary_mean=0.0):
        """Device-side construction of an sGP design (B and X) from the covariate; P comes from the host."""
        x = fvec(x)
        reg = fvec(np.asarray(region, dtype=np.float64)[:2])
        Pm = fmat(P)
        check(self._lib.bgp_model_add_sgp(self._h, dptr(x), float(initial_location), float(a), int(k), int(m), dptr(reg),
                                          dptr(Pm), float(logPdet), float(u), float(alpha), float(boundary_prec),
                                          float(boundary_mean)))

    def add_sgp_auto(self, x, initial_location, a, k, m, region, accuracy=0.01, u=1.0, alpha=0.5, boundary_prec=0.01,
                     boundary_mean=0.0):
        """sGP term entirely on the device: design from the covariate, precision by Compute_Q_sB (bgp_sgp_precision)."""
        x = fvec(x)
        reg = fvec(np.asarray(region, dtype=np.float64)[:2])
        check(self._lib.bgp_model_add_sgp_auto(self._h, dptr(x), float(initial_location), float(a), int(k), int(m),
                                               dptr(reg), float(accuracy), float(u), float(alpha), float(boundary_prec),
                                               float(boundary_mean)))

    def set_noise_prior(self, u=1.0, alpha=0.5):
        check(self._lib.bgp_model_set_noise_prior(self._h, float(u), float(alpha)))

    # -- lifecycle ------------------------------------------------------------------------------
    def set_shard(self, rank: int, world: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        check(self._lib.bgp_model_set_shard(self._h, rank, world, buf))

    def set_node_group(self, rank: int, world: int, unique_id: bytes):
        """Join a node group: replicas that split quadrature nodes, sample blocks and prediction rows."""
        buf = C.create_string_buffer(unique_id, 128)
        check(self._lib.bgp_model_set_node_group(self._h, rank, world, buf))
        self.node_rank, self.node_world = rank, world

    def set_hessian_retry(self, allow=True):
        check(self._lib.bgp_model_set_hessian_retry(self._h, int(allow)))

    def set_ospline(self, on=True):
        """Select the O-spline moment path (eligible models: one IWP term), the dense DMMA path (on=False), or the
        moment path with the gradient's leverages taken from the dense design (on=2; A/B)."""
        check(self._lib.bgp_model_set_ospline(self._h, int(on)))

    def set_lanes(self, lanes):
        """Concurrent evaluation contexts for batches on the moment path (bgp_model_set_lanes)."""
        check(self._lib.bgp_model_set_lanes(self._h, int(lanes)))

    def lanes(self):
        n = C.c_int()
        check(self._lib.bgp_model_get_lanes(self._h, C.byref(n)))
        return n.value

    def ospline_bytes(self):
        """Algorithmic bytes one likelihood pass of the moment path moves."""
        b = C.c_double()
        check(self._lib.bgp_model_ospline_bytes(self._h, C.byref(b)))
        return b.value

    def ospline(self):
        """(eligible, on) of the O-spline moment path."""
        e, o = C.c_int(), C.c_int()
        check(self._lib.bgp_model_get_ospline(self._h, C.byref(e), C.byref(o)))
        return bool(e.value), bool(o.value)

    def finalize(self):
        check(self._lib.bgp_model_finalize(self._h))
        n, p, S = C.c_int64(), C.c_int(), C.c_int()
        check(self._lib.bgp_model_dims(self._h, C.byref(n), C.byref(p), C.byref(S)))
        self.n, self.p, self.S = n.value, p.value, S.value
        self.par = np.zeros(self.S)                      # tmbparams theta = 0 (R/02_model_fit.R:249-252)
        self.env.last_par = np.zeros(self.p)
        self._finalized = True
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.bgp_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the TMB objective itself (objective_function::operator()) ---------------------------------
    def objective(self, W, theta, want_grad=True, want_hess=False):
        W = fvec(W)
        theta = fvec(np.atleast_1d(theta))
        f = C.c_double()
        g = np.empty(self.p) if want_grad else None
        H = np.empty((self.p, self.p), order="F") if want_hess else None
        check(self._lib.bgp_objective(self._h, dptr(W), dptr(theta), C.byref(f), dptr(g), dptr(H)))
        return f.value, g, H

    # -- ff$fn / ff$gr ------------------------------------------------------------------------------
    def _eval(self, theta, want_grad=False, want_mode=True, want_hess=False):
        if not self._finalized:
            raise BgpError(7, "model not finalized")          # BGP_ERR_STATE, as the C entry points answer
        theta = fvec(np.atleast_1d(theta))
        val = C.c_double()
        iters = C.c_int()
        g = np.empty(self.S) if want_grad else None
        w = np.empty(self.p) if want_mode else None
        H = np.empty((self.p, self.p), order="F") if want_hess else None
        code = self._lib.bgp_laplace_eval(self._h, dptr(theta), C.byref(val), dptr(g), dptr(w), dptr(H),
                                          C.byref(iters))
        self.newton_iters += iters.value
        if code in (3, 4, 5):        # NOT_PD / NONFINITE / NO_CONVERGENCE: TMB returns NaN (+ warning)
            self.last_warning = "%d: %s" % (code, self._lib.bgp_last_error().decode("utf-8", "replace"))
            return float("nan"), (np.full(self.S, np.nan) if want_grad else None), None, None
        check(code)
        if want_mode:
            self.env.last_par = w
            self.env.last_theta = theta.copy()
        if want_hess:
            self.env.last_H = H
        return val.value, g, w, H

    def fn(self, theta, want_hess=False):
        self.n_fn += 1
        return self._eval(theta, want_hess=want_hess)[0]

    def gr(self, theta):
        self.n_gr += 1
        return self._eval(theta, want_grad=True)[1]

    def fn_batch(self, thetas, want_modes=True, want_hess=False):
        """K Laplace evaluations through one C call; thetas is K x S."""
        thetas = np.ascontiguousarray(np.asarray(thetas, dtype=np.float64).reshape(-1, self.S))
        K = thetas.shape[0]
        vals = np.empty(K)
        modes = np.empty((K, self.p)) if want_modes else None
        Hs = np.empty((K, self.p, self.p)) if want_hess else None
        iters = C.c_int()
        code = self._lib.bgp_laplace_eval_batch(self._h, K, dptr(thetas), dptr(vals), dptr(modes), dptr(Hs),
                                                C.byref(iters))
        self.newton_iters += iters.value
        self.n_fn += K
        if code not in (0, 3, 4, 5):
            check(code)
        return vals, modes, Hs, iters.value

    def set_start(self, W=None):
        check(self._lib.bgp_model_set_start(self._h, dptr(fvec(W)) if W is not None else None))

    def get_tangent(self):
        """(theta, T) of the most recent evaluation: T = d w_hat / d theta, p x S."""
        th = np.empty(self.S)
        T = np.empty((self.p, self.S), order="F")
        check(self._lib.bgp_model_get_tangent(self._h, dptr(th), dptr(T)))
        return th, T

    def set_start_at(self, theta, W, T=None):
        """Restore the state an optimiser leaves at its last evaluation (mode + tangent there)."""
        th = fvec(np.atleast_1d(theta))
        Tm = None if T is None else fmat(np.asarray(T, dtype=np.float64).reshape(self.p, self.S))
        check(self._lib.bgp_model_set_start_at(self._h, dptr(th), dptr(fvec(W)), dptr(Tm)))

    def set_newton(self, grad_tol=1e-8, step_tol=1e-8, maxit=100):
        check(self._lib.bgp_model_set_newton(self._h, grad_tol, step_tol, maxit))

    def last_timing(self):
        t = [C.c_double() for _ in range(4)]
        k = [C.c_int64() for _ in range(3)]
        check(self._lib.bgp_model_last_timing(self._h, *[C.byref(v) for v in t], *[C.byref(v) for v in k]))
        return {"total_ms": t[0].value, "lik_ms": t[1].value, "hess_ms": t[2].value, "chol_ms": t[3].value,
                "lik_launches": k[0].value, "hess_launches": k[1].value, "chol_launches": k[2].value}


    def counters(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self._lib.bgp_model_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"laplace_evals": a.value, "newton_iters": b.value, "factor_reuses": c.value}

    def gradient_timing(self):
        t, n, f, d = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
        check(self._lib.bgp_model_gradient_timing(self._h, C.byref(t), C.byref(n), C.byref(f), C.byref(d)))
        return {"leverage_ms": t.value, "leverage_launches": n.value, "leverage_flops": f.value, "dense_flops": d.value}

    def set_factor_reuse(self, allow=True, eta_tol=1e-7, rel_tol=1e-10):
        check(self._lib.bgp_model_set_factor_reuse(self._h, int(allow), float(eta_tol), float(rel_tol)))

    def lik_bytes(self):
        d, u = C.c_double(), C.c_double()
        check(self._lib.bgp_model_lik_bytes(self._h, C.byref(d), C.byref(u)))
        return {"dense": d.value, "structural": u.value}

    def hessian_flops(self):
        d, u = C.c_double(), C.c_double()
        check(self._lib.bgp_model_hessian_flops(self._h, C.byref(d), C.byref(u)))
        return {"dense": d.value, "structural": u.value}


def make_objective(data: TMBData, device: int = 0) -> LaplaceObjective:
    return LaplaceObjective(data, device).finalize()
