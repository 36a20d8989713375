"""model_fit() / predict() — host-side mirror of BayesGP's user-facing calls on top of libbgp.

Same names, argument meaning and error behaviour as ``/root/reference/R/02_model_fit.R:336-701``
(``model_fit``), ``/root/reference/R/03_post_fit.R:53-125`` (``predict.FitResult``), ``:159-165``
(``sample_fixed_effect``), ``:200-296`` (``compute_post_fun_IWP`` / ``compute_post_fun_sGP`` /
``extract_mean_interval_given_samps``), with the formula DSL replaced by a list of ``Term``s.
Everything numerical happens behind the C ABI; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from ._lib import check, dptr, fmat, fvec
from .objective import FAMILY_CODES, LaplaceObjective
from .terms import Term, iid_design, prepare_term


class AGHQ:
    """The ``c("marginallaplace", "aghq")`` object: fields BayesGP reads."""

    def __init__(self, ff: LaplaceObjective, handle):
        self._lib = _lib.load()
        self.ff = ff
        self._h = handle
        S, K, p, k = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self._lib.bgp_fit_dims(self._h, C.byref(S), C.byref(K), C.byref(p), C.byref(k)))
        self.S, self.K, self.p, self.k = S.value, K.value, p.value, k.value
        mode, hess = np.empty(self.S), np.empty((self.S, self.S), order="F")
        conv, nfn, ngr = C.c_int(), C.c_int(), C.c_int()
        check(self._lib.bgp_fit_get_opt(self._h, dptr(mode), dptr(hess), C.byref(conv), C.byref(nfn), C.byref(ngr)))
        self.optresults = {"mode": mode, "hessian": np.array(hess), "convergence": conv.value, "ff": ff,
                           "fn_count": nfn.value, "gr_count": ngr.value}
        nodes = np.empty((self.K, self.S), order="F")
        w, lp, lpn = np.empty(self.K), np.empty(self.K), np.empty(self.K)
        lnc = C.c_double()
        check(self._lib.bgp_fit_get_grid(self._h, dptr(nodes), dptr(w), dptr(lp), dptr(lpn), C.byref(lnc)))
        self.normalized_posterior = {"nodesandweights": {"theta": np.array(nodes), "weights": w, "logpost": lp,
                                                         "logpost_normalized": lpn},
                                     "lognormconst": lnc.value, "grid": {"level": self.k}}
        self.marginals = []
        for j in range(self.S):
            th, lm, ww = np.empty(self.k), np.empty(self.k), np.empty(self.k)
            check(self._lib.bgp_fit_get_marginal(self._h, j, dptr(th), dptr(lm), dptr(ww)))
            self.marginals.append({"theta": th, "logmargpost": lm, "w": ww})
        self._modes = None
        self._Hs = None
        fb, it, t_opt, t_grid = C.c_int(), C.c_int64(), C.c_double(), C.c_double()
        check(self._lib.bgp_fit_get_diagnostics(self._h, C.byref(fb), C.byref(it), C.byref(t_opt), C.byref(t_grid)))
        # hessian_fallback > 0: the Richardson Hessian needed larger steps than numDeriv's default (opt-in retry;
        # the reference would have stopped in chol())
        self.optresults["hessian_fallback"] = fb.value
        self.diagnostics = {"hessian_fallback": fb.value, "grid_newton_iters": it.value, "opt_ms": t_opt.value,
                            "grid_ms": t_grid.value}
        owner = np.empty(self.K, dtype=np.int32)
        check(self._lib.bgp_fit_node_owner(self._h, owner.ctypes.data_as(_lib.c_int32_p)))
        self.node_owner = owner

    @property
    def lognormconst(self):
        return self.normalized_posterior["lognormconst"]

    def theta_moments(self):
        """Posterior mean and sd of theta by the quadrature itself, as ``summary(mod)`` prints them
        (aghq::compute_moment on ``normalized_posterior``; /root/reference/README.md:83-85)."""
        nw = self.normalized_posterior["nodesandweights"]
        lam = nw["weights"] * np.exp(nw["logpost_normalized"])
        mean = lam @ nw["theta"]
        var = lam @ (nw["theta"] - mean[None, :]) ** 2
        return mean, np.sqrt(var)

    @property
    def modesandhessians(self):
        if self._modes is None:
            modes = np.empty((self.K, self.p))
            Hs = np.empty((self.K, self.p, self.p))
            check(self._lib.bgp_fit_get_modes(self._h, dptr(modes), dptr(Hs)))
            self._modes, self._Hs = modes, Hs
        return {"theta": self.normalized_posterior["nodesandweights"]["theta"], "mode": self._modes, "H": self._Hs}

    def modesandhessians_view(self):
        """Zero-copy views of the fit's own page-locked host arrays (bgp_fit_host_arrays): valid until ``close()``.
        ``modesandhessians`` returns owned copies instead."""
        pm, pH = _lib.c_double_p(), _lib.c_double_p()
        check(self._lib.bgp_fit_host_arrays(self._h, C.byref(pm), C.byref(pH)))
        modes = np.ctypeslib.as_array(pm, shape=(self.K, self.p))
        Hs = np.ctypeslib.as_array(pH, shape=(self.K, self.p, self.p))
        return {"theta": self.normalized_posterior["nodesandweights"]["theta"], "mode": modes, "H": Hs}

    def close(self):
        if self._h:
            self._lib.bgp_fit_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def marginal_laplace_tmb(ff: LaplaceObjective, k: int, startingvalue, optresults=None) -> AGHQ:
    """aghq::marginal_laplace_tmb(ff, k, startingvalue)  (R/02_model_fit.R:284)."""
    lib = _lib.load()
    h = C.c_void_p()
    if optresults is None:
        th0 = fvec(startingvalue)
        check(lib.bgp_aghq_fit(ff._h, int(k), dptr(th0), C.byref(h)))
    else:
        mode = fvec(optresults["mode"])
        hess = fmat(np.atleast_2d(optresults["hessian"]))
        check(lib.bgp_aghq_fit_at(ff._h, int(k), dptr(mode), dptr(hess), C.byref(h)))
    return AGHQ(ff, h)


def sample_marginal(quad: AGHQ, M: int, Z=None, node_idx=None, seed: int = 0):
    """aghq::sample_marginal(quad, M)  (R/02_model_fit.R:687-689).  Returns ``{"samps": p x M,
    "theta": M x S, "node": M}``.  Pass (Z, node_idx) to fix the random inputs."""
    lib = _lib.load()
    samps = np.empty((quad.p, M), order="F")
    if Z is not None:
        Zf = fmat(Z)
        idx = np.ascontiguousarray(node_idx, dtype=np.int32)
        check(lib.bgp_sample(quad._h, M, dptr(Zf), idx.ctypes.data_as(_lib.c_int32_p), dptr(samps)))
    else:
        idx = np.empty(M, dtype=np.int32)
        check(lib.bgp_sample_draw(quad._h, M, int(seed), dptr(samps), idx.ctypes.data_as(_lib.c_int32_p)))
    theta = quad.normalized_posterior["nodesandweights"]["theta"][idx]
    # "resident": the same p x M matrix is still on the device behind `quad` (predict reads it in place) until the
    # next sample_marginal call on that object replaces it
    quad._resident_token = token = object()
    return {"samps": samps, "theta": theta, "node": idx, "resident": (quad, token, samps)}


class FitResult:
    """class(fit_result) <- "FitResult"  (R/02_model_fit.R:677-700)."""

    def __init__(self, instances, mod, ff, boundary_samp_indexes, random_samp_indexes, fixed_samp_indexes, family,
                 samps=None):
        self.instances = instances
        self.mod = mod
        self.ff = ff
        self.boundary_samp_indexes = boundary_samp_indexes
        self.random_samp_indexes = random_samp_indexes
        self.fixed_samp_indexes = fixed_samp_indexes
        self.family = family
        self.samps = samps

    def close(self):
        self.mod.close()
        self.ff.close()


def build_objective(y, terms: List[Term], fixed: Optional[Dict[str, np.ndarray]] = None, family="Gaussian", size=None,
                    control_family=None, control_fixed=None, device=0, shard=None, node_group=None):
    """get_result_by_method up to MakeADFun (R/02_model_fit.R:1-183,249-282): returns (ff, index maps).
    ``shard = (rank, world, nccl_unique_id)``: y / x / fixed are this rank's rows of an observation-sharded
    problem; the terms must then carry the GLOBAL knots / initial_location (IWP) or region / initial_location (sGP):
    defaults taken from the local rows would give every rank a different basis.
    ``node_group = (rank, world, nccl_unique_id)``: ranks holding the same rows split quadrature nodes, sample blocks
    and prediction rows (bgp_model_set_node_group)."""
    if family not in FAMILY_CODES:
        raise ValueError("family %r is outside the B200 hot path (Gaussian / Poisson / Binomial / none)" % family)
    if shard is not None and shard[1] > 1:
        for t in terms:
            if t.kind == "IWP" and (t.knots is None or t.initial_location is None):
                raise ValueError("observation-sharded model: IWP term %r needs explicit global knots and "
                                 "initial_location" % t.name)
            if t.kind == "sGP" and (t.region is None or t.initial_location is None):
                raise ValueError("observation-sharded model: sGP term %r needs an explicit global region and "
                                 "initial_location" % t.name)
            if t.kind == "IID":
                raise ValueError("observation-sharded model: IID term %r needs the global level set; build its "
                                 "design with add_random" % t.name)
    fixed = fixed or {}
    control_fixed = dict(control_fixed or {})
    control_family = control_family or {"u": 1.0, "alpha": 0.5}
    y = fvec(y)
    n = len(y)
    terms = [prepare_term(t) for t in terms]
    ff = LaplaceObjective(y=y, family=family, size=size, device=device)
    try:
        for t in terms:
            if t.kind == "IWP":
                # random + boundary block generated on the device (blocks keep the order of the calls)
                ff.add_iwp(t.x, t.initial_location, t.knots, t.order, t.u, t.alpha, t.boundary_prec, t.boundary_mean)
            elif t.kind == "sGP":
                # design AND precision (Compute_Q_sB) on the device from the covariate / the term's parameters
                ff.add_sgp_auto(t.x, t.initial_location, t.a, t.k, t.m, [float(t.region.min()), float(t.region.max())],
                                t.accuracy, t.u, t.alpha, t.boundary_prec, t.boundary_mean)
            else:
                B, Pd = iid_design(t)
                ff.add_random(B, Pd, 0.0, t.u, t.alpha)
        names = ["intercept"] + list(fixed)
        ff.add_fixed(np.ones(n), control_fixed.get("intercept", {}).get("prec", 0.01),
                     control_fixed.get("intercept", {}).get("mean", 0.0))
        for nm, col in fixed.items():
            cf = control_fixed.get(nm, {})
            ff.add_fixed(np.asarray(col, dtype=np.float64), cf.get("prec", 0.01), cf.get("mean", 0.0))
        if FAMILY_CODES[family] == 0:
            ff.set_noise_prior(control_family.get("u", 1.0), control_family.get("alpha", 0.5))
        if shard is not None:
            ff.set_shard(*shard)
        if node_group is not None:
            ff.set_node_group(*node_group)
        ff.finalize()
    except Exception:
        ff.close()
        raise
    # index maps (R/02_model_fit.R:627-675), 0-based
    rand_idx, bnd_idx, fix_idx = {}, {}, {}
    o = 0
    for t in terms:
        rand_idx[t.name] = np.arange(o, o + t.n_basis)
        o += t.n_basis
    for t in terms:
        if t.kind in ("IWP", "sGP"):
            bnd_idx[t.name] = np.arange(o, o + t.n_boundary)
            o += t.n_boundary
    for nm in names:
        fix_idx[nm] = o
        o += 1
    assert o == ff.p, (o, ff.p)
    return ff, terms, rand_idx, bnd_idx, fix_idx


def model_fit(y, terms: List[Term], fixed=None, method="aghq", family="Gaussian", control_family=None,
              control_fixed=None, aghq_k=4, size=None, M=3000, device=0, Z=None, node_idx=None, seed=0,
              optresults=None, shard=None, node_group=None) -> FitResult:
    """model_fit(formula, data, method = "aghq", family, ..., aghq_k = 4, M = 3000)."""
    if method != "aghq":
        raise ValueError("only method = 'aghq' is on the B200 hot path (nlminb / MCMC stay in R)")
    ff, terms, rand_idx, bnd_idx, fix_idx = build_objective(y, terms, fixed, family, size, control_family,
                                                            control_fixed, device, shard, node_group)
    if ff.S == 0:
        ff.close()
        raise ValueError("For model with no hyper-parameter, the method cannot be aghq or MCMC.")
    mod = marginal_laplace_tmb(ff, aghq_k, np.zeros(ff.S), optresults)
    res = FitResult(terms, mod, ff, bnd_idx, rand_idx, fix_idx, family)
    res.control_family = dict(control_family) if control_family else {"u": 1.0, "alpha": 0.5}   # sd.prior of the noise
    if M:
        res.samps = sample_marginal(mod, M, Z, node_idx, seed)
    return res


def model_fit_loop(loop_values, fit_args, prior_func=None, parallel=False, group=None):
    """model_fit_loop(loop_holder, loop_values, prior_func, parallel, ...)  (R/02_model_fit.R:725-778): one model_fit
    per value of the looping variable, the log marginal likelihoods (``mod$normalized_posterior$lognormconst``) and the
    posterior of the variable normalised by ``sfsmisc::integrate.xy``.

    ``fit_args(value)`` returns the keyword arguments of ``model_fit`` for that value (the R version substitutes the
    value for the ``LOOP`` placeholder inside the call).  ``parallel = TRUE`` in the reference is a PSOCK cluster of
    whole-model workers; here it is one process per GPU under ``torch.distributed``: rank r fits the values
    r, r + world, ... on its own device and the log marginal likelihoods are all-gathered (``group``: the process
    group, default the world)."""
    from .post_fit import integrate_xy
    loop_values = np.asarray(loop_values, dtype=np.float64)
    L = len(loop_values)
    log_ml = np.zeros(L)
    rank, world = 0, 1
    if parallel:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    for i in range(rank, L, world):
        kw = dict(fit_args(float(loop_values[i])))
        kw.setdefault("M", 0)                    # only lognormconst is read (R/02_model_fit.R:751)
        res = model_fit(**kw)
        log_ml[i] = res.mod.lognormconst
        res.close()
    if parallel and world > 1:
        import torch
        import torch.distributed as dist
        dev = (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl"
               else torch.device("cpu"))
        t = torch.from_numpy(log_ml).to(dev)
        dist.all_reduce(t, group=group)          # every value has one owner, the others contribute zeros
        log_ml = t.cpu().numpy()
    prior = np.ones(L) if prior_func is None else np.asarray(prior_func(loop_values), dtype=np.float64) * np.ones(L)
    log_joint = log_ml + np.log(prior)
    log_joint = log_joint - np.max(log_joint)
    post = np.exp(log_joint)
    post = post / integrate_xy(loop_values, post)
    return {"var": loop_values, "post": post, "log_ml": log_ml}


def sample_fixed_effect(model_fit_result: FitResult, variables):
    """R/03_post_fit.R:159-165."""
    samps = model_fit_result.samps["samps"]
    return samps[[model_fit_result.fixed_samp_indexes[v] for v in variables], :].T


def compute_post_fun_IWP(samps, global_samps=None, knots=None, refined_x=None, p=None, degree=0, intercept_samps=None,
                         level=0.95, only_samples=False, device=0):
    """R/03_post_fit.R:200-241 fused with extract_mean_interval_given_samps (:287-296)."""
    lib = _lib.load()
    samps = fmat(samps)
    M = samps.shape[1]
    if p <= degree:
        print("Error: The degree of derivative to compute is not defined. Should consider higher order smoothing "
              "model or lower order of the derivative degree.")
        return None
    if global_samps is not None and np.asarray(global_samps).reshape(-1, M).shape[0] != p - 1:
        print("Error: Incorrect dimension of global_samps. Check whether the choice of p is consistent with the "
              "fitted model.")
        return None
    gs = None if global_samps is None or p == 1 else fmat(np.asarray(global_samps).reshape(-1, M))
    ic = None if intercept_samps is None else fvec(intercept_samps)
    x = fvec(refined_x)
    G = len(x)
    kn = fvec(knots)
    mean, lo, hi = np.empty(G), np.empty(G), np.empty(G)
    F = np.empty((G, M), order="F") if only_samples else None
    check(lib.bgp_predict_iwp(dptr(samps), dptr(gs), dptr(ic), M, dptr(kn), len(kn), int(p), int(degree), dptr(x), G,
                              float(level), device, dptr(mean), dptr(lo), dptr(hi), dptr(F)))
    out = {"x": x, "plower": lo, "pupper": hi, "mean": mean}
    if only_samples:
        out["samples"] = F
    return out


def compute_post_fun_sGP(samps, global_samps=None, k=None, refined_x=None, a=None, region=None, boundary=True, m=1,
                         intercept_samps=None, level=0.95, only_samples=False, device=0):
    """R/03_post_fit.R:261-276 fused with extract_mean_interval_given_samps."""
    lib = _lib.load()
    samps = fmat(samps)
    M = samps.shape[1]
    gs = None if global_samps is None else fmat(np.asarray(global_samps).reshape(-1, M))
    ic = None if intercept_samps is None else fvec(intercept_samps)
    x = fvec(refined_x)
    G = len(x)
    reg = fvec(region)
    mean, lo, hi = np.empty(G), np.empty(G), np.empty(G)
    F = np.empty((G, M), order="F") if only_samples else None
    check(lib.bgp_predict_sgp(dptr(samps), dptr(gs), dptr(ic), M, float(a), int(k), int(m), dptr(reg), int(boundary),
                              dptr(x), G, float(level), device, dptr(mean), dptr(lo), dptr(hi), dptr(F)))
    out = {"x": x, "plower": lo, "pupper": hi, "mean": mean}
    if only_samples:
        out["samples"] = F
    return out


def _predict_resident(object: FitResult, term, variable, refined_x, degree, include_intercept, level):
    """compute_post_fun_IWP / compute_post_fun_sGP + extract_mean_interval_given_samps on the device-resident
    samples (bgp_fit_predict_*): same rows of samps$samps as the host path selects (R/03_post_fit.R:65-76)."""
    lib = _lib.load()
    x = fvec(refined_x)
    G = len(x)
    mean, lo, hi = np.empty(G), np.empty(G), np.empty(G)
    r0 = int(object.random_samp_indexes[variable][0])
    bidx = object.boundary_samp_indexes.get(variable)
    g0 = int(bidx[0]) if bidx is not None and len(bidx) else -1
    i0 = int(object.fixed_samp_indexes["intercept"]) if include_intercept else -1
    h = object.mod._h
    if term.kind == "IWP":
        if term.order <= degree:
            print("Error: The degree of derivative to compute is not defined. Should consider higher order smoothing "
                  "model or lower order of the derivative degree.")
            return None
        kn = fvec(term.knots)
        check(lib.bgp_fit_predict_iwp(h, r0, g0, i0, dptr(kn), len(kn), int(term.order), int(degree), dptr(x), G,
                                      float(level), dptr(mean), dptr(lo), dptr(hi)))
    else:
        reg = fvec(term.region)
        check(lib.bgp_fit_predict_sgp(h, r0, g0, i0, float(term.a), int(term.k), int(term.m), dptr(reg),
                                      int(term.boundary), dptr(x), G, float(level), dptr(mean), dptr(lo), dptr(hi)))
    return {"x": x, "plower": lo, "pupper": hi, "mean": mean}


def predict(object: FitResult, newdata=None, variable=None, degree=0, include_intercept=True, only_samples=False,
            level=0.95):
    """predict.FitResult(object, newdata, variable, degree, include.intercept, only.samples)."""
    names = list(object.random_samp_indexes)
    if names.count(variable) == 0:
        raise ValueError("The specified variable cannot be found in the fitted model, please check the name.")
    samps = object.samps["samps"]
    term = next(t for t in object.instances if t.name == variable)
    global_samps = samps[object.boundary_samp_indexes[variable], :] if variable in object.boundary_samp_indexes else None
    coefsamps = samps[object.random_samp_indexes[variable], :]
    if newdata is None:
        refined_x = term.observed_x
    else:
        if isinstance(newdata, dict):          # R: newdata[[variable]]  (R/03_post_fit.R:73)
            newdata = newdata[variable]
        refined_x = np.sort(np.asarray(newdata, dtype=np.float64) - term.initial_location)
    intercept = samps[object.fixed_samp_indexes["intercept"], :] if include_intercept else None
    dev = object.ff.device
    quad, token, arr = object.samps.get("resident") or (None, None, None)
    if (not only_samples and quad is not None and quad is object.mod and quad._h and term.kind in ("IWP", "sGP")
            and getattr(quad, "_resident_token", None) is token and object.samps["samps"] is arr):
        # the samples are still on the device behind the fit: no p x M host copy, grid rows split over the node group
        f = _predict_resident(object, term, variable, refined_x, degree, include_intercept, level)
        if f is None:
            return None
        f["x"] = f["x"] + term.initial_location
        return f
    if term.kind == "IWP":
        f = compute_post_fun_IWP(coefsamps, global_samps, term.knots, refined_x, term.order, degree, intercept, level,
                                 only_samples, dev)
    elif term.kind == "sGP":
        f = compute_post_fun_sGP(coefsamps, global_samps, term.k, refined_x, term.a, term.region, term.boundary,
                                 term.m, intercept, level, only_samples, dev)
    else:
        raise ValueError("predict supports IWP and sGP terms")
    if f is None:
        return None
    f["x"] = f["x"] + term.initial_location
    return f
