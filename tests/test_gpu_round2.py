"""GPU tests of the round-2 entry points: the optimiser end state (mode + tangent) capture / restore, the fit's
page-locked host arrays against the copying getter, prediction from the device-resident samples against the
host-buffer path, the fit diagnostics, and the node-group API on a trivial (one-rank) group."""
import numpy as np
import pytest

from helpers import relerr, synth_poisson, tmbdata_from_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_fit():
    import bayesgp_b200 as bg
    rng = np.random.default_rng(31)
    n = 5000
    x = rng.uniform(0, 1, n)
    y = rng.poisson(np.exp(1.0 + np.sin(2 * np.pi * x))).astype(np.float64)
    res = bg.model_fit(y, [bg.Term("IWP", "x", x, order=3, k=24)], {}, family="Poisson", aghq_k=5, M=800, seed=3)
    yield res, x
    res.close()


def test_start_state_restores_the_optimisers_end_state():
    from bayesgp_b200 import make_objective
    model = synth_poisson(n=20000, k=40, order=3)[0]
    ff = make_objective(tmbdata_from_oracle(model))
    try:
        th0 = np.array([-4.0])
        v0 = ff.fn(th0)
        w0 = ff.env.last_par.copy()
        th_rec, T = ff.get_tangent()
        assert np.array_equal(th_rec, th0) and T.shape == (ff.p, 1) and np.all(np.isfinite(T))
        # the tangent is d w_hat / d theta: a central difference of the modes
        h = 1e-4
        ff.fn(th0 + h)
        wp = ff.env.last_par.copy()
        ff.fn(th0 - h)
        wm = ff.env.last_par.copy()
        fd = (wp - wm) / (2 * h)
        assert relerr(T[:, 0], fd) < 1e-5
        # restoring (theta, mode, tangent) gives the same values as a cold history, in no more Newton iterations
        th1 = np.array([-3.7])
        ff.set_start(w0)
        it0 = ff.newton_iters
        plain = ff.fn(th1)
        it_plain = ff.newton_iters - it0
        ff.set_start_at(th0, w0, T)
        it0 = ff.newton_iters
        warm = ff.fn(th1)
        it_warm = ff.newton_iters - it0
        assert abs(plain - warm) <= 1e-10 * abs(plain)
        assert it_warm <= it_plain
        assert abs(ff.fn(th0) - v0) <= 1e-10 * abs(v0)
    finally:
        ff.close()


def test_host_arrays_equal_the_copying_getter_and_diagnostics(small_fit):
    res, _ = small_fit
    mod = res.mod
    mh = mod.modesandhessians
    view = mod.modesandhessians_view()
    assert np.array_equal(mh["mode"], view["mode"]) and np.array_equal(mh["H"], view["H"])
    assert mh["H"].shape == (mod.K, mod.p, mod.p)
    for j in range(mod.K):
        assert np.array_equal(mh["H"][j], mh["H"][j].T)
    d = mod.diagnostics
    assert d["hessian_fallback"] == 0 and d["grid_newton_iters"] >= 0 and d["grid_ms"] > 0 and d["opt_ms"] > 0
    assert mod.optresults["hessian_fallback"] == 0
    assert mod.node_owner.tolist() == [0] * mod.K          # no node group: every node on this rank


def test_resident_predict_equals_host_buffer_path(small_fit):
    import bayesgp_b200 as bg
    res, x = small_fit
    xg = np.linspace(x.min(), x.max(), 777)
    assert res.samps.get("resident") is not None
    for degree in (0, 1, 2):
        dev = bg.predict(res, newdata=xg, variable="x", degree=degree)
        host_samps = {k: v for k, v in res.samps.items() if k != "resident"}
        res_h = bg.FitResult(res.instances, res.mod, res.ff, res.boundary_samp_indexes, res.random_samp_indexes,
                             res.fixed_samp_indexes, res.family, host_samps)
        host = bg.predict(res_h, newdata=xg, variable="x", degree=degree)
        for key in ("x", "mean", "plower", "pupper"):
            assert np.array_equal(dev[key], host[key]), (degree, key)
    # a replaced sample matrix must not be served from the device copy
    other = dict(res.samps, samps=np.asfortranarray(res.samps["samps"] * 1.5))
    res_o = bg.FitResult(res.instances, res.mod, res.ff, res.boundary_samp_indexes, res.random_samp_indexes,
                         res.fixed_samp_indexes, res.family, other)
    scaled = bg.predict(res_o, newdata=xg, variable="x", degree=0)
    base = bg.predict(res, newdata=xg, variable="x", degree=0)
    assert relerr(scaled["mean"], 1.5 * base["mean"]) < 1e-12


def test_trivial_node_group_and_failing_nodes():
    """A one-rank node group changes nothing; a grid that reaches a theta where the inner solve fails is an error
    (aghq stops there), not a NaN-filled fit."""
    import bayesgp_b200 as bg
    from bayesgp_b200 import _lib
    from bayesgp_b200.objective import LaplaceObjective
    rng = np.random.default_rng(5)
    n = 3000
    x = rng.uniform(0, 1, n)
    y = rng.poisson(np.exp(0.5 + np.cos(3 * x))).astype(np.float64)
    knots = np.linspace(0, x.max() - x.min(), 12)

    def build(group):
        ff = LaplaceObjective(y=y, family="Poisson")
        ff.add_iwp(x, float(x.min()), knots, 2)
        ff.add_fixed(np.ones(n))
        if group:
            ff.set_node_group(0, 1, bytes(128))
        return ff.finalize()

    a, b = build(False), build(True)
    try:
        opt = {"mode": np.array([1.0]), "hessian": np.array([[4.0]])}
        ma = bg.marginal_laplace_tmb(a, 5, None, optresults=opt)
        mb = bg.marginal_laplace_tmb(b, 5, None, optresults=opt)
        assert ma.lognormconst == mb.lognormconst
        assert np.array_equal(ma.modesandhessians["H"], mb.modesandhessians["H"])
        ma.close()
        mb.close()
        # a grid so wide that exp(theta) overflows at its outer nodes
        with pytest.raises(_lib.BgpError) as ei:
            bg.marginal_laplace_tmb(a, 5, None, optresults={"mode": np.array([0.0]), "hessian": np.array([[1e-6]])})
        assert "quadrature node" in str(ei.value)
    finally:
        a.close()
        b.close()


def test_model_fit_loop_matches_the_oracle():
    """model_fit_loop (R/02_model_fit.R:725-778): one fit per value of the looped variable (here the period of an sGP
    term, the package's own use case), log marginal likelihoods and the normalised posterior of the variable."""
    import bayesgp_b200 as bg
    from oracle import fit as ofit
    rng = np.random.default_rng(41)
    n = 1500
    x = rng.uniform(0, 4, n)
    y = rng.poisson(np.exp(0.3 + 0.8 * np.sin(2 * np.pi * x / 1.0))).astype(np.float64)
    periods = np.array([0.8, 0.9, 1.0, 1.1, 1.25])
    region = np.array([0.0, 4.0])

    def args(T, Term):
        return dict(y=y, terms=[Term("sGP", "x", x, a=2 * np.pi / T, k=8, m=1, region=region, initial_location=0.0)],
                    fixed={}, family="Poisson", aghq_k=3)

    prior = lambda v: np.exp(-0.5 * ((v - 1.0) / 0.5) ** 2)
    got = bg.model_fit_loop(periods, lambda T: args(T, bg.Term), prior_func=prior)
    want = ofit.model_fit_loop(periods, lambda T: args(T, ofit.Term), prior_func=prior)
    assert np.max(np.abs(got["log_ml"] - want["log_ml"]) / np.abs(want["log_ml"])) < 2e-7     # own BFGS on both sides
    assert relerr(got["post"], want["post"]) < 1e-4
    assert int(np.argmax(got["post"])) == 2                                                  # the true period
