"""GPU parity on the edge cases of the path: IWP order 1 (no boundary columns, R/02_model_fit.R:460,651-652),
mixed-sign knots (mirrored O-spline blocks, R/01_utility.R:378-401), an IID term (one-hot design with
structural zeros, R/01_utility.R:214-219), Binomial with the default size = 1 (R/02_model_fit.R:176-183),
n and p that are not multiples of any tile size, derivatives in predict, and the error behaviour of the ABI."""
import numpy as np
import pytest

from helpers import relerr, tmbdata_from_oracle

pytestmark = pytest.mark.gpu


def _compare(model, thetas, tol_val=1e-8):
    from bayesgp_b200 import make_objective
    from oracle.laplace import LaplaceObjective as OFF
    off = OFF(model)
    ff = make_objective(tmbdata_from_oracle(model))
    try:
        rng = np.random.default_rng(5)
        W = 0.05 * rng.standard_normal(model.p)
        o = model.objective(W, thetas[0], "fgH")
        f, g, H = ff.objective(W, thetas[0], want_grad=True, want_hess=True)
        assert abs(f - o["f"]) <= 1e-11 * abs(o["f"])
        assert relerr(g, o["g"]) < 1e-10 and relerr(H, o["H"]) < 1e-10
        for th in thetas:
            want = off.fn(th)
            got, _, w, Hm = ff._eval(th, want_hess=True)
            assert abs(got - want) <= tol_val * abs(want), (th, got, want)
            assert relerr(w, off.last_par) < 1e-6 and relerr(Hm, off.sp_hess()) < 1e-6
            gw, gg = off.gr(th), ff.gr(th)
            assert np.max(np.abs(gw - gg)) <= 2e-6 * max(1.0, np.max(np.abs(gw))), (th, gw, gg)
    finally:
        ff.close()


def test_iwp_order_one_has_no_boundary_block():
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(11)
    n = 1237                                     # not a multiple of 8 / 16 / 64
    x = rng.uniform(0, 3, n)
    y = rng.poisson(np.exp(0.3 + np.sin(x))).astype(np.float64)
    model = build_model(y, [Term("IWP", "x", x, order=1, k=23)], {}, family="Poisson")[0]
    assert model.X[0].shape[1] == 0 and model.p == 22 + 1
    _compare(model, [np.array([0.0]), np.array([2.5])])


def test_mixed_sign_knots_and_fixed_effects():
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(12)
    n = 3001
    x = rng.uniform(-1.0, 2.0, n)
    z = rng.standard_normal(n)
    y = rng.poisson(np.exp(0.2 + 0.5 * np.cos(2 * x) + 0.1 * z)).astype(np.float64)
    knots = np.array([-1.0, -0.6, -0.3, -0.1, 0.0, 0.2, 0.5, 0.9, 1.4, 2.0])
    model = build_model(y, [Term("IWP", "x", x, order=3, knots=knots, initial_location=0.0)], {"z": z}, family="Poisson")[0]
    _compare(model, [np.array([0.0]), np.array([-1.5])])


@pytest.mark.parametrize("k", [400, 470])
def test_designs_of_seven_and_eight_column_groups(k):
    """384 < p <= 512: the likelihood pass with two observations per warp on 8-observation stages (ring depths 6
    and 5, the odd one included), the Hessian and Cholesky at panel counts that are not a power of two."""
    from helpers import synth_poisson
    model = synth_poisson(n=15001, k=k, order=3, seed=77 + k)[0]
    assert (model.p + 63) // 64 in (7, 8)
    _compare(model, [np.array([-4.0]), np.array([-4.6])])


def test_iid_term_and_default_binomial_size():
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(13)
    n = 2500
    grp = rng.integers(0, 37, n).astype(np.float64)
    x = rng.uniform(0, 1, n)
    eff = rng.standard_normal(37) * 0.7
    eta = -0.2 + eff[grp.astype(int)] + np.sin(3 * x)
    y = rng.binomial(1, 1 / (1 + np.exp(-eta))).astype(np.float64)
    model = build_model(y, [Term("IID", "g", grp), Term("IWP", "x", x, order=2, k=12)], {}, family="Binomial")[0]
    assert model.size is not None and np.all(model.size == 1.0)
    _compare(model, [np.array([0.0, 0.0]), np.array([1.0, 2.0])])


@pytest.mark.parametrize("a,k,m,region,acc", [(2 * np.pi * 3, 9, 2, (0.0, 1.5), 0.01), (2 * np.pi * 5, 20, 1, (0.0, 1.0), 0.01),
                                              (1.7, 4, 1, (0.2, 2.0), 0.05), (2 * np.pi, 31, 3, (-0.5, 0.75), 0.002)])
def test_sgp_precision_on_the_device_matches_compute_q_sb(a, k, m, region, acc):
    """Compute_Q_sB (R/01_utility.R:67-174) on the device — one Gram GEMM of the nine basis families + the block
    formulas — against the oracle's restatement, and determinant(P)$modulus against numpy's slogdet."""
    import ctypes as C
    from bayesgp_b200 import _lib
    from oracle import basis as ob
    lib = _lib.load()
    d = 3 * (k - 2) * m
    P = np.empty((d, d), order="F")
    reg = np.array(region, dtype=np.float64)
    ld = C.c_double()
    _lib.check(lib.bgp_sgp_precision(a, k, m, _lib.dptr(reg), acc, 0, _lib.dptr(P), C.byref(ld)))
    want = ob.compute_P_sGP(a, k, m, reg, acc)
    assert np.array_equal(P, P.T)
    assert relerr(P, want) < 1e-11
    # determinant(P)$modulus is only as well defined as P is conditioned: the k = 31, m = 3 Gram matrix has cond ~1e16
    # (Compute_Q_sB is numerically singular at fine bases) and a 1e-15 relative perturbation of P
    # moves its LU log-determinant by ~0.05; the well-conditioned cases are held to 1e-6
    tol = 1e-6 * max(1.0, abs(ld.value)) if np.linalg.cond(want) < 1e9 else 0.5
    assert abs(ld.value - np.linalg.slogdet(want)[1]) <= tol


def test_gradient_with_a_wide_dense_precision_block():
    """ff$gr with an sGP term of 84 columns (two harmonics of 42, dense block-diagonal P): the block quadratic forms and
    traces of the one-CTA gradient algebra (grad.cu) against the oracle's closed form."""
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(15)
    n = 6000
    x = rng.uniform(0, 3, n)
    z = rng.uniform(0, 1, n)
    eta = 0.2 + 0.7 * np.sin(2 * np.pi * x) + 0.3 * np.sin(4 * np.pi * x) + np.cos(3 * z)
    y = rng.poisson(np.exp(eta)).astype(np.float64)
    terms = [Term("sGP", "x", x, a=2 * np.pi, k=16, m=2, region=np.array([0.0, 3.0]), initial_location=0.0),
             Term("IWP", "z", z, order=2, k=15)]
    model = build_model(y, terms, {}, family="Poisson")[0]
    assert model.B[0].shape[1] == 84
    _compare(model, [np.array([1.0, -2.0]), np.array([0.4, -1.2])])


def test_predict_derivatives_match_oracle():
    """f, f', f'' from the same samples (degree < order, R/03_post_fit.R:201-203 rejects the rest)."""
    import bayesgp_b200 as bg
    from oracle import fit as ofit
    rng = np.random.default_rng(14)
    knots = np.linspace(0, 2.0, 17)
    M, G, order = 700, 333, 3
    coef = rng.standard_normal((16, M)) * 0.3
    glob = rng.standard_normal((order - 1, M))
    icpt = rng.standard_normal(M)
    xg = np.sort(rng.uniform(0, 2.0, G))
    for degree in (0, 1, 2):
        out = bg.compute_post_fun_IWP(coef, glob, knots, xg, order, degree, icpt)
        F = ofit.compute_post_fun_IWP(coef, glob, knots, xg, order, degree, icpt)
        lo, hi, mean = ofit.extract_mean_interval_given_samps(F)
        assert relerr(out["mean"], mean) < 1e-10
        assert relerr(out["plower"], lo) < 1e-10 and relerr(out["pupper"], hi) < 1e-10
    assert bg.compute_post_fun_IWP(coef, glob, knots, xg, order, 3, icpt) is None      # degree >= order


def test_abi_error_behaviour():
    import ctypes as C
    from bayesgp_b200 import _lib
    from bayesgp_b200.objective import LaplaceObjective
    lib = _lib.load()
    y = np.arange(10, dtype=np.float64)
    with pytest.raises(_lib.BgpError):
        LaplaceObjective(y=y, family=3)                         # Coxph: outside the path
    ff = LaplaceObjective(y=y, family="Poisson")
    with pytest.raises(_lib.BgpError):
        ff.fn(np.array([0.0]))                                  # not finalized
    with pytest.raises(_lib.BgpError):
        ff.finalize()                                           # no design columns
    ff.close()
    # latent dimension above the supported maximum is refused, not truncated
    ff = LaplaceObjective(y=np.zeros(40), family="Poisson")
    ff.add_fixed(np.ones(40))
    ff.add_random(np.zeros((40, 1100)), np.ones(1100), 0.0)
    with pytest.raises(_lib.BgpError) as ei:
        ff.finalize()
    assert "exceeds" in str(ei.value)
    ff.close()
    # a non positive definite inner Hessian comes back as NaN (TMB behaviour), not as an exception
    ff = LaplaceObjective(y=np.array([1.0, 2.0, 0.0, 3.0]), family="Poisson")
    ff.add_random(np.zeros((4, 2)), np.array([-1.0, 1.0]), 0.0)    # indefinite prior precision, no data information
    ff.add_fixed(np.ones(4))
    ff.finalize()
    assert np.isnan(ff.fn(np.array([0.0])))
    ff.close()
