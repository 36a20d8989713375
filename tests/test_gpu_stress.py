"""GPU robustness: degenerate sizes (n < one tile, p around the 16 / 64 column boundaries), two models alive at
once, and create / destroy cycles without device-memory growth."""
import numpy as np
import pytest

from helpers import relerr, tmbdata_from_oracle

pytestmark = pytest.mark.gpu


def _model(n, k, seed, order=2, fixed_cols=0):
    from oracle.fit import Term, build_model
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, n)
    fixed = {"z%d" % i: rng.standard_normal(n) for i in range(fixed_cols)}
    y = rng.poisson(np.exp(0.5 + np.sin(3 * x))).astype(np.float64)
    return build_model(y, [Term("IWP", "x", x, order=order, k=k)], fixed, family="Poisson")[0]


@pytest.mark.parametrize("n,k,fixed_cols", [(5, 3, 0), (63, 5, 1), (64, 15, 0), (65, 16, 0), (1000, 63, 0), (1000, 64, 1),
                                            (257, 130, 2)])
def test_degenerate_sizes(n, k, fixed_cols):
    from bayesgp_b200 import make_objective
    from oracle.laplace import LaplaceObjective as OFF
    model = _model(n, k, seed=100 + n + k, fixed_cols=fixed_cols)
    off = OFF(model)
    ff = make_objective(tmbdata_from_oracle(model))
    try:
        W = 0.1 * np.random.default_rng(1).standard_normal(model.p)
        th = np.array([1.0])
        o = model.objective(W, th, "fgH")
        f, g, H = ff.objective(W, th, want_grad=True, want_hess=True)
        assert abs(f - o["f"]) <= 1e-11 * max(1.0, abs(o["f"]))
        assert relerr(g, o["g"]) < 1e-10 and relerr(H, o["H"]) < 1e-10
        want = off.fn(th)
        got, _, w, _ = ff._eval(th)
        assert abs(got - want) <= 1e-8 * abs(want)
        assert relerr(w, off.last_par) < 1e-6
    finally:
        ff.close()


def test_two_models_interleaved():
    from bayesgp_b200 import make_objective
    from oracle.laplace import LaplaceObjective as OFF
    ma, mb = _model(3000, 40, 7), _model(2000, 25, 8, order=3)
    fa, fb = make_objective(tmbdata_from_oracle(ma)), make_objective(tmbdata_from_oracle(mb))
    oa, ob = OFF(ma), OFF(mb)
    try:
        for th in (0.0, 2.0, 4.0):
            va, vb = fa.fn(np.array([th])), fb.fn(np.array([th + 1.0]))
            assert abs(va - oa.fn(np.array([th]))) <= 1e-8 * abs(va)
            assert abs(vb - ob.fn(np.array([th + 1.0]))) <= 1e-8 * abs(vb)
    finally:
        fa.close()
        fb.close()


def test_create_destroy_does_not_leak_device_memory():
    import torch
    from bayesgp_b200 import make_objective
    model = _model(20000, 60, 9)
    data = tmbdata_from_oracle(model)

    def cycle():
        ff = make_objective(data)
        ff.fn(np.array([1.0]))
        ff.gr(np.array([1.0]))
        ff.close()

    cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(20):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 8 << 20, (free0, free1)
