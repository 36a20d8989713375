import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bayesgp_b200 as bg
from bayesgp_b200 import api, _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000
rng = np.random.default_rng(20244)
x1, x2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
eta = -0.3 + np.sin(2 * np.pi * x1) + 0.6 * np.cos(2 * np.pi * 5 * x2)
size = 1.0 + rng.poisson(9, n)
y = rng.binomial(size.astype(int), 1 / (1 + np.exp(-eta))).astype(np.float64)
terms = [bg.Term("IWP", "x1", x1, order=2, k=200), bg.Term("sGP", "x2", x2, a=2 * np.pi * 5, k=int(sys.argv[2]) if len(sys.argv) > 2 else 100, m=1, region=np.array([0.0, 1.0]), accuracy=0.01)]
t0 = time.time()
ff, terms, *_ = api.build_objective(y, terms, {}, "Binomial", size)
print("build", time.time() - t0, "p", ff.p, "S", ff.S, ff.hessian_flops())
lib = _lib.load()
for th in ([0.0, 0.0], [2.0, 2.0], [5.0, 5.0], [-2.0, -2.0]):
    ff.set_start(None)
    v = ff.fn(np.array(th))
    print(th, v, "err:", lib.bgp_last_error().decode(), "iters", ff.newton_iters, ff.last_timing())
f, g, H = ff.objective(np.zeros(ff.p), np.array([0.0, 0.0]), True, True)
print("f", f, "gmax", np.abs(g).max(), "H finite", np.isfinite(H).all(), "min eig", np.linalg.eigvalsh(H)[:3], "max", np.linalg.eigvalsh(H)[-1])
