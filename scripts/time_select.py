import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
from bayesgp_b200.api import compute_post_fun_IWP
rng = np.random.default_rng(1)
M, G = 10000, 20000
knots = np.linspace(0, 1, 300)
coef = 0.05 * rng.standard_normal((299, M)); glob = rng.standard_normal((2, M)); icpt = rng.standard_normal(M)
xg = np.linspace(0, 1, G)
kw = dict(global_samps=glob, knots=knots, refined_x=xg, p=3, degree=0, intercept_samps=icpt)
out = compute_post_fun_IWP(coef, **kw)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); out = compute_post_fun_IWP(coef, **kw); ts.append(time.perf_counter() - t0)
print("best ms", min(ts) * 1e3, "sum lo", out["plower"].sum(), "sum hi", out["pupper"].sum())
