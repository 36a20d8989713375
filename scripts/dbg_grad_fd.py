"""Laplace gradient at full C3 size against finite differences of the values (two stencils), for A/B runs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bench
from bayesgp_b200.workloads import c3_data
x, y = c3_data(1000000)
ff = bench.build_b200(x, y, 0)
th = np.array([-10.5])
g = ff.gr(th)
ff.set_factor_reuse(False)
v0 = ff.fn(th)
for eps in (1e-4, 1e-3, 2e-3, 4e-3):
    f = {k: ff.fn(th + k * eps) for k in (-2, -1, 1, 2)}
    fd2 = (f[1] - f[-1]) / (2 * eps)
    fd4 = (f[-2] - 8 * f[-1] + 8 * f[1] - f[2]) / (12 * eps)
    print("eps %g  g %.10f  fd2 %.10f (%.2e)  fd4 %.10f (%.2e)  value %.6f" % (eps, g[0], fd2, g[0] - fd2, fd4, g[0] - fd4, v0))
