"""Short, fixed launch sequence for ncu: build C3, one warm-up pass, then 4 Laplace evaluations."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bench
from bayesgp_b200.workloads import c3_data
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
x, y = c3_data(n)
ff = bench.build_b200(x, y, 0)
thetas = np.array([[-10.9], [-10.7], [-10.5], [-10.3]])
ff.set_start(None)
ff.fn_batch(thetas[:2], want_modes=False)
vals, _, _, iters = ff.fn_batch(thetas, want_modes=False)
g = ff.gr(thetas[-1])
print("values", vals, "iters", iters, "grad", g, "timing", ff.last_timing())
