"""Short, fixed launch sequence for ncu: build C3, a warm-up batch, then — inside the profiler range — two Laplace
evaluations and one gradient (ncu --profile-from-start off captures only the range)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bench
from bayesgp_b200 import _lib
from bayesgp_b200.workloads import c3_data
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
x, y = c3_data(n)
ff = bench.build_b200(x, y, 0)
thetas = np.array([[-10.9], [-10.7], [-10.5], [-10.3]])
ff.set_start(None)
ff.fn_batch(thetas[:2], want_modes=False)
ff.gr(thetas[1])                       # allocates the gradient plan outside the range
lib = _lib.load()
lib.bgp_profiler_range(1)
vals, _, _, iters = ff.fn_batch(thetas[2:], want_modes=False)
g = ff.gr(thetas[-1] + 0.05)
lib.bgp_profiler_range(0)
print("values", vals, "iters", iters, "grad", g, "timing", ff.last_timing())
