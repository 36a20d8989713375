"""Trace of model_fit's optimisation phase on C3 (BGP_FIT_DEBUG=1): every ff evaluation with its inner iteration counts."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bench
import bayesgp_b200 as bg
from bayesgp_b200.workloads import c3_data
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
x, y = c3_data(n)
ff = bench.build_b200(x, y, 0)
t0 = time.time()
mod = bg.marginal_laplace_tmb(ff, 15, np.zeros(ff.S))
print("fit %.3f s" % (time.time() - t0), mod.optresults["mode"], mod.optresults["fn_count"], mod.optresults["gr_count"], mod.diagnostics)
