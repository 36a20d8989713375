import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bench
from bayesgp_b200.workloads import c3_data
x, y = c3_data(1000000)
ff = bench.build_b200(x, y, 0)
W = np.zeros(ff.p); th = np.array([-10.5])
for _ in range(3): ff.objective(W, th, want_grad=True)
t0 = ff.last_timing()
for _ in range(20): ff.objective(W, th, want_grad=True)
t1 = ff.last_timing()
print("lik ms per pass", (t1["lik_ms"] - t0["lik_ms"]) / (t1["lik_launches"] - t0["lik_launches"]))
