"""Repeats the revisit flow of tests/test_gpu_largep.py::test_gaussian_above_512_columns (evaluate at theta, then the
gradient at the same theta) to measure how often the factorisation at the mode reports a non-positive pivot."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
from helpers import tmbdata_from_oracle
from oracle.fit import Term, build_model
from bayesgp_b200 import make_objective
rng = np.random.default_rng(77)
n, k = 20000, 199
xs = [rng.uniform(0, 1, n) for _ in range(3)]
eta = 0.5 + np.sin(2 * np.pi * xs[0]) + 0.4 * np.sin(3 * np.pi * xs[1]) + 0.6 * np.sin(2.5 * np.pi * xs[2] + 1.0)
y = eta + 0.3 * rng.standard_normal(n)
model = build_model(y, [Term("IWP", "x%d" % (i + 1), xs[i], order=3, k=k) for i in range(3)], {}, family="Gaussian")[0]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
msgs = {}
ff = make_objective(tmbdata_from_oracle(model))
ths = [np.array([-3.0, -3.5, -4.0, 2.0]), np.array([-3.4, -3.2, -4.3, 2.4])]
for r in range(reps):
    th = ths[r % 2]
    v = ff._eval(th, want_hess=True)[0]
    g = ff.gr(th)
    if not np.all(np.isfinite(g)) or not np.isfinite(v):
        bad += 1
        msgs[getattr(ff, "last_warning", "?")] = msgs.get(getattr(ff, "last_warning", "?"), 0) + 1
ff.close()
print("env", {k: v for k, v in os.environ.items() if k.startswith("BGP_")}, "reps", reps, "failures", bad, msgs)
