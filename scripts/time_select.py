"""Quantile-select timing / exactness: fast path vs BGP_SELECT_RADIX=1.  usage: time_select.py [M] [G]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import ctypes as C
from bayesgp_b200 import _lib
from bayesgp_b200.api import compute_post_fun_IWP
rng = np.random.default_rng(1)
M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000
G = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20000
knots = np.linspace(0, 1, 300)
coef = 0.05 * rng.standard_normal((299, M)); glob = rng.standard_normal((2, M)); icpt = rng.standard_normal(M)
xg = np.linspace(0, 1, G)
kw = dict(global_samps=glob, knots=knots, refined_x=xg, p=3, degree=0, intercept_samps=icpt)
out = compute_post_fun_IWP(coef, **kw)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); out = compute_post_fun_IWP(coef, **kw); ts.append(time.perf_counter() - t0)
tg, tsel, tt = C.c_double(), C.c_double(), C.c_double()
_lib.load().bgp_predict_last_timing(C.byref(tg), C.byref(tsel), C.byref(tt))
print("M %d G %d best ms %.2f device gemm %.2f select %.2f total %.2f | sum lo %.12e hi %.12e mean %.12e" % (
    M, G, min(ts) * 1e3, tg.value, tsel.value, tt.value, out["plower"].sum(), out["pupper"].sum(), out["mean"].sum()))
