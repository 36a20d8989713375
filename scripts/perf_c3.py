"""Quick C3-scale timing: n = 1e6 Poisson, IWP3 k = 300 (p = 302), 15 theta nodes."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
from bayesgp_b200.objective import LaplaceObjective

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rng = np.random.default_rng(20243)
x = rng.uniform(0, 1, n)
eta = 1.0 + np.sin(2 * np.pi * x) + 0.5 * np.cos(6 * np.pi * x)
y = rng.poisson(np.exp(eta)).astype(np.float64)
x0 = x.min(); xi = x - x0
knots = np.linspace(xi.min(), xi.max(), k)
t0 = time.time()
ff = LaplaceObjective(y=y, family="Poisson")
ff.add_iwp(x, x0, knots, 3)
ff.add_fixed(np.ones(n))
ff.finalize()
print("model build %.3f s  n=%d p=%d" % (time.time() - t0, ff.n, ff.p))
thetas = np.linspace(4.0, 9.0, 15)[:, None]
for rep in range(3):
    t0 = time.time()
    if rep == 0:
        ff.set_start(None)
    vals, modes, Hs, iters = ff.fn_batch(thetas, want_modes=True, want_hess=False)
    dt = time.time() - t0
    tm = ff.last_timing()
    print("rep %d: 15 evals %.3f s (%.1f evals/s), newton iters %d, device total %.1f ms" % (rep, dt, 15 / dt, iters, tm["total_ms"]))
    print("   cumulative: lik %.1f ms / %d launches (%.3f ms each); hess %.1f ms / %d (%.3f ms each); chol %.1f ms / %d (%.3f each)" % (
        tm["lik_ms"], tm["lik_launches"], tm["lik_ms"] / max(1, tm["lik_launches"]),
        tm["hess_ms"], tm["hess_launches"], tm["hess_ms"] / max(1, tm["hess_launches"]),
        tm["chol_ms"], tm["chol_launches"], tm["chol_ms"] / max(1, tm["chol_launches"])))
print(vals)
p = ff.p
flops = n * p * (p + 1)
hf = ff.hessian_flops()
print("hessian flops dense %.4e structural %.4e (%.1f%%) -> executed %.2f TFLOP/s" % (hf["dense"], hf["structural"], 100 * hf["structural"] / hf["dense"], hf["structural"] / (tm["hess_ms"] / tm["hess_launches"] * 1e-3) / 1e12))
print("SYRK algorithmic flops %.3e -> %.2f TFLOP/s ; lik bytes %.3e -> %.1f GB/s" % (
    flops, flops / (tm["hess_ms"] / tm["hess_launches"] * 1e-3) / 1e12, 8.0 * n * (ff.p + 2),
    8.0 * n * (round((p + 15) // 16 * 16) + 2) / (tm["lik_ms"] / tm["lik_launches"] * 1e-3) / 1e9))
