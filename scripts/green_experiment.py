"""Experiment: two models of the same C3 data driven from two host threads, each evaluating half of the 15-node grid,
with and without the SM partition (BGP_GREEN_SMS) — does the Cholesky of one chain overlap the big kernels of the other?
usage: [BGP_GREEN_SMS=8|16] python scripts/green_experiment.py"""
import os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bench
from bayesgp_b200.workloads import c3_data, iwp_knots

x, y = c3_data(1_000_000)
x0, knots = iwp_knots(x, bench.P_KNOTS)
ffs = [bench.build_b200(x, y, 0, x0, knots) for _ in range(2)]
mode, sd, thetas, w_mode, t_mode = bench.node_grid(ffs[0])
print("green", os.environ.get("BGP_GREEN_SMS"), "lik blocks / SM count in use:", ffs[0].last_timing())
left = thetas[:8][::-1].copy()       # centre first, then outwards
right = thetas[8:].copy()


def run_single(ff, reps=6):
    best = 1e9
    for _ in range(reps):
        ff.set_start_at(np.array([mode]), w_mode, t_mode)
        t0 = time.perf_counter()
        ff.fn_batch(thetas, want_modes=False)
        best = min(best, time.perf_counter() - t0)
    return best


def run_pair(reps=6):
    best = 1e9
    for _ in range(reps):
        for ff in ffs:
            ff.set_start_at(np.array([mode]), w_mode, t_mode)
        bar = threading.Barrier(3)
        out = [None, None]

        def work(i, th):
            bar.wait()
            out[i] = ffs[i].fn_batch(th, want_modes=False)[0]

        ts = [threading.Thread(target=work, args=(0, left)), threading.Thread(target=work, args=(1, right))]
        for t in ts:
            t.start()
        bar.wait()
        t0 = time.perf_counter()
        for t in ts:
            t.join()
        best = min(best, time.perf_counter() - t0)
    return best, out


t1 = run_single(ffs[0])
t2, out = run_pair()
vals = ffs[0].fn_batch(thetas, want_modes=False)[0]
err = max(np.max(np.abs(out[0] - vals[:8][::-1]) / np.abs(vals[:8])), np.max(np.abs(out[1] - vals[8:]) / np.abs(vals[8:])))
print("single model, 15 nodes: %.2f ms (%.1f evals/s)" % (t1 * 1e3, 15 / t1))
print("two models on two threads, 8 + 7 nodes: %.2f ms (%.1f evals/s), max rel diff vs single %.2e" % (t2 * 1e3, 15 / t2, err))
for ff in ffs:
    ff.close()
