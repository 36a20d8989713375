"""Per-run cost of the node-sharded C3 grid, measured on ONE GPU: every contiguous run of nodes is evaluated from the
optimiser's end state (mode + tangent at the centre) as its owner rank would, for several partitions of the 15 nodes."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bench
from bayesgp_b200.workloads import c3_data, iwp_knots

x, y = c3_data(1_000_000)
x0, knots = iwp_knots(x, bench.P_KNOTS)
ff = bench.build_b200(x, y, 0, x0, knots)
mode, sd, thetas, w_mode, t_mode = bench.node_grid(ff)
z = (thetas[:, 0] - mode) / sd
print("z:", np.round(z, 2).tolist())
ff.fn_batch(thetas, want_modes=False)       # warm-up
single = {}
for j in range(15):                         # cost of every node alone from the centre state
    ff.set_start_at(np.array([mode]), w_mode, t_mode)
    it0 = ff.newton_iters
    t0 = time.perf_counter()
    ff.fn_batch(thetas[j:j + 1], want_modes=False)
    single[j] = ((time.perf_counter() - t0) * 1e3, ff.newton_iters - it0)
print("alone (ms, iters):", {j: (round(a, 2), b) for j, (a, b) in single.items()})
for name, sizes in (("count", [2, 2, 2, 2, 2, 2, 2, 1]), ("cost-v1", [1, 1, 2, 3, 3, 3, 1, 1]), ("alt", [1, 2, 2, 3, 2, 2, 2, 1]),
                    ("n4-count", [4, 4, 4, 3]), ("n4-alt", [3, 4, 5, 3])):
    lo = 0
    rows = []
    for s in sizes:
        ff.set_start_at(np.array([mode]), w_mode, t_mode)
        it0 = ff.newton_iters
        t0 = time.perf_counter()
        ff.fn_batch(thetas[lo:lo + s], want_modes=False)
        rows.append((round((time.perf_counter() - t0) * 1e3, 2), ff.newton_iters - it0))
        lo += s
    print(name, sizes, "per run (ms, iters):", rows, "max %.2f ms" % max(r[0] for r in rows))
ff.close()
