"""Multi-GPU check (run under torchrun on N GPUs of one box):
   observation-sharded Laplace evaluation / gradient / AGHQ fit vs the single-GPU result."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import torch
import torch.distributed as dist
import bayesgp_b200 as bg
from bayesgp_b200.distributed import broadcast_unique_id, nccl_unique_id, shard_bounds
from bayesgp_b200.objective import LaplaceObjective
from bayesgp_b200.workloads import c3_data, iwp_knots

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
x, y = c3_data(n)
x0, knots = iwp_knots(x, k)

def build(lo, hi, shard):
    ff = LaplaceObjective(y=y[lo:hi], family="Poisson", device=local)
    ff.add_iwp(x[lo:hi], x0, knots, 3)
    ff.add_fixed(np.ones(hi - lo))
    if shard:
        uid = broadcast_unique_id(nccl_unique_id, rank)
        ff.set_shard(rank, world, uid)
    return ff.finalize()

lo, hi = shard_bounds(n, rank, world)
ffs = build(lo, hi, True)
thetas = [np.array([-10.0]), np.array([-12.5]), np.array([-7.0])]
t0 = time.time()
vals = [ffs._eval(th, want_grad=True, want_hess=True) for th in thetas]
dt_s = time.time() - t0
fit_s = bg.marginal_laplace_tmb(ffs, 5, np.zeros(1))
ok = True
if rank == 0:
    ff1 = build(0, n, False)
    t0 = time.time()
    ref = [ff1._eval(th, want_grad=True, want_hess=True) for th in thetas]
    dt_1 = time.time() - t0
    fit_1 = bg.marginal_laplace_tmb(ff1, 5, np.zeros(1))
    for (v, g, w, H), (v1, g1, w1, H1) in zip(vals, ref):
        e = [abs(v - v1) / abs(v1), np.max(np.abs(g - g1)) / max(1, np.max(np.abs(g1))),
             np.max(np.abs(w - w1)) / np.max(np.abs(w1)), np.max(np.abs(H - H1)) / np.max(np.abs(H1))]
        print("value/grad/mode/H rel err sharded(%d) vs single: %.2e %.2e %.2e %.2e" % (world, *e))
        ok &= e[0] < 1e-10 and e[1] < 1e-7 and e[2] < 1e-8 and e[3] < 1e-9
    m1, ms = fit_1.optresults["mode"], fit_s.optresults["mode"]
    print("aghq mode single %.10f sharded %.10f | lognormconst %.8f vs %.8f" % (m1[0], ms[0], fit_1.lognormconst, fit_s.lognormconst))
    ok &= abs(m1[0] - ms[0]) < 1e-6 * max(1, abs(m1[0])) and abs(fit_1.lognormconst - fit_s.lognormconst) < 1e-8 * abs(fit_1.lognormconst)
    print("3 evals + grads: sharded %.3f s, single %.3f s" % (dt_s, dt_1))
    print("MGPU_CHECK", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
