"""Multi-GPU check (run under torchrun on N >= 2 GPUs of one box; tests/test_gpu_multi.py launches it):
   (1) observation-sharded Laplace evaluation / gradient / AGHQ fit vs the single-GPU result;
   (2) node group: bgp_aghq_fit with the quadrature nodes split over the ranks, bgp_sample with the per-node sample
       blocks split, bgp_fit_predict_* with the grid rows split — each against the same call on one GPU;
   (3) both at once (2-D layout) when N >= 4: N/2 observation shards x 2 node-group ranks."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import torch
import torch.distributed as dist
import bayesgp_b200 as bg
from bayesgp_b200 import api
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
ok = True


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(b))))


def build(lo, hi, shard=None, node_group=None):
    ff = LaplaceObjective(y=y[lo:hi], family="Poisson", device=local)
    ff.add_iwp(x[lo:hi], x0, knots, 3)
    ff.add_fixed(np.ones(hi - lo))
    if shard:
        ff.set_shard(*shard)
    if node_group:
        ff.set_node_group(*node_group)
    return ff.finalize()


# ---- (1) observation shards ---------------------------------------------------------------------------------------
lo, hi = shard_bounds(n, rank, world)
ffs = build(lo, hi, shard=(rank, world, broadcast_unique_id(nccl_unique_id, rank)))
thetas = [np.array([-10.0]), np.array([-12.5]), np.array([-7.0])]
t0 = time.time()
vals = [ffs._eval(th, want_grad=True, want_hess=True) for th in thetas]
dt_s = time.time() - t0
fit_s = bg.marginal_laplace_tmb(ffs, 5, np.zeros(1))
ff1 = build(0, n)                       # every rank: the single-GPU reference on its own device
t0 = time.time()
ref = [ff1._eval(th, want_grad=True, want_hess=True) for th in thetas]
dt_1 = time.time() - t0
fit_1 = bg.marginal_laplace_tmb(ff1, 5, np.zeros(1))
for (v, g, w, H), (v1, g1, w1, H1) in zip(vals, ref):
    e = [abs(v - v1) / abs(v1), np.max(np.abs(g - g1)) / max(1, np.max(np.abs(g1))), rel(w, w1), rel(H, H1)]
    say("value/grad/mode/H rel err obs-sharded(%d) vs single: %.2e %.2e %.2e %.2e" % (world, *e))
    ok &= e[0] < 1e-10 and e[1] < 1e-7 and e[2] < 1e-8 and e[3] < 1e-9
m1, ms = fit_1.optresults["mode"], fit_s.optresults["mode"]
say("aghq mode single %.10f obs-sharded %.10f | lognormconst %.8f vs %.8f" % (m1[0], ms[0], fit_1.lognormconst, fit_s.lognormconst))
ok &= abs(m1[0] - ms[0]) < 1e-6 * max(1, abs(m1[0])) and abs(fit_1.lognormconst - fit_s.lognormconst) < 1e-8 * abs(fit_1.lognormconst)
say("3 evals + grads: obs-sharded %.3f s, single %.3f s" % (dt_s, dt_1))
fit_s.close()
ffs.close()

# ---- (2) node group on replicas -----------------------------------------------------------------------------------
ffn = build(0, n, node_group=(rank, world, broadcast_unique_id(nccl_unique_id, rank)))
opt = {"mode": fit_1.optresults["mode"], "hessian": fit_1.optresults["hessian"]}
K = 15
for f in (ffn, ff1):
    f.set_start(None)
t0 = time.time()
mod_n = bg.marginal_laplace_tmb(ffn, K, None, optresults=opt)
t_n = time.time() - t0
t0 = time.time()
mod_1 = bg.marginal_laplace_tmb(ff1, K, None, optresults=opt)
t_1 = time.time() - t0
e_lnc = abs(mod_n.lognormconst - mod_1.lognormconst) / abs(mod_1.lognormconst)
e_lp = rel(mod_n.normalized_posterior["nodesandweights"]["logpost"], mod_1.normalized_posterior["nodesandweights"]["logpost"])
mh_n, mh_1 = mod_n.modesandhessians, mod_1.modesandhessians
e_m, e_H = rel(mh_n["mode"], mh_1["mode"]), rel(mh_n["H"], mh_1["H"])
say("node-sharded(%d) grid vs single: lognormconst %.2e logpost %.2e modes %.2e Hessians %.2e | %.3f s vs %.3f s; owners %s"
    % (world, e_lnc, e_lp, e_m, e_H, t_n, t_1, mod_n.node_owner.tolist()))
ok &= e_lnc <= 1e-12 and e_lp <= 1e-10 and e_m <= 1e-8 and e_H <= 1e-6
# sampling: same (Z, node ids) on every rank; each rank draws the blocks of its own nodes
rng = np.random.default_rng(7)
M = 2000
Z = rng.standard_normal((ffn.p, M))
lam = mod_1.normalized_posterior["nodesandweights"]["weights"] * np.exp(mod_1.normalized_posterior["nodesandweights"]["logpost_normalized"])
idx = rng.choice(K, size=M, p=lam / lam.sum()).astype(np.int32)
s_n = api.sample_marginal(mod_n, M, Z, idx)
s_1 = api.sample_marginal(mod_1, M, Z, idx)
e_s = rel(s_n["samps"], s_1["samps"])
say("node-sharded sampling vs single: %.2e (nodes hit: %d)" % (e_s, len(np.unique(idx))))
ok &= e_s <= 1e-6
# predict from the resident samples, grid rows split over the ranks
from bayesgp_b200.terms import prepare_term
term = prepare_term(bg.Term("IWP", "x", x, order=3, knots=knots, initial_location=x0))
p = ffn.p
res_n = api.FitResult([term], mod_n, ffn, {"x": np.arange(p - 3, p - 1)}, {"x": np.arange(0, p - 3)}, {"intercept": p - 1}, "Poisson", s_n)
res_1 = api.FitResult([term], mod_1, ff1, {"x": np.arange(p - 3, p - 1)}, {"x": np.arange(0, p - 3)}, {"intercept": p - 1}, "Poisson", s_1)
xg = np.linspace(x.min(), x.max(), 1001)
for degree in (0, 1):
    pr_n = api.predict(res_n, newdata=xg, variable="x", degree=degree)
    pr_1 = api.predict(res_1, newdata=xg, variable="x", degree=degree)
    host = dict(s_1)
    host.pop("resident")                # the host-buffer path (bgp_predict_iwp) on the same samples
    res_h = api.FitResult([term], mod_1, ff1, res_1.boundary_samp_indexes, res_1.random_samp_indexes, res_1.fixed_samp_indexes, "Poisson", host)
    pr_h = api.predict(res_h, newdata=xg, variable="x", degree=degree)
    e = [rel(pr_n[k_], pr_1[k_]) for k_ in ("mean", "plower", "pupper")]
    eh = [rel(pr_1[k_], pr_h[k_]) for k_ in ("mean", "plower", "pupper")]
    say("predict degree %d: row-sharded vs single %.2e %.2e %.2e | resident vs host-buffer path %.2e %.2e %.2e" % (degree, *e, *eh))
    ok &= max(e) <= 1e-6 and max(eh) <= 1e-12
mod_n.close()
ffn.close()

# ---- (3) 2-D layout -----------------------------------------------------------------------------------------------
if world >= 4 and world % 2 == 0:
    nshard = world // 2
    srank, nrank = rank // 2, rank % 2                  # ranks (2s, 2s+1) hold observation shard s
    # one communicator per observation-shard index is not needed: ranks with equal nrank form an observation group,
    # ranks with equal srank form a node group; ids are created by the lowest rank of each group and shared
    obs_groups = [dist.new_group([2 * s + r for s in range(nshard)]) for r in range(2)]
    node_groups = [dist.new_group([2 * s, 2 * s + 1]) for s in range(nshard)]
    def bcast_id(group, root):
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == root:
            buf.copy_(torch.tensor(list(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, src=root, group=group)
        return bytes(buf.cpu().tolist())
    oid = bcast_id(obs_groups[nrank], nrank)
    nid = bcast_id(node_groups[srank], 2 * srank)
    lo, hi = shard_bounds(n, srank, nshard)
    ff2 = build(lo, hi, shard=(srank, nshard, oid), node_group=(nrank, 2, nid))
    ff2.set_start(None)
    mod_2 = bg.marginal_laplace_tmb(ff2, K, None, optresults=opt)
    e2 = abs(mod_2.lognormconst - mod_1.lognormconst) / abs(mod_1.lognormconst)
    say("2-D layout (%d observation shards x 2 node ranks) vs single: lognormconst %.2e" % (nshard, e2))
    ok &= e2 <= 1e-10
    mod_2.close()
    ff2.close()

flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
say("MGPU_CHECK", "PASS" if float(flag[0]) == 1.0 else "FAIL")
mod_1.close()
fit_1.close()
ff1.close()
dist.barrier()
dist.destroy_process_group()
