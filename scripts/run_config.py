"""End-to-end model_fit() + predict() on BASELINE.json's large synthetic configs:
   c4: Binomial n = 1e6, IWP2 k=440 + sGP (a = 2 pi 5, k = 20, m = 1) + intercept (p = 497), 2-D AGHQ 7^2
       (SURVEY 8d proposed IWP2 k=200 + sGP k=100 for the same p; the sGP precision of Compute_Q_sB is
       numerically singular beyond k ~ 30 at this frequency — cond(H) 2e13 at k=40, not positive definite in
       FP64 at k>=60, measured — so the columns are moved to the IWP term)
   c5: Poisson  n = 1e7, three IWP3 k = 334 terms + intercept (p = 1006), 3-D AGHQ 5^3, M = 1e5, G = 1e5
   Under torchrun:  c4 — every rank holds all rows and the ranks form ONE NODE GROUP ("7^2 nodes sharded across 8
   GPUs": grid nodes, sample blocks and prediction rows are split inside the library, BFGS / Richardson replicated);
   c5 — 2-D layout: `BGP_NODE_RANKS` (default 2) node-group ranks x world / that many observation shards (NCCL
   all-reduce of g and the packed H per Newton iteration inside each observation group).
   usage: run_config.py c4|c5 [n] [aghq_k] [M] [G]"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import bayesgp_b200 as bg
from bayesgp_b200 import api
from bayesgp_b200.distributed import broadcast_unique_id, nccl_unique_id, shard_bounds

cfg = sys.argv[1]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
arg = lambda i, d: type(d)(float(sys.argv[i])) if len(sys.argv) > i else d
t_all = time.time()
if cfg == "c4":
    n, k, M, G = arg(2, 1_000_000), arg(3, 7), arg(4, 10_000), arg(5, 1000)
    rng = np.random.default_rng(20244)
    x1, x2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    eta = -0.3 + np.sin(2 * np.pi * x1) + 0.6 * np.cos(2 * np.pi * 5 * x2)
    size = 1.0 + rng.poisson(9, n)
    y = rng.binomial(size.astype(int), 1 / (1 + np.exp(-eta))).astype(np.float64)
    cols = {"x1": x1, "x2": x2}
    mk = lambda sl: [bg.Term("IWP", "x1", x1[sl], order=2, knots=np.linspace(0, x1.max() - x1.min(), 440), initial_location=float(x1.min())),
                     bg.Term("sGP", "x2", x2[sl], a=2 * np.pi * 5, k=20, m=1, region=np.array([0.0, 1.0]), accuracy=0.01)]
    family = "Binomial"
else:
    n, k, M, G = arg(2, 10_000_000), arg(3, 5), arg(4, 100_000), arg(5, 100_000)
    rng = np.random.default_rng(20245)
    xs = [rng.uniform(0, 1, n) for _ in range(3)]
    # none of the three effects lies in the null space of its IWP3 penalty (a quadratic would send theta -> +inf)
    eta = 0.5 + np.sin(2 * np.pi * xs[0]) + 0.4 * np.sin(3 * np.pi * xs[1]) + 0.6 * np.sin(2.5 * np.pi * xs[2] + 1.0)
    y = rng.poisson(np.exp(eta)).astype(np.float64)
    size = None
    cols = {"x%d" % (i + 1): xs[i] for i in range(3)}
    mk = lambda sl: [bg.Term("IWP", "x%d" % (i + 1), xs[i][sl], order=3, knots=np.linspace(0, xs[i].max() - xs[i].min(), 334),
                             initial_location=float(xs[i].min())) for i in range(3)]
    family = "Poisson"
# layout: node-group ranks (same rows) x observation shards
nnode = 1
if world > 1:
    nnode = world if cfg == "c4" else int(os.environ.get("BGP_NODE_RANKS", "2"))
    nnode = max(1, min(world, nnode))
    while world % nnode:
        nnode -= 1
nshard = world // nnode
srank, nrank = rank // nnode, rank % nnode            # ranks (s * nnode + r) hold observation shard s


def group_id(members, root):
    """ncclUniqueId made by `root`, shared with `members` (a torch.distributed subgroup carries the 128 bytes)."""
    import torch
    g = dist.new_group(members)
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == root:
        buf.copy_(torch.tensor(list(nccl_unique_id()), dtype=torch.uint8))
    if rank in members:
        dist.broadcast(buf, src=root, group=g)
    return bytes(buf.cpu().tolist())


shard = node_group = None
if world > 1:
    # every rank takes part in every new_group call (torch.distributed requirement)
    ids_obs = [group_id([s * nnode + r for s in range(nshard)], r) for r in range(nnode)]
    ids_node = [group_id([s * nnode + r for r in range(nnode)], s * nnode) for s in range(nshard)]
    if nshard > 1:
        shard = (srank, nshard, ids_obs[nrank])
    if nnode > 1:
        node_group = (nrank, nnode, ids_node[srank])
lo, hi = shard_bounds(n, srank, nshard)
sl = slice(lo, hi)
t0 = time.time()
ff, terms, rand_idx, bnd_idx, fix_idx = api.build_objective(y[sl], mk(sl), {}, family, None if size is None else size[sl],
                                                            device=local, shard=shard, node_group=node_group)
if cfg == "c5":
    ff.set_hessian_retry(True)        # n = 1e7: numDeriv's default step can land on rounding noise
t_build = time.time() - t0
hf = ff.hessian_flops()
t0 = time.time()
mod = api.marginal_laplace_tmb(ff, k, np.zeros(ff.S))
t_fit = time.time() - t0
out = {"config": cfg, "n": n, "p": ff.p, "S": ff.S, "K": mod.K, "gpus": world, "observation_shards": nshard,
       "node_group_ranks": nnode, "build_s": t_build, "fit_s": t_fit, "opt_s": mod.diagnostics["opt_ms"] * 1e-3,
       "grid_s": mod.diagnostics["grid_ms"] * 1e-3, "hessian_fallback": mod.diagnostics["hessian_fallback"],
       "data_s": t0 - t_all - t_build, "hessian_structural_fraction": hf["structural"] / hf["dense"],
       "theta_mode": mod.optresults["mode"].tolist(), "convergence": mod.optresults["convergence"],
       "fn_count": mod.optresults["fn_count"], "gr_count": mod.optresults["gr_count"], "lognormconst": mod.lognormconst,
       "laplace_evals": ff.counters()["laplace_evals"], "newton_iters": ff.counters()["newton_iters"],
       "factor_reuses": ff.counters()["factor_reuses"]}
# sampling and predict are collective over the node group (every rank draws the sample blocks of its nodes and
# summarises its block of prediction rows); with observation shards every shard group repeats them (replicated)
t0 = time.time()
samps = api.sample_marginal(mod, M, seed=1)
out["sample_s"] = time.time() - t0
res = api.FitResult(terms, mod, ff, bnd_idx, rand_idx, fix_idx, family)
res.samps = samps
out["predict_s"] = {}
for t in terms:
    xg = np.linspace(cols[t.name].min(), cols[t.name].max(), G)
    t0 = time.time()
    pr = api.predict(res, newdata={t.name: xg}, variable=t.name, degree=0)
    out["predict_s"][t.name] = time.time() - t0
    out.setdefault("predict_mean_range", {})[t.name] = [float(np.min(pr["mean"])), float(np.max(pr["mean"]))]
out["total_s"] = time.time() - t_all
# per-kernel rooflines of this rank's share (CUDA events around the phases, cumulative over the whole fit; with
# observation shards the Hessian phase includes the all-reduce of the packed triangle)
tm, lb = ff.last_timing(), ff.lik_bytes()
nh, nl, nc = max(1, tm["hess_launches"]), max(1, tm["lik_launches"]), max(1, tm["chol_launches"])
out["kernels"] = {"hess_ms_per_launch": tm["hess_ms"] / nh, "hess_launches": tm["hess_launches"],
                  "hess_executed_tflops": hf["structural"] / (tm["hess_ms"] / nh * 1e-3) / 1e12,
                  "hess_dense_equivalent_tflops": hf["dense"] / (tm["hess_ms"] / nh * 1e-3) / 1e12,
                  "lik_ms_per_launch": tm["lik_ms"] / nl, "lik_launches": tm["lik_launches"],
                  "lik_gbs_on_moved_bytes": lb["structural"] / (tm["lik_ms"] / nl * 1e-3) / 1e9,
                  "chol_ms_per_launch": tm["chol_ms"] / nc, "chol_launches": tm["chol_launches"],
                  "rows_on_this_rank": int(hi - lo)}
out["node_owner_counts"] = np.bincount(mod.node_owner, minlength=nnode).tolist()
if rank == 0:
    print("RUN_CONFIG " + json.dumps(out), flush=True)
if dist:
    dist.barrier()
    dist.destroy_process_group()
