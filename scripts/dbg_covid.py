"""Diagnostics: per-fixture GPU-vs-oracle differences of the grid quantities (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
import test_gpu_fit as T
import bayesgp_b200 as bg
from oracle import fit as ofit
from oracle.aghq import marginal_laplace_tmb as o_mlt
from oracle.laplace import LaplaceObjective as OFF
from helpers import relerr
for name in ["covid", "sim1_gaussian", "binomial_sgp"]:
    oargs, pargs, k = T._both(name)
    model = ofit.build_model(oargs["y"], oargs["terms"], oargs["fixed"], oargs["family"], oargs.get("size"))[0]
    off = OFF(model)
    omod = o_mlt(off, k, np.zeros(model.S))
    pfit = bg.model_fit(pargs["y"], pargs["terms"], pargs["fixed"], family=pargs["family"], size=pargs.get("size"),
                        aghq_k=k, M=0, optresults={"mode": omod.mode, "hessian": omod.hessian})
    mod = pfit.mod
    nw = mod.normalized_posterior["nodesandweights"]
    print(name, "lognormconst diff", mod.lognormconst - omod.lognormconst, "rel", abs(mod.lognormconst - omod.lognormconst) / abs(omod.lognormconst))
    print("  logpost diff", nw["logpost"] - omod.logpost)
    mh = mod.modesandhessians
    print("  mode relerr", [relerr(mh["mode"][j], omod.modes[j]) for j in range(mod.K)][:6])
    print("  H relerr", [relerr(mh["H"][j], omod.hessians[j]) for j in range(mod.K)][:6])
    for j in range(mod.S):
        print("  marg", j, np.max(np.abs(mod.marginals[j]["logmargpost"] - omod.marginals[j]["logmargpost"])))
    # conditioning of H at node 0
    H = omod.hessians[0]
    print("  cond(H0) %.3e" % np.linalg.cond(H))
    pfit.close()
