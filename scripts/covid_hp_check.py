"""GPU Laplace values on the README model vs the 40-digit reference (tests/golden/covid_hp.json)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
from helpers import GOLDEN, covid_model, tmbdata_from_oracle
from bayesgp_b200 import make_objective
hp = json.load(open(os.path.join(GOLDEN, "covid_hp.json")))
ff = make_objective(tmbdata_from_oracle(covid_model()[0]))
for reuse in (True, False):
    ff.set_factor_reuse(reuse)
    ff.set_start(None)
    errs = []
    for t, v, o in zip(hp["theta"], hp["value"], hp["oracle_fp64_value"]):
        got = ff.fn(np.array([t]))
        errs.append((t, got - v, o - v))
    print("reuse", reuse, " theta: gpu - truth | fp64 oracle - truth")
    for t, e, eo in errs:
        print("  %9.5f  %+.3e (%.1e rel)  | %+.3e" % (t, e, abs(e) / abs(v), eo))
ff.close()
