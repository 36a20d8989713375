"""Fixed launch sequence for ncu: the predict leg of bench.py (G = 1e5, M = 1e4, IWP3 k = 300); the profiler range
covers the first strips of one call."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import bench
from bayesgp_b200 import _lib
G = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
r = bench.predict_leg(0, 35.4, G=G)
print(r)
lib = _lib.load()
lib.bgp_profiler_range(1)
r = bench.predict_leg(0, 35.4, G=4352, reps=1)        # four strips of 1088 rows, same M and design
lib.bgp_profiler_range(0)
print(r)
