"""Fixed launch sequence for ncu: the predict leg of bench.py (G = 1e5, M = 1e4, IWP3 k = 300), two calls."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import bench
t0 = time.time()
r = bench.predict_leg(0, 35.4, G=int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000)
print(r, "total", time.time() - t0)
