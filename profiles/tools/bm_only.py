"""Brownian generation only, C4 shape (40 steps x 3 factors) or C3 shape (1000 x 2): the target of the source-level ncu captures.

    python profiles/tools/bm_only.py [paths] [steps] [factors] [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nv = pkg.native
nv.init(0)
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
factors = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
td = pkg.TimeDiscretizationFromArray(0.0, steps, 0.5)
best = 1e9
for r in range(reps):
    bm = pkg.BrownianMotionCuda(td, factors, paths, 3141 + r)
    nv.synchronize()
    nv.timer_start()
    bm.getBrownianIncrement(0, 0)
    ms = nv.timer_stop_ms()
    best = min(best, ms)
    del bm
print("bm %d x %d x %d: best %.3f ms (incl. jump-ahead), %.1f G increments/s" % (paths, steps, factors, best, paths * steps * factors / best / 1e6))
