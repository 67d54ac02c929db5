"""Host-side profile of one C5 Bermudan valuation (1 M paths): where do the ~20 ms go?  cProfile over the Python mirror."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from common import lmm_setup, lmm_device, bermudan_spec  # noqa: E402

pkg = graft.load_package()
nv = pkg.native
nv.init(0)
s = lmm_setup(pkg)
b = bermudan_spec(s)
product = pkg.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])


def run(seed):
    sim = lmm_device(pkg, s, 1_000_000, seed=seed)
    t0 = time.perf_counter()
    sim.getProcess().getProcessValue(40, 39)
    nv.synchronize()
    t1 = time.perf_counter()
    v = product.getValue(sim)
    nv.synchronize()
    t2 = time.perf_counter()
    return v, (t1 - t0) * 1e3, (t2 - t1) * 1e3


for k in range(3):
    print("price %.10f  simulate %.2f ms  product %.2f ms  launches so far %d" % (run(3141 + k) + (nv.launch_count(),)))
l0 = nv.launch_count()
pr = cProfile.Profile()
pr.enable()
run(4000)
pr.disable()
print("launches in one valuation:", nv.launch_count() - l0)
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
