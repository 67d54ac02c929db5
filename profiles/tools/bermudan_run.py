"""One C5 Bermudan valuation (1 M paths by default) a few times: wall time split into simulation and backward induction, and - under
`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...` - the launch list of the last repetition.

    python profiles/tools/bermudan_run.py [paths] [reps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nv = pkg.native
nv.init(0)
from common import lmm_setup, bermudan_spec  # noqa: E402

paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
s = lmm_setup(pkg)
b = bermudan_spec(s)
factory = pkg.RandomVariableCudaFactory()
model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], factory, s["cov"], None, {"measure": "SPOT", "stateSpace": "LOGNORMAL"})
product = pkg.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
for rep in range(reps):
    nv.synchronize()
    l0 = nv.launch_count()
    t0 = time.perf_counter()
    bm = pkg.BrownianMotionCuda(s["sim"], s["F"], paths, 3141, factory)
    sim = pkg.LIBORMonteCarloSimulationFromLIBORModel(pkg.EulerSchemeFromProcessModel(model, bm, 2))
    sim.getProcess().getProcessValue(s["T"], s["N"] - 1)
    t1 = time.perf_counter()
    nv.synchronize()
    t2 = time.perf_counter()
    price = product.getValue(sim)
    nv.synchronize()
    t3 = time.perf_counter()
    print("rep %d: simulate (host enqueue %.2f ms, device done %.2f ms), induction %.2f ms, total %.2f ms, launches %d, price %.12f"
          % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t0), 1e3 * (t3 - t2), 1e3 * (t3 - t0), nv.launch_count() - l0, price), flush=True)
    del sim, bm
