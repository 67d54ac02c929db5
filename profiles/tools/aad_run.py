"""Black-Scholes greeks by AAD on device vectors (1 M paths, 10 steps): forward recording, backward sweep, averages - wall time and launches
of each part (FMB_LAZY_MIN_N=0: deferred chains for every size).

    python profiles/tools/aad_run.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
pkg = g.load_package(); nv = pkg.native; nv.init(0)
factory = pkg.RandomVariableCudaFactory()
td = pkg.TimeDiscretizationFromArray(0.0, 10, 0.5)
f = pkg.RandomVariableDifferentiableAADFactory(factory)
bmA = pkg.BrownianMotionCuda(td, 1, 1_000_000, 3141, factory)
bmA.getBrownianIncrement(0, 0)
def greeks():
    model = pkg.BlackScholesModel(1.0, 0.05, 0.30, f)
    mc = pkg.MonteCarloAssetModel(model, pkg.EulerSchemeFromProcessModel(model, bmA))
    nv.synchronize(); t0 = time.perf_counter(); l0 = nv.launch_count()
    value = pkg.EuropeanOption(5.0, 1.05).getValueRV(0.0, mc)
    price = value.getAverage()
    t1 = time.perf_counter(); l1 = nv.launch_count()
    g_ = value.getGradient()
    t2 = time.perf_counter(); l2 = nv.launch_count()
    out = [g_[i].getAverage() for i in (model.getInitialValue()[0].getID(), model.getRiskFreeRate().getID(), model.getVolatility().getID())]
    t3 = time.perf_counter(); l3 = nv.launch_count()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, l1 - l0, l2 - l1, l3 - l2, out
for i in range(3):
    print(greeks())
