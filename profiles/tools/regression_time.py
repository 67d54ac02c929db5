import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g.load_package(); nv = pkg.native
if os.environ.get("FMB_AB_LIB"):
    nv.LIB_PATH = os.environ["FMB_AB_LIB"]          # A/B builds of the library (timing tool only)
nv.init(0)
RV = pkg.RandomVariableCuda
rng = np.random.default_rng(1)
for n in ([int(a) for a in sys.argv[1:]] or [2000, 1_000_000]):
    z1, z2, z3 = rng.standard_normal((3, n))
    d = 1 / (1 + 0.05 * np.exp(0.2 * z1 - 0.02) * 0.5)
    D = 1 / (1 + 0.05 * np.exp(0.15 * (0.7 * z1 + 0.7 * z2) - 0.01) * 10)
    N = np.exp(0.05 * 5 + 0.02 * z3)
    basis = [RV(1.0, np.ones(n)).mult(1.0), RV(1.0, d), RV(1.0, d * d), RV(1.0, D), RV(1.0, D * D), RV(1.0, 1 / N)]
    basis[0] = pkg.RandomVariableFromDoubleArray(1.0)
    y = RV(1.0, rng.standard_normal(n) + d)
    for b in basis[1:]:
        b.dv.h
    y.dv.h
    def fit(reps):
        nv.synchronize(); nv.timer_start()
        for _ in range(reps):
            est = pkg.MonteCarloConditionalExpectationRegression(basis)
            est.getConditionalExpectation(y).dv.h
        return nv.timer_stop_ms() / reps * 1e3
    fit(3)
    print("n=%d: fit+predict %.1f us per regression (device time, back to back)" % (n, fit(50)))
