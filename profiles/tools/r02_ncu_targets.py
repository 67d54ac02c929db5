"""One launch of each round-2 kernel at a representative size, for `ncu --set full` (profiles/r02_*.csv are made from its report):

    ncu --set full --clock-control none -k regex:'bmGenerateKernel|eulerLmmKernel|eulerHestonKernel|eulerHullWhiteKernel|eulerTwoFactorTmaKernel|momentsKernel|selectHistogramKernel|predictKernelV' \
        -o gpurun_out/r02_kernels python profiles/tools/r02_ncu_targets.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nv = pkg.native
nv.init(0)
from common import lmm_setup, lmm_device, bermudan_spec  # noqa: E402

# C4: Brownian generation (bulk-store flush) + LMM Euler kernel, 4 M paths
s = lmm_setup(pkg)
sim = lmm_device(pkg, s, 4_000_000, scheme=2)
sim.getProcess().getProcessValue(40, 39)
nv.synchronize()
# regression on 1 M-element basis functions (C5 per-GPU size): moments with last-block finalisation + solve, prediction
b = bermudan_spec(s)
sim1 = lmm_device(pkg, s, 1_000_000, scheme=2)
product = pkg.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
basis = product.getBasisFunctions(b["fixing"][5], sim1)
y = sim1.getNumeraire(b["payment"][5]).invert()
pkg.MonteCarloConditionalExpectationRegression(basis).getConditionalExpectation(y)
q = sim.getNumeraire(5.0).getQuantile(0.99)                     # radix select over 4 M elements
nv.synchronize()
del sim, sim1, basis, y
nv.load().fmb_pool_trim()
# C3 shape: Heston 1 M x 1000, plain kernel and bulk-copy pipeline; C2 shape: Hull-White 2 M x 200, both
td = pkg.TimeDiscretizationFromArray(0.0, 1000, 0.005)
bm = pkg.BrownianMotionCuda(td, 2, 1_000_000, 31415)
model = pkg.HestonModel(1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.5, 0.1, 1, bm.randomVariableFactory)
for tma in ("0", "1"):
    os.environ["FMB_EULER_TMA"] = tma
    pkg.EulerSchemeFromProcessModel(model, bm).getProcessValue(1000, 0)
    nv.synchronize()
del bm
nv.load().fmb_pool_trim()
td = pkg.TimeDiscretizationFromArray(0.0, 200, 0.1)
vt = np.arange(0, 21.0)
vm = pkg.ShortRateVolatilityModelAsGiven(pkg.TimeDiscretizationFromArray(vt), 0.005 + 0.0005 * np.floor(vt) / 20, np.full(vt.size, 0.1))
bm = pkg.BrownianMotionCuda(td, 2, 2_000_000, 3141)
hw = pkg.HullWhiteModel(bm.randomVariableFactory, pkg.TimeDiscretizationFromArray(0.0, 40, 0.5), vm)
for tma in ("0", "1"):
    os.environ["FMB_EULER_TMA"] = tma
    pkg.EulerSchemeFromProcessModel(hw, bm, 0).getProcessValue(200, 1)
    nv.synchronize()
print("done", q)
