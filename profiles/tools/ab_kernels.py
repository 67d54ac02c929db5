"""A/B timings of round-2 kernel variants on one B200 (device events on the library's stream, best of 3):
Brownian generation with the element-wise flush vs the bulk-store (TMA) flush, Heston / Hull-White Euler with plain loads vs the
bulk-copy + mbarrier pipeline.  Environment switches FMB_BM_TMA / FMB_EULER_TMA are read per call.

    python profiles/tools/ab_kernels.py [c4] [c3] [c2]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
nv = pkg.native
nv.init(0)


def best(fn, reps=3):
    fn()
    nv.synchronize()
    ms = []
    for _ in range(reps):
        nv.timer_start()
        fn()
        ms.append(nv.timer_stop_ms())
    return min(ms)


def bm_only(T, dt, F, P):
    td = pkg.TimeDiscretizationFromArray(0.0, T, dt)

    def run():
        bm = pkg.BrownianMotionCuda(td, F, P, 3141)
        bm.getBrownianIncrement(0, 0)
    return run


def main():
    which = sys.argv[1:] or ["c4", "c3", "c2", "lmm"]
    out = []
    if "lmm" in which:
        # production LMM Euler kernel (lane per path, serial prefix sum in registers) vs the experimental lane-per-rate kernel (warp-shuffle scans)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from common import lmm_setup
        s_ = lmm_setup(pkg)
        P = 4_000_000
        factory = pkg.RandomVariableCudaFactory()
        model = pkg.LIBORMarketModelFromCovarianceModel.of(s_["tenor"], None, s_["L0"], s_["df"], factory, s_["cov"], None, {"measure": "SPOT", "stateSpace": "LOGNORMAL"})
        bm = pkg.BrownianMotionCuda(s_["sim"], 3, P, 3141, factory)
        bm.getBrownianIncrement(0, 0)
        vals = {}
        for variant in ("production", "shuffle"):
            os.environ["FMB_LMM_VARIANT"] = variant

            def run():
                pkg.EulerSchemeFromProcessModel(model, bm, 2).getProcessValue(40, 39)
            ms = best(run)
            proc = pkg.EulerSchemeFromProcessModel(model, bm, 2)
            vals[variant] = [proc.getProcessValue(t, j).getRealizations()[:200_000] for t, j in ((1, 39), (20, 21), (40, 39))]
            out.append({"what": "LMM Euler kernel C4 (4M paths, 780 live rate-steps per path)", "FMB_LMM_VARIANT": variant, "ms": ms})
            del proc
        os.environ.pop("FMB_LMM_VARIANT", None)
        dev = max(float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 0.05))) for a, b in zip(vals["production"], vals["shuffle"]))
        out.append({"what": "max relative deviation lane-per-rate vs production (scale 0.05)", "value": dev})
        del bm, vals
        nv.load().fmb_pool_trim()
    if "c4" in which:
        for tma in ("0", "1"):
            os.environ["FMB_BM_TMA"] = tma
            ms = best(bm_only(40, 0.5, 3, 4_000_000))
            out.append({"what": "Brownian generation C4 (4M paths x 40 x 3)", "FMB_BM_TMA": tma, "ms": ms, "GBps": 4e6 * 120 * 8 / ms / 1e6})
    if "c3" in which:
        P = 2_000_000
        for tma in ("0", "1"):
            os.environ["FMB_BM_TMA"] = tma
            ms = best(bm_only(1000, 0.005, 2, P), reps=2)
            out.append({"what": "Brownian generation C3 shape (2M paths x 1000 x 2)", "FMB_BM_TMA": tma, "ms": ms, "GBps": P * 2000 * 8 / ms / 1e6})
        os.environ["FMB_BM_TMA"] = "1"
        td = pkg.TimeDiscretizationFromArray(0.0, 1000, 0.005)
        bm = pkg.BrownianMotionCuda(td, 2, P, 31415)
        bm.getBrownianIncrement(0, 0)
        model = pkg.HestonModel(1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.5, 0.1, 1, bm.randomVariableFactory)
        for tma, cfg in (("0", "-"), ("1", "0"), ("1", "1"), ("1", "2"), ("1", "3"), ("1", "4"), ("1", "5")):
            os.environ["FMB_EULER_TMA"] = tma
            if cfg != "-":
                os.environ["FMB_TMA_CFG"] = cfg

            def run():
                pkg.EulerSchemeFromProcessModel(model, bm).getProcessValue(1000, 0)
            ms = best(run, reps=2)
            out.append({"what": "Heston Euler kernel (2M paths x 1000)", "FMB_EULER_TMA": tma, "FMB_TMA_CFG (NT,KS,S: 0=256,2,4 1=256,4,3 2=256,8,2 3=128,4,4 4=128,8,3 5=128,2,6)": cfg,
                        "ms": ms, "GBps": P * 1000 * 32 / ms / 1e6})
        os.environ.pop("FMB_TMA_CFG", None)
        os.environ.pop("FMB_EULER_TMA", None)
        del bm
        nv.load().fmb_pool_trim()
    if "c2" in which:
        P = 4_000_000
        td = pkg.TimeDiscretizationFromArray(0.0, 200, 0.1)
        vt = np.arange(0, 21.0)
        vm = pkg.ShortRateVolatilityModelAsGiven(pkg.TimeDiscretizationFromArray(vt), 0.005 + 0.0005 * np.floor(vt) / 20, np.full(vt.size, 0.1))
        bm = pkg.BrownianMotionCuda(td, 2, P, 3141)
        bm.getBrownianIncrement(0, 0)
        hw = pkg.HullWhiteModel(bm.randomVariableFactory, pkg.TimeDiscretizationFromArray(0.0, 40, 0.5), vm)
        pkg.EulerSchemeFromProcessModel(hw, bm, 0).getProcessValue(200, 1)          # host-side coefficient tables built once
        for tma, cfg in (("0", "-"), ("1", "0"), ("1", "1"), ("1", "2"), ("1", "3"), ("1", "4"), ("1", "5")):
            os.environ["FMB_EULER_TMA"] = tma
            if cfg != "-":
                os.environ["FMB_TMA_CFG"] = cfg
            best_ms = 1e9
            for _ in range(3):
                proc = pkg.EulerSchemeFromProcessModel(hw, bm, 0)
                spec = hw.getFusedSpecification(proc)                                   # (host work outside the timed region)
                nv.synchronize()
                nv.timer_start()
                proc._precalculate_fused(spec)
                best_ms = min(best_ms, nv.timer_stop_ms())
            out.append({"what": "Hull-White Euler kernel (4M paths x 200)", "FMB_EULER_TMA": tma, "FMB_TMA_CFG": cfg, "ms": best_ms, "GBps": P * 200 * 32 / best_ms / 1e6})
        os.environ.pop("FMB_TMA_CFG", None)
        os.environ.pop("FMB_EULER_TMA", None)
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
