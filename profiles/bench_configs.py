#!/usr/bin/env python
"""Side benchmark: path-steps/s of the other BASELINE.json configurations (C1 Black-Scholes, C2 Hull-White, C3 Heston, C5 Bermudan)
on one B200, through the host API.  Output: one JSON object per configuration (kept in profiles/r01_configs.jsonl).
Not the driver's bench (that is /bench.py, C4); timing with device events on the library's stream, 1 warm-up + 3 timed repetitions."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

pkg = None
nv = None


def bind(package):
    global pkg, nv
    pkg, nv = package, package.native


def timed(fn, reps=3):
    fn(0)
    nv.synchronize()
    ms = []
    for r in range(reps):
        nv.timer_start()
        fn(r + 1)
        ms.append(nv.timer_stop_ms())
    return float(np.mean(ms))


def report(name, paths, steps, bytes_per_path_step, ms, extra=None, emit=True):
    d = {"config": name, "paths": paths, "steps": steps, "ms": ms, "path_steps_per_s": paths * steps / (ms * 1e-3),
         "algorithmic_GBps": paths * steps * bytes_per_path_step / (ms * 1e-3) / 1e9}
    d.update(extra or {})
    if emit:
        print(json.dumps(d), flush=True)
    return d


def kernel_only(process, model):
    """Device time of the fused Euler kernel alone (host-side table set-up outside the timed region), best of 3."""
    best_ms = 1e9
    for _ in range(3):
        proc = process()
        spec = model.getFusedSpecification(proc)
        proc.stochasticDriver.getBrownianIncrement(0, 0)
        nv.synchronize()
        nv.timer_start()
        proc._precalculate_fused(spec)
        best_ms = min(best_ms, nv.timer_stop_ms())
        del proc
    return best_ms


def c2_kernel():
    td = pkg.TimeDiscretizationFromArray(0.0, 200, 0.1)
    vt = np.arange(0, 21.0)
    vm = pkg.ShortRateVolatilityModelAsGiven(pkg.TimeDiscretizationFromArray(vt), 0.005 + 0.0005 * np.floor(vt) / 20, np.full(vt.size, 0.1))
    bm = pkg.BrownianMotionCuda(td, 2, 1_000_000, 3141)
    hw = pkg.HullWhiteModel(bm.randomVariableFactory, pkg.TimeDiscretizationFromArray(0.0, 40, 0.5), vm)
    return kernel_only(lambda: pkg.EulerSchemeFromProcessModel(hw, bm, 0), hw)


def c3_kernel(paths):
    td = pkg.TimeDiscretizationFromArray(0.0, 1000, 0.005)
    bm = pkg.BrownianMotionCuda(td, 2, paths, 31415)
    model = pkg.HestonModel(1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.5, 0.1, 1, bm.randomVariableFactory)
    return kernel_only(lambda: pkg.EulerSchemeFromProcessModel(model, bm), model)


def run_all(package, c3_paths=4_000_000):
    """C1, C2, C3 for bench.py's `configs` key (C5 is measured by bench.py itself): 1 warm-up + 2 timed repetitions each."""
    bind(package)
    out = [report("C1 Black-Scholes 100k x 100, EULER_FUNCTIONAL, incl. call price", 100_000, 100, 16, timed(c1, reps=2), emit=False),
           report("C2 Hull-White 1M x 200 (dt 0.1y), piecewise-constant sigma(t), EULER (model built once)", 1_000_000, 200, 32, timed(c2, reps=2), emit=False)]
    nv.load().fmb_pool_trim()
    out.append(report("C3 Heston full truncation %dM x 1000, 8-strike smile" % (c3_paths // 1_000_000), c3_paths, 1000, 32, timed(c3(c3_paths), reps=2), emit=False))
    nv.load().fmb_pool_trim()
    # kernel-level numbers of the two-component Euler kernels (algorithmic 32 B per path-step: 16 read + 16 written)
    hbm = None
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ms = c2_kernel()
    gbs = 1_000_000 * 200 * 32 / ms / 1e6
    out[1]["euler_kernel"] = {"kernel": "eulerTwoFactorTmaKernel<HullWhiteStep,256,2,4> (bulk-copy + mbarrier pipeline)", "ms": ms, "achieved_GBps": gbs,
                              "hbm_frac": gbs / hbm if hbm else None, "bound": "hbm"}
    nv.load().fmb_pool_trim()
    kp = min(c3_paths, 2_000_000)
    ms = c3_kernel(kp)
    gbs = kp * 1000 * 32 / ms / 1e6
    out[2]["euler_kernel"] = {"kernel": "eulerHestonKernel", "paths": kp, "ms": ms, "achieved_GBps": gbs, "hbm_frac": gbs / hbm if hbm else None,
                              "bound": "issue / FP64 latency (exp + log + sqrt per step): ncu issue slots 59 %, FP64 pipe 43 %, DRAM 43 %"}
    nv.load().fmb_pool_trim()
    return out


def next_rows(package, calibration_paths=100_000, aad_paths=1_000_000):
    """SURVEY §8f rank 4 on the device, for bench.py's `f4` key (wall clock around whole host-API calls; N = 1 only):
    * calibration: the C4 model shape (40 rates x 3 factors x 40 steps), five covariance parameters against 16 swaption prices of a
      known model, Levenberg-Marquardt; every evaluation = table set-up + fused LMM kernel on resident increments + 16 valuations;
    * aad: Black-Scholes European call, 10 steps, forward recording + backward sweep for delta / rho / vega on device vectors."""
    bind(package)
    from common import lmm_setup
    s = lmm_setup(pkg)

    def cov(a, b, c, d, decay):
        vol = pkg.LIBORVolatilityModelFourParameterExponentialForm(s["sim"], s["tenor"], a, b, c, d, True)
        corr = pkg.LIBORCorrelationModelExponentialDecay(s["sim"], s["tenor"], s["F"], decay, True)
        return pkg.LIBORCovarianceModelFromVolatilityAndCorrelation(s["sim"], s["tenor"], vol, corr)
    factory = pkg.RandomVariableCudaFactory()
    bm = pkg.BrownianMotionCuda(s["sim"], s["F"], calibration_paths, 31415, factory)
    products = []
    for e in (1.0, 2.0, 5.0, 10.0):
        for n in (2, 4, 10, 20):
            fix = [e + 0.5 * i for i in range(n)]
            if fix[-1] + 0.5 <= 20.0:
                products.append(pkg.Swaption(e, fix, [t + 0.5 for t in fix], [0.05] * n))

    def simulate(c):
        m = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], factory, c, None, {"measure": "SPOT"})
        return m, pkg.LIBORMonteCarloSimulationFromLIBORModel(pkg.EulerSchemeFromProcessModel(m, bm))
    _, truth = simulate(cov(0.25, 0.02, 0.30, 0.20, 0.15))
    targets = [p.getValue(truth) for p in products]
    items = [pkg.CalibrationProduct(p, t, 1.0) for p, t in zip(products, targets)]
    start = cov(0.15, 0.0, 0.20, 0.30, 0.05)
    model0, sim0 = simulate(start)
    rms0 = float(np.sqrt(np.mean([(p.getValue(sim0) - t) ** 2 for p, t in zip(products, targets)])))
    nv.synchronize()
    t0 = time.perf_counter()
    done = start.getCloneCalibrated(model0, items, {"brownianMotion": bm, "maxIterations": 100, "accuracy": 1e-12, "parameterStep": 1e-5})
    wall = time.perf_counter() - t0
    info = done.lastCalibration
    calibration = {"model": "LMM 40 rates x 3 factors x 40 steps, %d paths, increments resident" % calibration_paths, "parameters": 5,
                   "products": len(products), "iterations": info["iterations"], "evaluations": info["evaluations"], "wall_ms": wall * 1e3,
                   "ms_per_evaluation": wall * 1e3 / info["evaluations"], "rms_start": rms0, "rms_end": info["rootMeanSquaredError"]}
    del truth, sim0, done
    nv.load().fmb_pool_trim()

    td = pkg.TimeDiscretizationFromArray(0.0, 10, 0.5)
    f = pkg.RandomVariableDifferentiableAADFactory(factory)
    bmA = pkg.BrownianMotionCuda(td, 1, aad_paths, 3141, factory)
    bmA.getBrownianIncrement(0, 0)

    def greeks():
        model = pkg.BlackScholesModel(1.0, 0.05, 0.30, f)
        mc = pkg.MonteCarloAssetModel(model, pkg.EulerSchemeFromProcessModel(model, bmA))
        nv.synchronize()
        t0 = time.perf_counter()
        launches0 = nv.launch_count()
        value = pkg.EuropeanOption(5.0, 1.05).getValueRV(0.0, mc)
        price = value.getAverage()
        t1 = time.perf_counter()
        launches1 = nv.launch_count()
        g = value.getGradient()
        out = [g[i].getAverage() for i in (model.getInitialValue()[0].getID(), model.getRiskFreeRate().getID(), model.getVolatility().getID())]
        t2 = time.perf_counter()
        return price, out, (t1 - t0) * 1e3, (t2 - t1) * 1e3, launches1 - launches0, nv.launch_count() - launches1
    greeks()
    price, g, fwd, bwd, lf, lb = greeks()
    aad = {"model": "Black-Scholes 10 steps, %d paths, European call: recorded generic Euler recipe on device vectors" % aad_paths,
           "price": price, "delta": g[0], "rho": g[1], "vega": g[2], "forward_ms": fwd, "backward_ms": bwd, "forward_launches": lf, "backward_launches": lb}
    nv.load().fmb_pool_trim()
    return {"calibration": calibration, "aad": aad}


def c1(seed):
    td = pkg.TimeDiscretizationFromArray(0.0, 100, 0.05)
    m = pkg.MonteCarloBlackScholesModel(1.0, 0.05, 0.30, pkg.BrownianMotionCuda(td, 1, 100_000, 3141 + seed))
    return pkg.EuropeanOption(5.0, 1.05).getValue(m)


_c2_model = {}


def c2(seed):
    # the model is built once (like the C4 model of bench.py); a repetition = Brownian generation + fused Euler evolution + one average
    if not _c2_model:
        vt = np.arange(0, 21.0)
        vm = pkg.ShortRateVolatilityModelAsGiven(pkg.TimeDiscretizationFromArray(vt), 0.005 + 0.0005 * np.floor(vt) / 20, np.full(vt.size, 0.1))
        _c2_model["factory"] = pkg.RandomVariableCudaFactory()
        _c2_model["td"] = pkg.TimeDiscretizationFromArray(0.0, 200, 0.1)
        _c2_model["model"] = pkg.HullWhiteModel(_c2_model["factory"], pkg.TimeDiscretizationFromArray(0.0, 40, 0.5), vm)
    bm = pkg.BrownianMotionCuda(_c2_model["td"], 2, 1_000_000, 3141 + seed, _c2_model["factory"])
    p = pkg.EulerSchemeFromProcessModel(_c2_model["model"], bm, 0)
    return p.getProcessValue(200, 1).getAverage()


def c3(paths):
    def run(seed):
        td = pkg.TimeDiscretizationFromArray(0.0, 1000, 0.005)
        bm = pkg.BrownianMotionCuda(td, 2, paths, 31415 + seed)
        model = pkg.HestonModel(1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.5, 0.1, 1, bm.randomVariableFactory)
        mc = pkg.MonteCarloAssetModel(model, bm)
        return [pkg.EuropeanOption(5.0, 1.10 * k).getValue(mc) for k in (0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4, 1.5)]
    return run


def c5(paths):
    from common import lmm_setup, lmm_device, bermudan_spec
    s = lmm_setup(pkg)
    b = bermudan_spec(s)
    product = pkg.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])

    def run(seed):
        return product.getValue(lmm_device(pkg, s, paths, seed=3141 + seed))
    return run


if __name__ == "__main__":
    bind(graft.load_package())
    nv.init(0)
    which = sys.argv[1:] or ["c1", "c2", "c3", "c5"]
    if "c1" in which:
        report("C1 Black-Scholes 100k x 100, EULER_FUNCTIONAL, incl. call price", 100_000, 100, 16, timed(c1))
    if "c2" in which:
        report("C2 Hull-White 1M x 200 (dt 0.1y), piecewise-constant sigma(t), EULER (model built once)", 1_000_000, 200, 32, timed(c2))
    if "c3" in which:
        for paths in (1_000_000, 4_000_000):
            t0 = time.time()
            report("C3 Heston full truncation %dM x 1000, 8-strike smile" % (paths // 1_000_000), paths, 1000, 32, timed(c3(paths), reps=2), {"wall_s_total": None})
            nv.load().fmb_pool_trim()
    if "c5" in which:
        for paths in (1_000_000, 8_000_000):
            report("C5 LMM Bermudan swaption %dM paths (simulate + 20 exercise dates x 6 basis functions)" % (paths // 1_000_000), paths, 40, 180, timed(c5(paths), reps=2))
            nv.load().fmb_pool_trim()
