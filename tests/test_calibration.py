"""calibration.py (second half): the Levenberg-Marquardt optimiser on the reference's own test problems
(T/optimizer/LevenbergMarquardtTest.java:30-330) and the generic calibration of a parametric LIBOR covariance model
(J/montecarlo/interestrate/models/covariance/AbstractLIBORCovarianceModelParametric.java:333-470) — on the numpy stand-in (``-m "not gpu"``:
generic Euler recipe) and on the device (``-m gpu``: every evaluation is one fused-kernel simulation on resident increments)."""
import time

import numpy as np
import pytest

from numpy_rv import make
from common import lmm_setup


# ---- the optimiser (host logic) ------------------------------------------------------------------------------------------------------
def test_small_linear_system_and_clone_with_new_targets(pkg):
    class Problem(pkg.LevenbergMarquardt):
        def setValues(self, p, v):
            v[0] = p[0] * 0.0 + p[1]
            v[1] = p[0] * 2.0 + p[1]
    o = Problem()
    o.setInitialParameters([0, 0]).setWeights([1, 1]).setMaxIteration(100).setTargetValues([5, 10])
    o.run()
    best = o.getBestFitParameters()
    assert abs(best[0] - 2.5) < 1e-12 and abs(best[1] - 5.0) < 1e-12
    with pytest.raises(NotImplementedError):                 # "Solver cannot be modified after it has run."
        o.setMaxIteration(3)
    o2 = o.getCloneWithModifiedTargetValues([5.1, 10.2], [1, 1], True)
    o2.run()
    best2 = o2.getBestFitParameters()
    assert abs(best2[0] - 2.55) < 1e-12 and abs(best2[1] - 5.10) < 1e-12
    assert o2.getIterations() <= o.getIterations()           # starts from the previous optimum


def test_nonlinear_systems(pkg):
    def f(p, v):
        v[0] = 1.0 * p[0] + 2.0 * p[1] + p[2] + p[0] * p[1]
        v[1] = 2.0 * p[0] + 1.0 * p[1] + p[2] + p[1] * p[2]
        v[2] = 3.0 * p[0] + 0.0 * p[1] + p[2]
    o = pkg.LevenbergMarquardt([0, 0, 0], [5, 10, 2], 100, 10, objectiveFunction=f)
    o.run()
    assert o.getRootMeanSquaredError() < 1e-1

    class Rosenbrock(pkg.LevenbergMarquardt):
        def setValues(self, p, v):
            v[0] = 10.0 * (p[1] - p[0] * p[0])
            v[1] = 1.0 - p[0]
    r = Rosenbrock([0.5, 0.5], [0.0, 0.0], 100, 10)
    r.run()
    assert abs(r.getBestFitParameters()[0] - 1.0) < 1e-10 and abs(r.getBestFitParameters()[1] - 1.0) < 1e-10


def test_booth_function_with_steps_and_with_analytic_derivative(pkg):
    def booth(p, v):
        v[0] = (p[0] + 2 * p[1] - 7) ** 2 + (2 * p[0] + p[1] - 5) ** 2
    o = pkg.LevenbergMarquardt([2.0, 2.0], [0.0], 1000, objectiveFunction=booth)
    o.setParameterSteps([1e-8, 1e-8])
    o.run()
    assert abs(o.getRootMeanSquaredError()) < 2e-4

    class Analytic(pkg.LevenbergMarquardt):
        def setValues(self, p, v):
            booth(p, v)

        def setDerivatives(self, p, d):
            d[0][0] = (p[0] + 2 * p[1] - 7) * 2 + (2 * p[0] + p[1] - 5) * 4
            d[1][0] = (p[0] + 2 * p[1] - 7) * 4 + (2 * p[0] + p[1] - 5) * 2
    a = Analytic([2.0, 2.0], [0.0], 1000)
    a.run()
    assert abs(a.getRootMeanSquaredError()) < 2e-4
    # Levenberg regularisation, error tolerance, lambda accessors, failing evaluations
    l = pkg.LevenbergMarquardt([2.0, 2.0], [0.0], 1000, regularizationMethod=pkg.RegularizationMethod.LEVENBERG, objectiveFunction=booth)
    l.setErrorTolerance(1e-9)
    l.run()
    assert l.getRootMeanSquaredError() < 1e-3 and l.getIterations() < 1000
    with pytest.raises(ValueError):
        l.setLambdaDivisor(1.0)
    with pytest.raises(ValueError):
        l.setLambdaMultiplicator(0.5)

    def sometimes_nan(p, v):
        v[0] = float("nan") if p[0] > 3.0 else (p[0] - 1.0)
    n = pkg.LevenbergMarquardt([0.0], [0.0], 50, objectiveFunction=sometimes_nan)
    n.run()
    assert abs(n.getBestFitParameters()[0] - 1.0) < 1e-9
    assert np.allclose(pkg.solveLinearEquationSVD([[2.0, 0.0], [0.0, 0.0]], [4.0, 1.0]), [2.0, 0.0])     # pseudo-inverse of a singular system


# ---- calibration of the covariance model ---------------------------------------------------------------------------------------------
def _environment(request, pkg):
    if request.param == "numpy":
        RV, Factory, BM = make(pkg)
        return pkg, Factory, BM, False
    pkg.native.init(0)
    return pkg, pkg.RandomVariableCudaFactory, pkg.BrownianMotionCuda, True


@pytest.fixture(scope="module", params=["numpy", pytest.param("device", marks=pytest.mark.gpu)])
def env(request, pkg):
    return _environment(request, pkg)


def _covariance(pkg, s, a, b, c, d, decay, calibrateVol=True, calibrateCorr=True):
    vol = pkg.LIBORVolatilityModelFourParameterExponentialForm(s["sim"], s["tenor"], a, b, c, d, calibrateVol)
    corr = pkg.LIBORCorrelationModelExponentialDecay(s["sim"], s["tenor"], s["F"], decay, calibrateCorr)
    return pkg.LIBORCovarianceModelFromVolatilityAndCorrelation(s["sim"], s["tenor"], vol, corr)


def _swaptions(pkg, s, exercises, tenors, rate=0.05):
    period = s["tenor"].getTimeStep(0)
    out = []
    for e in exercises:
        for n in tenors:
            fix = [e + i * period for i in range(n)]
            if fix[-1] + period > s["tenor"].getTime(s["N"]) + 1e-12:
                continue
            out.append(pkg.Swaption(e, fix, [t + period for t in fix], [rate] * n))
    return out


def test_parametric_interface(pkg):
    s = lmm_setup(pkg, n_libors=8, n_factors=2, period=0.5, dt=0.5)
    cov = _covariance(pkg, s, 0.2, 0.05, 0.25, 0.3, 0.1)
    assert cov.getParameterAsDouble() == [0.2, 0.05, 0.25, 0.3, 0.1]
    clone = cov.getCloneWithModifiedParameters([0.1, 0.0, 0.2, 0.25, 0.05])
    assert clone.getParameterAsDouble() == [0.1, 0.0, 0.2, 0.25, 0.05] and cov.getParameterAsDouble()[0] == 0.2
    volOnly = _covariance(pkg, s, 0.2, 0.05, 0.25, 0.3, 0.1, calibrateCorr=False)
    assert volOnly.getParameterAsDouble() == [0.2, 0.05, 0.25, 0.3]
    assert volOnly.getCloneWithModifiedParameters([0.1, 0.0, 0.2, 0.25]).getCorrelationModel() is volOnly.getCorrelationModel()
    corrOnly = _covariance(pkg, s, 0.2, 0.05, 0.25, 0.3, 0.1, calibrateVol=False)
    assert corrOnly.getParameterAsDouble() == [0.1] and corrOnly.getVolatilityModel().getParameterAsDouble() is None
    assert corrOnly.getCloneWithModifiedParameters([0.1]).getCorrelationModel() is corrOnly.getCorrelationModel()      # unchanged parameter: same object
    # the vectorised tables equal the element-wise definitions bit for bit
    fl, var = cov.getFactorLoadingTable()
    vm, cm = cov.getVolatilityModel(), cov.getCorrelationModel()
    for t in (0, 3, 7):
        for j in (0, 4, 7):
            v = vm.getVolatility(t, j)
            assert var[t, j] == (v * v) * cm.getCorrelation(t, j, j)
            assert all(fl[t, j, k] == v * cm.getFactorLoading(t, k, j) for k in range(2))
    item = pkg.CalibrationProduct(pkg.Caplet(1.0, 0.5, 0.05), 0.01, 2.0, name="caplet")
    assert (item.getName(), item.getTargetValue().doubleValue(), item.getWeight(), item.getPriority()) == ("caplet", 0.01, 2.0, 0)
    with pytest.raises(TypeError):
        class Fixed:
            pass
        pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], None, None, Fixed(), [item], None)


def test_calibration_recovers_the_prices_of_a_known_model(env):
    pkg, Factory, BM, device = env
    n = 10 if device else 6
    paths = 20_000 if device else 2_000
    s = lmm_setup(pkg, n_libors=n, n_factors=2, period=0.5, dt=0.5)
    truth = _covariance(pkg, s, 0.25, 0.02, 0.30, 0.20, 0.15)
    bm = BM(s["sim"], s["F"], paths, 31415, Factory())
    products = _swaptions(pkg, s, [0.5, 1.0, 1.5, 2.0], [1, 2, 3])
    assert len(products) >= 9

    def simulate(cov):
        model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], Factory(), cov, None, {"measure": "SPOT"})
        return model, pkg.LIBORMonteCarloSimulationFromLIBORModel(pkg.EulerSchemeFromProcessModel(model, bm))
    _, simTruth = simulate(truth)
    targets = [p.getValue(simTruth) for p in products]
    items = [pkg.CalibrationProduct(p, t, 1.0) for p, t in zip(products, targets)]
    start = _covariance(pkg, s, 0.15, 0.0, 0.20, 0.30, 0.05)
    model0, sim0 = simulate(start)
    error0 = np.sqrt(np.mean([(p.getValue(sim0) - t) ** 2 for p, t in zip(products, targets)]))
    t0 = time.perf_counter()
    calibrated = start.getCloneCalibrated(model0, items, {"brownianMotion": bm, "maxIterations": 60, "accuracy": 1e-12, "parameterStep": 1e-5})
    elapsed = time.perf_counter() - t0
    info = calibrated.lastCalibration
    _, sim1 = simulate(calibrated)
    error1 = np.sqrt(np.mean([(p.getValue(sim1) - t) ** 2 for p, t in zip(products, targets)]))
    assert abs(error1 - info["rootMeanSquaredError"]) < 1e-12
    assert error0 > 1e-4 and error1 < 2e-7 and error1 < 1e-3 * error0, (error0, error1, info)
    assert info["evaluations"] >= info["iterations"] and list(info["bestParameters"]) == calibrated.getParameterAsDouble()
    if device:
        assert sim1.getProcess().usedFusedKernel == "lmm"
        print("calibration: %d evaluations in %.3f s (%.2f ms per evaluation: simulation + %d swaptions), rms %.2e -> %.2e" % (
            info["evaluations"], elapsed, 1e3 * elapsed / info["evaluations"], len(products), error0, error1))
    # the same through the model factory (calibration items given): LIBORMarketModelFromCovarianceModel.java:296-318
    viaOf = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], Factory(), start, items,
                                                       {"measure": "SPOT", "calibrationParameters": {"brownianMotion": bm, "maxIterations": 60,
                                                                                                     "accuracy": 1e-12, "parameterStep": 1e-5}})
    assert viaOf.getCovarianceModel().getParameterAsDouble() == calibrated.getParameterAsDouble()
