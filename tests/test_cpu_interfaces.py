"""-m "not gpu": the host-side methods of the reference's interfaces either side of the hot path (TimeDiscretization, LIBORCovarianceModel,
LIBORMarketModel / TermStructureModel, the simulation models' clone family, RandomVariableFactory.getRandomVariableOrDefault), on
deterministic values and on the numpy stand-in of tests/numpy_rv.py (the product has no CPU arithmetic for stochastic values)."""
import numpy as np
import pytest

import numpy_rv
from common import lmm_setup


def test_time_discretization_set_operations_and_lookups(pkg):
    """TimeDiscretizationFromArray.java:227-249 (stub period), :286-292, :301-352 and TimeDiscretization.java:84-120."""
    T = pkg.TimeDiscretizationFromArray
    a, b = T(0.0, 4, 0.5), T([0.25, 0.5, 1.0, 3.0])
    assert a.union(b).getAsArrayList() == [0.0, 0.25, 0.5, 1.0, 1.5, 2.0, 3.0]
    assert a.intersect(b).getAsArrayList() == [0.5, 1.0]
    assert a.filter(lambda t: t >= 1.0).getAsArrayList() == [1.0, 1.5, 2.0]
    assert (a.getFirstTime(), a.getLastTime(), a.getTickSize()) == (0.0, 2.0, 1.0 / (365.0 * 24.0))
    assert [a.getTimeIndexNearestGreaterOrEqual(t) for t in (-1.0, 0.0, 0.6, 1.0, 5.0)] == [0, 0, 2, 2, 5]
    assert [a.getTimeIndexNearestLessOrEqual(t) for t in (0.6, 1.0, 5.0)] == [1, 2, 4]
    assert list(a) == list(a.doubleStream()) == a.getAsArrayList() == list(a.getAsDoubleArray())
    shifted = a.getTimeShiftedTimeDiscretization(0.1)
    assert shifted.getAsArrayList() == [T._round(a, t + 0.1) for t in a.getAsArrayList()] and shifted.getTime(0) == 0.09999999999999999
    assert T(0.0, 1.1, 0.5, T.SHORT_PERIOD_AT_END).getAsArrayList() == [0.0, 0.5, 1.0, T._round(a, 1.1)]
    assert T(0.0, 1.1, 0.5, T.SHORT_PERIOD_AT_START).getAsArrayList() == [0.0, T._round(a, 0.1), T._round(a, 0.6), T._round(a, 1.1)]
    coarse = T([0.0, 1.0], tickSize=0.5)
    assert a.union(coarse).getTickSize() == a.getTickSize() and a.intersect(coarse).getTickSize() == 0.5      # finer for union, coarser for intersect
    assert "timeTickSize" in repr(a)


def test_covariance_model_pointwise_accessors_equal_the_table(pkg):
    """LIBORCovarianceModelFromVolatilityAndCorrelation.java:47-93 and the time / fixing-date overloads of AbstractLIBORCovarianceModel.java:45-76
    against getFactorLoadingTable (what the fused kernel is fed)."""
    s = lmm_setup(pkg)
    cov = s["cov"]
    fl, var = cov.getFactorLoadingTable()
    for t, j in ((0, 0), (3, 10), (39, 39), (5, 2)):
        assert [x.doubleValue() for x in cov.getFactorLoading(t, j)] == list(fl[t, j])
        assert cov.getCovariance(t, j, j).doubleValue() == var[t, j]
    assert [x.doubleValue() for x in cov.getFactorLoading(1.5, 5.0)] == list(fl[3, 10])           # (time, fixing date)
    assert [x.doubleValue() for x in cov.getFactorLoading(1.7, 10)] == list(fl[3, 10])            # (time between grid points, component index)
    vol = cov.getVolatilityModel()
    corr = cov.getCorrelationModel()
    assert cov.getCovariance(1.6, 10, 12).doubleValue() == vol.getVolatility(3, 10) * vol.getVolatility(3, 12) * corr.getCorrelation(3, 10, 12)
    weight = sum(corr.getFactorLoading(3, 1, c) ** 2 for c in range(40))
    assert cov.getFactorLoadingPseudoInverse(3, 10, 1).doubleValue() == (1.0 / vol.getVolatility(3, 10)) * corr.getFactorLoading(3, 1, 10) * (1 / weight)


def test_integrated_libor_covariance_equals_the_reference_loop(pkg):
    """LIBORMarketModelFromCovarianceModel.java:1552-1596 restated as plain loops (incl. the unsymmetrised first time index)."""
    s = lmm_setup(pkg, n_libors=12, n_factors=2)
    model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], None, s["cov"], None, {"measure": "SPOT"})
    got = model.getIntegratedLIBORCovariance(s["sim"])
    assert got is model.getIntegratedLIBORCovariance(s["sim"])
    fl, _ = s["cov"].getFactorLoadingTable()
    T, N, F = fl.shape
    ref = np.zeros((T, N, N))
    for t in range(T):
        dt = s["sim"].getTimeStep(t)
        for i in range(N):
            for j in range(i, N):
                v = 0.0
                if s["tenor"].getTime(i) > s["sim"].getTime(t):
                    for f in range(F):
                        v += fl[t, i, f] * fl[t, j, f] * dt
                ref[t, i, j] = v
    for t in range(1, T):
        for i in range(N):
            for j in range(i, N):
                ref[t, i, j] = ref[t - 1, i, j] + ref[t, i, j]
                ref[t, j, i] = ref[t, i, j]
    assert np.array_equal(got, ref)


def _numpy_lmm(pkg, s, paths=500, with_discount_curve=True):
    _, Factory, BM = numpy_rv.make(pkg)
    factory = Factory()
    model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"] if with_discount_curve else None, factory, s["cov"], None,
                                                       {"measure": "SPOT", "stateSpace": "LOGNORMAL"})
    bm = BM(s["sim"], s["F"], paths, 3141, factory)
    return pkg.LIBORMonteCarloSimulationFromLIBORModel(pkg.EulerSchemeFromProcessModel(model, bm))


def test_lmm_clone_family_and_simulation_accessors(pkg):
    s = lmm_setup(pkg, n_libors=8, n_factors=2)
    sim = _numpy_lmm(pkg, s)
    model = sim.getModel()
    assert sim.getNumberOfFactors() == 2 and sim.getNumberOfComponents() == 8 and sim.getReferenceDate() is None
    libors = sim.getLIBORs(3)
    assert len(libors) == 8 and all(np.array_equal(np.atleast_1d(a.getRealizations()), np.atleast_1d(sim.getLIBOR(3, c).getRealizations())) for c, a in enumerate(libors))
    # forward-rate shift (the bump of a delta): same covariance model and properties, shifted curve (:1679-1691)
    shifted = model.getCloneWithModifiedData({"forwardRateShift": np.full(8, 1e-4)})
    assert np.array_equal(shifted.L0, model.L0 + 1e-4) and shifted.covarianceModel is model.covarianceModel
    assert (shifted.measure, shifted.stateSpace, shifted.liborCap) == (model.measure, model.stateSpace, model.liborCap)
    with pytest.raises(RuntimeError):
        model.getCloneWithModifiedData({"swaptionMarketData": object()})
    # only the discount curve changes: the simulated paths are re-used (LIBORMonteCarloSimulationFromLIBORModel.java:176-181); else re-simulated
    df2 = np.asarray(s["df"]) * np.exp(-0.001 * np.arange(len(s["df"])))
    other = sim.getCloneWithModifiedData("discountCurve", df2)
    assert other.getProcess() is sim.getProcess() and np.array_equal(other.getModel().discountFactors, df2)
    bumped = sim.getCloneWithModifiedData({"forwardRateShift": np.full(8, 1e-4)})
    assert bumped.getProcess() is not sim.getProcess() and bumped.getProcess().getStochasticDriver() is sim.getProcess().getStochasticDriver()
    up, base = bumped.getLIBOR(2, 5).getAverage(), sim.getLIBOR(2, 5).getAverage()
    assert 0.5e-4 < up - base < 2e-4
    # forward discount bond seen from time 0 is the discount curve's forward bond (:947-952)
    bond = model.getForwardDiscountBond(sim.getProcess(), 0.0, 2.0)
    assert bond.getAverage() == pytest.approx(s["df"][4] / s["df"][0], rel=1e-13)
    with pytest.raises(ValueError):
        _numpy_lmm(pkg, s, with_discount_curve=False).getModel().getForwardDiscountBond(sim.getProcess(), 0.0, 2.0)


def test_asset_model_clone_family(pkg):
    """BlackScholesModel.java:157-166, HestonModel.java:448-465, MonteCarloAssetModel.java:116-150, EulerSchemeFromProcessModel.java:370-391,
    MonteCarloBlackScholesModel.java:112-155."""
    _, Factory, BM = numpy_rv.make(pkg)
    td = pkg.TimeDiscretizationFromArray(0.0, 10, 0.1)
    bm = BM(td, 1, 4000, 3141)
    sim = pkg.MonteCarloAssetModel(pkg.BlackScholesModel(1.0, 0.05, 0.2, Factory()), bm)
    up = sim.getCloneWithModifiedData({"initialValue": 1.1})
    assert up.getProcess().getStochasticDriver() is bm and up.getModel().getVolatility().doubleValue() == 0.2
    assert up.getAssetValue(1.0, 0).getAverage() / sim.getAssetValue(1.0, 0).getAverage() == pytest.approx(1.1, rel=1e-12)      # same paths, scaled
    other = sim.getCloneWithModifiedData({"volatility": 0.3, "scheme": pkg.Scheme.EULER})
    assert other.getProcess().scheme == pkg.Scheme.EULER and other.getModel().getVolatility().doubleValue() == 0.3
    assert sim.getReferenceDate() is None
    with pytest.raises(NotImplementedError):
        sim.getCloneWithModifiedSeed(1)
    reseeded = sim.getProcess().getCloneWithModifiedData({"seed": 7})
    assert reseeded.getStochasticDriver().seed == 7 and reseeded.getModel() is sim.getModel()
    with pytest.raises(ValueError):
        sim.getProcess().getCloneWithModifiedData({"seed": 7, "stochasticDriver": bm})
    h = pkg.HestonModel(1.0, 0.05, 0.2, 0.05, 0.04, 1.0, 0.3, -0.5, pkg.HestonModel.FULL_TRUNCATION, Factory())
    h2 = h.getCloneWithModifiedData({"xi": 0.0, "rho": pkg.Scalar(0.1)})
    assert (h2.xi.doubleValue(), h2.rho.doubleValue(), h2.kappa.doubleValue(), h2.scheme) == (0.0, 0.1, 1.0, h.scheme)
    # the Black-Scholes convenience class: clones sit on a NEW driver with the class's seed / the requested seed (nothing runs without a GPU)
    mc = pkg.MonteCarloBlackScholesModel(td, 1000, 1.0, 0.05, 0.2)
    c = mc.getCloneWithModifiedData({"riskFreeRate": 0.01})
    assert isinstance(c, pkg.MonteCarloBlackScholesModel) and c.getModel().getRiskFreeRate().doubleValue() == 0.01
    assert c.getProcess().getStochasticDriver() is not mc.getProcess().getStochasticDriver() and c.getProcess().getStochasticDriver().seed == 3141
    assert mc.getCloneWithModifiedSeed(17).getProcess().getStochasticDriver().seed == 17


def test_factory_value_or_default(pkg):
    F = pkg.RandomVariableCudaFactory
    assert F.getRandomVariableOrDefault(None, None, pkg.Scalar(2.0)).doubleValue() == 2.0
    x = pkg.Scalar(5.0)
    assert F.getRandomVariableOrDefault(None, x, None) is x
    assert F.getRandomVariableOrDefault(F(), 3, None).doubleValue() == 3.0
    with pytest.raises(TypeError):
        F.getRandomVariableOrDefault(None, 3.0, None)
    with pytest.raises(ValueError):
        F.getRandomVariableOrDefault(F(), "3", None)


def test_covariance_model_clone_with_modified_data(pkg):
    """LIBORCovarianceModelFromVolatilityAndCorrelation.java:180-208, LIBORVolatilityModelFourParameterExponentialForm.java:212-253,
    LIBORCorrelationModelExponentialDecay.java:150-167: a changed grid is handed down to the parts unless a part is given explicitly."""
    s = lmm_setup(pkg, n_libors=8, n_factors=2)
    cov = s["cov"]
    finer = pkg.TimeDiscretizationFromArray(0.0, 16, 0.25)
    c = cov.getCloneWithModifiedData({"timeDiscretization": finer})
    assert c.getTimeDiscretization() is finer and c.getVolatilityModel().getTimeDiscretization() is finer and c.getCorrelationModel().getTimeDiscretization() is finer
    assert c.getLiborPeriodDiscretization() is cov.getLiborPeriodDiscretization()
    assert c.getFactorLoadingTable()[0].shape == (16, 8, 2)
    assert np.array_equal(c.getFactorLoadingTable()[0][::2], cov.getFactorLoadingTable()[0])        # the same function of (t, T_j) on the common times
    v = cov.getVolatilityModel().getCloneWithModifiedData({"a": 0.5, "d": pkg.Scalar(0.1), "isCalibrateable": True})
    assert (v.a, v.b, v.c, v.d, v.isCalibrateable) == (0.5, cov.getVolatilityModel().b, cov.getVolatilityModel().c, 0.1, True)
    k = cov.getCorrelationModel().getCloneWithModifiedData({"numberOfFactors": 3, "a": 0.2})
    assert k.getNumberOfFactors() == 3 and k.a == 0.2
    given = cov.getCloneWithModifiedData({"volatilityModel": v})
    assert given.getVolatilityModel() is v and given.getCorrelationModel() is cov.getCorrelationModel()


def test_product_accessors_and_bump_and_revalue(pkg):
    """AbstractMonteCarloProduct.java:87-161 (getValues / getValuesForModifiedData) and the getters of the products on the path."""
    _, Factory, BM = numpy_rv.make(pkg)
    td = pkg.TimeDiscretizationFromArray(0.0, 10, 0.1)
    sim = pkg.MonteCarloAssetModel(pkg.BlackScholesModel(1.0, 0.05, 0.2, Factory()), BM(td, 1, 20000, 3141))
    option = pkg.EuropeanOption(1.0, 1.05)
    assert (option.getMaturity(), option.getStrike(), option.getCallOrPut(), option.getUnderlyingIndex(), option.getCurrency()) == (1.0, 1.05, 1.0, 0, None)
    v = option.getValues(sim)
    assert v["value"] == option.getValue(sim) and v["error"] == option.getValue(0.0, sim).getStandardError() and 0 < v["error"] < 0.01
    up = option.getValuesForModifiedData(sim, "initialValue", 1.01)
    same = option.getValuesForModifiedData(0.0, sim, {"initialValue": 1.01})
    assert up == same
    delta = (up["value"] - v["value"]) / 0.01
    assert 0.4 < delta < 0.8                                  # Black-Scholes delta of this slightly out-of-the-money call is about 0.58
    s = lmm_setup(pkg, n_libors=8, n_factors=2)
    lmm = _numpy_lmm(pkg, s, paths=2000)
    swaption = pkg.Swaption(1.0, [1.0, 1.5, 2.0], [1.5, 2.0, 2.5], [0.05] * 3)
    assert (swaption.getExerciseDate(), swaption.getFixingDates(), swaption.getPaymentDates(), swaption.getSwaprates(), swaption.getNotional()) == \
        (1.0, [1.0, 1.5, 2.0], [1.5, 2.0, 2.5], [0.05] * 3, 1.0)
    indicator = np.asarray(swaption.getExerciseIndicator(lmm).getRealizations())
    value = np.asarray(swaption.getValue(1.0, lmm).getRealizations())
    assert set(np.unique(indicator)) <= {0.0, 1.0} and np.array_equal(indicator == 1.0, value > 0)
    bermudan = pkg.BermudanSwaption([True, False, True], [1.0, 1.5, 2.0], [0.5] * 3, [1.5, 2.0, 2.5], [1.0] * 3, [0.05] * 3)
    assert bermudan.getExerciseTimes() == [1.0, 2.0] and bermudan.getFinalMaturity() == 2.5 and bermudan.getIsCallable() is True
    assert bermudan.getSwapRates() == [0.05] * 3 and bermudan.getPeriodNotionals() == [1.0] * 3
    assert _numpy_lmm(pkg, s).getModel().getNumeraireAdjustments() == {}        # nothing evaluated yet
    model = lmm.getModel()
    lmm.getNumeraire(1.0)
    adjustments = model.getNumeraireAdjustments()
    assert sorted(adjustments) == [0.5 * i for i in range(8)]
    assert adjustments[0.5].doubleValue() == (s["df"][1] / s["df"][2] - 1.0) / 0.5
    assert model.clone().covarianceModel is model.covarianceModel and model.getMeasure() == model.SPOT
    heston = pkg.HestonModel(1.0, 0.05, 0.2, 0.05, 0.04, 1.0, 0.3, -0.5, pkg.HestonModel.REFLECTION, Factory())
    assert [g().doubleValue() for g in (heston.getInitialValue, heston.getRiskFreeRate, heston.getVolatility, heston.getTheta, heston.getKappa, heston.getXi, heston.getRho)] == \
        [1.0, 0.05, 0.2, 0.04, 1.0, 0.3, -0.5] and heston.getScheme() == pkg.HestonModel.REFLECTION


def test_hull_white_and_parametric_accessors(pkg):
    """HullWhiteModel.java:360-365, :797-812; getParameter / clone of the parametric covariance parts; the regression's basis-function suppliers."""
    tenor = pkg.TimeDiscretizationFromArray(0.0, 10, 0.5)
    voltd = pkg.TimeDiscretizationFromArray([0.0, 2.0])
    vm = pkg.ShortRateVolatilityModelAsGiven(voltd, [0.01, 0.012], [0.1, 0.1])
    df = [np.exp(-0.03 * 0.5 * i) for i in range(11)]
    dff = [np.exp(-0.04 * 0.5 * i) for i in range(11)]
    hw = pkg.HullWhiteModel(None, tenor, vm, None, df, dff)
    assert hw.getVolatilityModel() is vm
    vm2 = pkg.ShortRateVolatilityModelAsGiven(voltd, [0.02, 0.02], [0.1, 0.1])
    clone = hw.getCloneWithModifiedVolatilityModel(vm2)
    assert clone.getVolatilityModel() is vm2 and clone.getLiborPeriodDiscretization() is tenor and np.array_equal(clone.dfDiscount, hw.dfDiscount)
    v = hw.getIntegratedBondSquaredVolatility(1.0, 3.0).doubleValue()
    assert v == hw.getShortRateConditionalVariance(0, 1.0).mult(hw.getB(1.0, 3.0).squared()).doubleValue() and v > 0
    s = lmm_setup(pkg, n_libors=8, n_factors=2)
    cov = s["cov"]
    vol, corr = cov.getVolatilityModel(), cov.getCorrelationModel()
    assert [x.doubleValue() for x in cov.getParameter()] == cov.getParameterAsDouble()
    assert (vol.getParameter() is None) == (vol.getParameterAsDouble() is None) and (corr.getParameter() is None) == (corr.getParameterAsDouble() is None)
    c = cov.clone()
    assert c is not cov and np.array_equal(c.getFactorLoadingTable()[0], cov.getFactorLoadingTable()[0])
    one, x = pkg.Scalar(1.0), pkg.Scalar(2.0)
    est = pkg.MonteCarloConditionalExpectationRegression([one, x])
    assert est.getBasisFunctionsEstimator().getBasisFunctions() == [one, x] and est.getBasisFunctionsPredictor().getBasisFunctions() == [one, x]


def test_lmm_model_parameters(pkg):
    """LIBORMarketModelFromCovarianceModel.java:1699-1733."""
    s = lmm_setup(pkg, n_libors=4, n_factors=2)
    lmm = _numpy_lmm(pkg, s)
    before = lmm.getModelParameters()
    assert before["FORWARD(0.0,0.5)"] is None and not any(k.startswith("NUMERAIREADJUSTMENT") for k in before)      # no process seen yet
    lmm.getNumeraire(1.0)
    p = lmm.getModelParameters()
    assert list(p) == sorted(p)
    assert [k for k in p if k.startswith("FORWARD")] == ["FORWARD(0.0,0.5)", "FORWARD(0.5,1.0)", "FORWARD(1.0,1.5)", "FORWARD(1.5,2.0)"]
    assert p["FORWARD(1.0,1.5)"].doubleValue() == pytest.approx(s["L0"][2], rel=1e-15)
    assert p["NUMERAIREADJUSTMENT(0.5)"].doubleValue() == (s["df"][1] / s["df"][2] - 1.0) / 0.5
    assert ("COVARIANCEMODELPARAMETER(0)" in p) == bool(s["cov"].getParameterAsDouble())
