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
