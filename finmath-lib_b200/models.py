"""ProcessModel mirrors (SDE specifications) and the simulation façades.  Host-side logic only: every model keeps the
reference's callback contract (J/montecarlo/model/ProcessModel.java:47-174) on RandomVariables — so the generic Euler
loop can run it unchanged — and additionally describes itself to the fused kernels through getFusedSpecification().

* BlackScholesModel                       J/montecarlo/assetderivativevaluation/models/BlackScholesModel.java:42-203
* HestonModel                             J/montecarlo/assetderivativevaluation/models/HestonModel.java:75-523
* LIBORMarketModelFromCovarianceModel     J/montecarlo/interestrate/models/LIBORMarketModelFromCovarianceModel.java:161-1757
  with LIBORVolatilityModelFourParameterExponentialForm (…/covariance/…:168-190), LIBORCorrelationModelExponentialDecay
  (…:87-134, PCA factor reduction J/functions/LinearAlgebra.java:392-482), LIBORCovarianceModelFromVolatilityAndCorrelation (…:47-93)
* MonteCarloAssetModel / MonteCarloBlackScholesModel   J/montecarlo/assetderivativevaluation/MonteCarloAssetModel.java:31-209
* LIBORMonteCarloSimulationFromLIBORModel  J/montecarlo/interestrate/LIBORMonteCarloSimulationFromLIBORModel.java:27-206
"""
import math
import weakref

import numpy as np

from .montecarlo import BrownianMotionCuda, EulerSchemeFromProcessModel, Scheme
from .stochastic import RandomVariable, RandomVariableCuda, RandomVariableCudaFactory, Scalar, _jexp
from . import native as nv


class BlackScholesModel:
    def __init__(self, initialValue, riskFreeRate, volatility, randomVariableFactory=None):
        f = randomVariableFactory if randomVariableFactory is not None else RandomVariableCudaFactory()
        self.randomVariableFactory = f
        # numbers or RandomVariables (BlackScholesModel.java:59-81): with a differentiable factory the three parameters are independents
        asRV = lambda v: v if isinstance(v, RandomVariable) else f.createRandomVariable(v)
        self.initialValue, self.riskFreeRate, self.volatility = asRV(initialValue), asRV(riskFreeRate), asRV(volatility)
        self.initialState = [self.initialValue.log()]                                          # :76
        self.drift = [self.riskFreeRate.sub(self.volatility.squared().div(2))]                 # :77
        self.factorLoadings = [self.volatility]

    def getNumberOfComponents(self): return 1
    def getNumberOfFactors(self): return 1
    def getInitialValue(self): return [self.initialValue]
    def getRiskFreeRate(self): return self.riskFreeRate
    def getVolatility(self): return self.volatility
    def getInitialState(self, process): return self.initialState
    def getDrift(self, process, timeIndex, realizationAtTimeIndex, realizationPredictor): return self.drift
    def getFactorLoading(self, process, timeIndex, component, realizationAtTimeIndex): return self.factorLoadings
    def applyStateSpaceTransform(self, process, timeIndex, componentIndex, rv): return rv.exp()
    def applyStateSpaceTransformInverse(self, process, timeIndex, componentIndex, rv): return rv.log()
    def getNumeraire(self, process, time): return self.riskFreeRate.mult(time).exp()
    def getRandomVariableForConstant(self, value): return self.randomVariableFactory.createRandomVariable(value)

    def getReferenceDate(self): return None

    def getCloneWithModifiedData(self, dataModified):        # BlackScholesModel.java:157-166: numbers, defaults = the averages of the current parameters
        d = dict(dataModified or {})
        pick = lambda key, current: float(d[key]) if d.get(key) is not None else current.getAverage()
        return BlackScholesModel(pick("initialValue", self.initialValue), pick("riskFreeRate", self.riskFreeRate), pick("volatility", self.volatility),
                                 self.randomVariableFactory)

    def getFusedSpecification(self, process):
        s0 = self.initialValue.doubleValue()
        return dict(kernel="black_scholes", initialValue=s0, riskFreeRate=self.riskFreeRate.doubleValue(),
                    volatility=self.volatility.doubleValue(), initialValues=[math.exp(math.log(s0))])


class HestonModel:
    REFLECTION, FULL_TRUNCATION = 0, 1

    def __init__(self, initialValue, riskFreeRate, volatility, discountRate, theta, kappa, xi, rho, scheme, randomVariableFactory=None):
        f = randomVariableFactory if randomVariableFactory is not None else RandomVariableCudaFactory()
        self.randomVariableFactory = f
        c = f.createRandomVariable
        self.initialValue, self.riskFreeRate, self.volatility, self.discountRate = c(initialValue), c(riskFreeRate), c(volatility), c(discountRate)
        self.theta, self.kappa, self.xi, self.rho = c(theta), c(kappa), c(xi), c(rho)
        self.rhoBar = self.rho.squared().sub(1).mult(-1).sqrt()                                # :182
        self.scheme = scheme
        self.ZERO = Scalar(0.0)                                                                 # :92

    def getNumberOfComponents(self): return 2
    def getNumberOfFactors(self): return 1                                                       # :437-440 (sic)
    def getInitialState(self, process): return [self.initialValue.log(), self.volatility.squared()]

    def _variance(self, rv):
        return rv.floor(0.0) if self.scheme == HestonModel.FULL_TRUNCATION else rv.abs()

    def getDrift(self, process, timeIndex, x, predictor):
        var = self._variance(x[1])
        return [self.riskFreeRate.sub(var.div(2.0)), self.theta.sub(var).mult(self.kappa)]     # :361-362

    def getFactorLoading(self, process, timeIndex, component, x):
        vol = self._variance(x[1]).sqrt()
        if component == 0:
            return [vol, self.ZERO]
        v = vol.mult(self.xi)
        return [v.mult(self.rho), v.mult(self.rhoBar)]

    def applyStateSpaceTransform(self, process, timeIndex, componentIndex, rv): return rv.exp() if componentIndex == 0 else rv
    def applyStateSpaceTransformInverse(self, process, timeIndex, componentIndex, rv): return rv.log() if componentIndex == 0 else rv
    def getNumeraire(self, process, time): return self.discountRate.mult(time).exp()
    def getRandomVariableForConstant(self, value): return self.randomVariableFactory.createRandomVariable(value)

    def getReferenceDate(self): return None
    def getInitialValue(self): return self.initialValue      # HestonModel.java:474-530
    def getRiskFreeRate(self): return self.riskFreeRate
    def getVolatility(self): return self.volatility
    def getTheta(self): return self.theta
    def getKappa(self): return self.kappa
    def getXi(self): return self.xi
    def getRho(self): return self.rho
    def getScheme(self): return self.scheme

    def getCloneWithModifiedData(self, dataModified):        # HestonModel.java:448-465 (parameters are deterministic here: numbers or Scalars)
        d = dict(dataModified or {})
        factory = d.get("randomVariableFactory", self.randomVariableFactory)

        def pick(key, current):
            v = d.get(key)
            if v is None:
                return current.doubleValue()
            return v.doubleValue() if isinstance(v, RandomVariable) else float(v)
        return HestonModel(pick("initialValue", self.initialValue), pick("riskFreeRate", self.riskFreeRate), pick("volatility", self.volatility),
                           pick("discountRate", self.discountRate), pick("theta", self.theta), pick("kappa", self.kappa), pick("xi", self.xi),
                           pick("rho", self.rho), self.scheme, factory)

    def getFusedSpecification(self, process):
        T = process.getTimeDiscretization().getNumberOfTimeSteps()
        s0, sigma = self.initialValue.doubleValue(), self.volatility.doubleValue()
        return dict(kernel="heston", hestonScheme=self.scheme, initialValue=s0, riskFreeRates=[self.riskFreeRate.doubleValue()] * T,
                    volatility=sigma, theta=self.theta.doubleValue(), kappa=self.kappa.doubleValue(), xi=self.xi.doubleValue(),
                    rho=self.rho.doubleValue(), initialValues=[math.exp(math.log(s0)), sigma * sigma])


class MonteCarloAssetModel:
    def __init__(self, model, process_or_driver, scheme=None):
        self.model = model
        if isinstance(process_or_driver, EulerSchemeFromProcessModel):
            self.process = process_or_driver
        else:
            self.process = EulerSchemeFromProcessModel(model, process_or_driver, scheme)      # MonteCarloAssetModel.java:54-56

    def getTimeDiscretization(self): return self.process.getTimeDiscretization()
    def getTime(self, i): return self.process.getTime(i)
    def getTimeIndex(self, t): return self.process.getTimeIndex(t)
    def getNumberOfPaths(self): return self.process.getNumberOfPaths()
    def getNumberOfAssets(self): return 1
    def getModel(self): return self.model
    def getProcess(self): return self.process
    def getRandomVariableForConstant(self, v): return self.model.getRandomVariableForConstant(v)

    def getAssetValue(self, time, assetIndex):
        if isinstance(time, float):
            timeIndex = self.getTimeIndex(time)
            if timeIndex < 0:
                raise ValueError("The model does not provide an interpolation of simulation time (time given was %r)." % time)
        else:
            timeIndex = time
        return self.process.getProcessValue(timeIndex, assetIndex)

    def getNumeraire(self, time):
        if not isinstance(time, float):
            time = self.getTime(time)
        return self.model.getNumeraire(self.process, time)

    def getMonteCarloWeights(self, time):
        timeIndex = self.getTimeIndex(time) if isinstance(time, float) else time
        return self.process.getMonteCarloWeights(timeIndex)

    def getReferenceDate(self): return self.model.getReferenceDate()

    def getCloneWithModifiedData(self, dataModified):        # MonteCarloAssetModel.java:116-138
        d = dict(dataModified or {})
        newModel = self.model.getCloneWithModifiedData(d)
        forProcess = dict(d)
        forProcess.setdefault("model", newModel)
        return MonteCarloAssetModel(newModel, self.process.getCloneWithModifiedData(forProcess))

    def getCloneWithModifiedSeed(self, seed):                # :146-150
        raise NotImplementedError("Method not implemented")


class MonteCarloBlackScholesModel(MonteCarloAssetModel):
    seed = 3141                                                                                 # MonteCarloBlackScholesModel.java:50

    def __init__(self, *args, shard=None):
        if len(args) == 5:                                   # (timeDiscretization, numberOfPaths, initialValue, riskFreeRate, volatility) :77-87
            td, paths, s0, r, sigma = args
            driver = BrownianMotionCuda(td, 1, paths, self.seed, shard=shard)
        else:                                                # (initialValue, riskFreeRate, volatility, brownianMotion)
            s0, r, sigma, driver = args
        super().__init__(BlackScholesModel(s0, r, sigma, driver.randomVariableFactory), driver)

    def _fresh_driver(self, seed):
        old = self.process.getStochasticDriver()
        return BrownianMotionCuda(self.getTimeDiscretization(), 1, self.getNumberOfPaths(), seed, getattr(old, "randomVariableFactory", None))

    def _with(self, model, driver):
        clone = MonteCarloBlackScholesModel.__new__(MonteCarloBlackScholesModel)
        MonteCarloAssetModel.__init__(clone, model, driver)
        return clone

    def getCloneWithModifiedData(self, dataModified):
        """MonteCarloBlackScholesModel.java:112-148: a new model from the map, on a NEW Brownian motion with this class's seed (the
        reference builds a seeded / time-shifted driver first and then does not use it, :146: mirrored as it behaves)."""
        return self._with(self.model.getCloneWithModifiedData(dataModified), self._fresh_driver(self.seed))

    def getCloneWithModifiedSeed(self, seed):                # :151-155
        return self._with(self.model, self._fresh_driver(int(seed)))


# ---- LIBOR market model ------------------------------------------------------------------------------------------------
class LIBORVolatilityModelFourParameterExponentialForm:
    def __init__(self, timeDiscretization, liborPeriodDiscretization, a, b, c, d, isCalibrateable=False):
        self.td, self.tenor, self.a, self.b, self.c, self.d = timeDiscretization, liborPeriodDiscretization, a, b, c, d
        self.isCalibrateable = isCalibrateable

    def getTimeDiscretization(self): return self.td
    def getLiborPeriodDiscretization(self): return self.tenor

    def getVolatility(self, timeIndex, liborIndex):          # :168-190
        ttm = self.tenor.getTime(liborIndex) - self.td.getTime(timeIndex)
        if ttm <= 0:
            return 0.0
        return (self.b * ttm + self.a) * _jexp(self.c * (-ttm)) + self.d

    def getVolatilityTable(self):
        """[T][N] of getVolatility (the same expression, evaluated for the whole grid at once: a calibration builds it per evaluation)."""
        t = np.asarray(self.td.getAsDoubleArray(), dtype=np.float64)[:self.td.getNumberOfTimeSteps()]
        ttm = np.asarray(self.tenor.getAsDoubleArray(), dtype=np.float64)[None, :self.tenor.getNumberOfTimeSteps()] - t[:, None]
        with np.errstate(over="ignore", invalid="ignore"):
            vol = (self.b * ttm + self.a) * _exp_like_libm(self.c * (-ttm)) + self.d
        return np.where(ttm <= 0, 0.0, vol)

    # parametric interface (:134-166): null unless calibrateable
    def getParameterAsDouble(self):
        return [self.a, self.b, self.c, self.d] if self.isCalibrateable else None

    def getCloneWithModifiedParameter(self, parameter):
        if not self.isCalibrateable:
            return self
        a, b, c, d = (float(v.doubleValue()) if hasattr(v, "doubleValue") else float(v) for v in parameter[:4])
        return LIBORVolatilityModelFourParameterExponentialForm(self.td, self.tenor, a, b, c, d, True)

    def getParameter(self):                                  # :134-150: the parameters as (deterministic) RandomVariables, null unless calibrateable
        p = self.getParameterAsDouble()
        return None if p is None else [Scalar(v) for v in p]

    def clone(self):
        return self.getCloneWithModifiedData(None)

    def getCloneWithModifiedData(self, dataModified):        # :212-253: timeDiscretization, liborPeriodDiscretization, isCalibrateable, a, b, c, d
        m = dict(dataModified or {})
        num = lambda v: float(v.doubleValue()) if hasattr(v, "doubleValue") else float(v)
        return LIBORVolatilityModelFourParameterExponentialForm(m.get("timeDiscretization", self.td), m.get("liborPeriodDiscretization", self.tenor),
                                                                num(m.get("a", self.a)), num(m.get("b", self.b)), num(m.get("c", self.c)), num(m.get("d", self.d)),
                                                                bool(m.get("isCalibrateable", self.isCalibrateable)))


def _exp_like_libm(x):
    # through math.exp, so that the table equals getVolatility() bit for bit (numpy's vectorised exp may differ by an ulp); once per
    # DISTINCT argument: on a regular grid a [T][N] table of times to maturity holds T + N distinct values, not T * N
    x = np.asarray(x, dtype=np.float64)
    u, inverse = np.unique(x.ravel(), return_inverse=True)
    e = np.array([_jexp(v) for v in u.tolist()], dtype=np.float64)
    return e[inverse].reshape(x.shape)


def _factor_matrix(correlation, numberOfFactors):
    """LinearAlgebra.getFactorMatrixUsingCommonsMath :392-443 (symmetric eigen-decomposition, largest eigenvalues first,
    first entry of each eigenvector positive, column scaled by sqrt(eigenvalue / |v|^2))."""
    ev, V = np.linalg.eigh(correlation)
    order = np.argsort(-ev, kind="stable")
    n = correlation.shape[0]
    fm = np.zeros((n, numberOfFactors))
    for f in range(numberOfFactors):
        idx = order[f]
        sign = 1.0 if V[0, idx] > 0.0 else -1.0
        norm2 = float(np.sum(V[:, idx] * V[:, idx]))
        e = max(float(ev[idx]), 0.0)
        fm[:, f] = sign * math.sqrt(e / norm2) * V[:, idx]
    return fm


def factorReduction(correlation, numberOfFactors):           # LinearAlgebra.factorReductionUsingCommonsMath :452-482
    fm = _factor_matrix(correlation, numberOfFactors)
    for row in range(fm.shape[0]):
        s = float(np.sum(fm[row] * fm[row]))
        fm[row] = fm[row] / math.sqrt(s) if s != 0 else 1.0
    return _factor_matrix(fm @ fm.T, numberOfFactors)


class LIBORCorrelationModelExponentialDecay:
    def __init__(self, timeDiscretization, liborPeriodDiscretization, numberOfFactors, a, isCalibrateable=False):
        self.td, self.tenor, self.numberOfFactors, self.a, self.isCalibrateable = timeDiscretization, liborPeriodDiscretization, numberOfFactors, a, isCalibrateable
        a = max(a, 0)                                                                               # :99 (the stored parameter stays as given)
        n = liborPeriodDiscretization.getNumberOfTimeSteps()
        T = liborPeriodDiscretization.getAsDoubleArray()
        corr = np.array([[math.exp(-a * abs(T[r] - T[c])) for c in range(n)] for r in range(n)])     # :102-112
        self.factorMatrix = factorReduction(corr, numberOfFactors)
        self.correlationMatrix = self.factorMatrix @ self.factorMatrix.T
        np.fill_diagonal(self.correlationMatrix, 1.0)                                               # :129

    def getNumberOfFactors(self): return self.factorMatrix.shape[1]
    def getFactorLoading(self, timeIndex, factor, component): return float(self.factorMatrix[component, factor])
    def getCorrelation(self, timeIndex, c1, c2): return float(self.correlationMatrix[c1, c2])
    def getTimeDiscretization(self): return self.td
    def getLiborPeriodDiscretization(self): return self.tenor

    # parametric interface (:74-80, :137-147)
    def getParameterAsDouble(self):
        return [self.a] if self.isCalibrateable else None

    def getCloneWithModifiedParameter(self, parameter):
        a = float(parameter[0].doubleValue()) if hasattr(parameter[0], "doubleValue") else float(parameter[0])
        if not self.isCalibrateable or self.a == a:
            return self
        # (the reference's clone drops the flag, :79 — a one-shot quirk that does not matter there because calibration always clones
        # from the ORIGINAL model; kept calibrateable here so that a calibrated model can be calibrated again)
        return LIBORCorrelationModelExponentialDecay(self.td, self.tenor, self.numberOfFactors, a, True)

    def getParameter(self):                                  # :137-147
        p = self.getParameterAsDouble()
        return None if p is None else [Scalar(v) for v in p]

    def clone(self):
        return self.getCloneWithModifiedData(None)

    def getCloneWithModifiedData(self, dataModified):        # :150-167: timeDiscretization, liborPeriodDiscretization, numberOfFactors, a, isCalibrateable
        m = dict(dataModified or {})
        return LIBORCorrelationModelExponentialDecay(m.get("timeDiscretization", self.td), m.get("liborPeriodDiscretization", self.tenor),
                                                     int(m.get("numberOfFactors", self.numberOfFactors)), float(m.get("a", self.a)),
                                                     bool(m.get("isCalibrateable", self.isCalibrateable)))


class LIBORCovarianceModelFromVolatilityAndCorrelation:
    def __init__(self, timeDiscretization, liborPeriodDiscretization, volatilityModel, correlationModel):
        self.td, self.tenor, self.volatilityModel, self.correlationModel = timeDiscretization, liborPeriodDiscretization, volatilityModel, correlationModel

    def getNumberOfFactors(self): return self.correlationModel.getNumberOfFactors()
    def getTimeDiscretization(self): return self.td
    def getLiborPeriodDiscretization(self): return self.tenor
    def getVolatilityModel(self): return self.volatilityModel
    def getCorrelationModel(self): return self.correlationModel

    def getFactorLoadingTable(self):
        """[T][N][F] deterministic factor loadings sigma_j(t_i) * F[j][k] (:47-57) and [T][N] variances sigma*sigma*corr_jj (:82-93)."""
        T, N, F = self.td.getNumberOfTimeSteps(), self.tenor.getNumberOfTimeSteps(), self.getNumberOfFactors()
        table = getattr(self.volatilityModel, "getVolatilityTable", None)
        fm, cm = getattr(self.correlationModel, "factorMatrix", None), getattr(self.correlationModel, "correlationMatrix", None)
        if table is not None and fm is not None and cm is not None:
            vol = table()                                    # the same products, whole grid at once (bit-identical to the loop below)
            return vol[:, :, None] * fm[None, :N, :F], (vol * vol) * np.diagonal(cm)[None, :N]
        fl = np.zeros((T, N, F))
        var = np.zeros((T, N))
        for t in range(T):
            for j in range(N):
                vol = self.volatilityModel.getVolatility(t, j)
                for k in range(F):
                    fl[t, j, k] = vol * self.correlationModel.getFactorLoading(t, k, j)
                var[t, j] = (vol * vol) * self.correlationModel.getCorrelation(t, j, j)
        return fl, var

    # ---- point-wise accessors of the LIBORCovarianceModel interface (deterministic here: Scalars) -------------------------------------
    @staticmethod
    def _is_index(x):
        return isinstance(x, (int, np.integer)) and not isinstance(x, bool)

    def _time_index(self, time):                             # AbstractLIBORCovarianceModel.java:54-60, :69-76
        ti = self.td.getTimeIndex(time)
        return ti if ti >= 0 else -ti - 2

    def getFactorLoading(self, time, component, realizationAtTimeIndex=None):
        """:47-57 for (timeIndex, componentIndex); AbstractLIBORCovarianceModel.java:45-60 for the overloads taking a time and / or a
        fixing date (Java picks them by int / double; so does this, by the Python type of the argument)."""
        if not self._is_index(component):
            ci = self.tenor.getTimeIndex(component)
            component = ci if ci >= 0 else -ci - 2
        timeIndex = time if self._is_index(time) else self._time_index(time)
        volatility = Scalar(self.volatilityModel.getVolatility(timeIndex, component))
        return [volatility.mult(self.correlationModel.getFactorLoading(timeIndex, f, component)) for f in range(self.getNumberOfFactors())]

    def getFactorLoadingPseudoInverse(self, timeIndex, component, factor, realizationAtTimeIndex=None):       # :60-77
        inverse = Scalar(self.volatilityModel.getVolatility(timeIndex, component)).invert().mult(self.correlationModel.getFactorLoading(timeIndex, factor, component))
        factorWeight = 0.0
        for c in range(self.tenor.getNumberOfTimeSteps()):
            e = self.correlationModel.getFactorLoading(timeIndex, factor, c)
            factorWeight += e * e
        return inverse.mult(1 / factorWeight)

    def getCovariance(self, time, component1, component2, realizationAtTimeIndex=None):                        # :82-93
        timeIndex = time if self._is_index(time) else self._time_index(time)
        v1, v2 = Scalar(self.volatilityModel.getVolatility(timeIndex, component1)), Scalar(self.volatilityModel.getVolatility(timeIndex, component2))
        return v1.mult(v2).mult(self.correlationModel.getCorrelation(timeIndex, component1, component2))

    # ---- parametric interface (LIBORCovarianceModelFromVolatilityAndCorrelation.java:100-160): volatility parameters, then correlation's
    def getParameterAsDouble(self):
        v, c = self.volatilityModel.getParameterAsDouble(), self.correlationModel.getParameterAsDouble()
        return list(v or []) + list(c or [])

    def getCloneWithModifiedParameters(self, parameters):
        v, c = self.volatilityModel.getParameterAsDouble(), self.correlationModel.getParameterAsDouble()
        nv_ = len(v) if v is not None else 0
        vol, corr = self.volatilityModel, self.correlationModel
        if v is not None:
            vol = vol.getCloneWithModifiedParameter(list(parameters[:nv_]))
        if c is not None:
            corr = corr.getCloneWithModifiedParameter(list(parameters[nv_:nv_ + len(c)]))
        return LIBORCovarianceModelFromVolatilityAndCorrelation(self.td, self.tenor, vol, corr)

    def getParameter(self):                                  # :95-115: the volatility model's parameters, then the correlation model's
        return [Scalar(v) for v in self.getParameterAsDouble()]

    def clone(self):
        return self.getCloneWithModifiedData(None)

    def getCloneWithModifiedData(self, dataModified):        # :180-208
        m = dict(dataModified or {})
        vol, corr = self.volatilityModel, self.correlationModel
        if "timeDiscretization" in m or "liborPeriodDiscretization" in m or "randomVariableFactory" in m:
            if "volatilityModel" not in m:
                vol = vol.getCloneWithModifiedData(m)
            if "correlationModel" not in m:
                corr = corr.getCloneWithModifiedData(m)
        return LIBORCovarianceModelFromVolatilityAndCorrelation(m.get("timeDiscretization", self.td), m.get("liborPeriodDiscretization", self.tenor),
                                                                m.get("volatilityModel", vol), m.get("correlationModel", corr))

    def getCloneCalibrated(self, calibrationModel, calibrationProducts, calibrationParameters=None):
        """AbstractLIBORCovarianceModelParametric.java:134-157 -> calibration.getCloneCalibrated."""
        from .calibration import getCloneCalibrated
        return getCloneCalibrated(self, calibrationModel, calibrationProducts, calibrationParameters)


def _accrue_chain(libors, subs, divisor):
    """((1 + L_a d_a) ... (1 + L_{b-1} d_{b-1}) - 1) / divisor in one kernel (fmb_rv_accrue_chain) when every rate is a device vector of
    the same shard; the same operations in the same order as the loop of accrue() calls it replaces (bit-identical).  None otherwise."""
    import ctypes as C
    first = libors[0]
    if len(libors) < 3 or not all(type(l) is RandomVariableCuda and l.dv is not None and l.dv.n == first.dv.n and l.shard is first.shard for l in libors):
        return None
    hs = np.array([l.dv.h for l in libors], dtype=np.uint64)
    ds = np.array(subs, dtype=np.float64)
    out = C.c_uint64()
    nv.check(nv.load().fmb_rv_accrue_chain(len(libors), nv.hptr(hs), nv.dptr(ds), float(divisor), C.byref(out)))
    return RandomVariableCuda(max(l.time for l in libors), None, first.shard, _dv=nv.DeviceVector(out.value, first.dv.n), _n=first.nGlobal)


class LIBORMarketModelFromCovarianceModel:
    SPOT, TERMINAL = 0, 1
    NORMAL, LOGNORMAL = 0, 1

    def __init__(self, liborPeriodDiscretization, forwardRates, discountFactors, randomVariableFactory, covarianceModel, properties=None,
                 factorLoadingTable=None):
        """forwardRates[j] = L_j(0) on the tenor grid (the forward curve evaluated there); discountFactors[i] = P^d(T_i) or None."""
        properties = dict(properties or {})
        self._properties = properties
        self.tenor = liborPeriodDiscretization
        self.L0 = np.asarray(forwardRates, dtype=np.float64)
        self.discountFactors = None if discountFactors is None else np.asarray(discountFactors, dtype=np.float64)
        self.randomVariableFactory = randomVariableFactory if randomVariableFactory is not None else RandomVariableCudaFactory()
        self.covarianceModel = covarianceModel
        self.measure = {"SPOT": 0, "TERMINAL": 1}[properties.get("measure", "SPOT")]                 # defaults :186-194
        self.stateSpace = {"NORMAL": 0, "LOGNORMAL": 1}[properties.get("stateSpace", "LOGNORMAL")]
        self.liborCap = float(properties.get("liborCap", 1e5))
        self.interpolationMethod = properties.get("interpolationMethod", "LOG_LINEAR_UNCORRECTED")    # :186-194
        self.forwardCurve = properties.get("forwardCurve")   # optional callable time -> forward (host-side curve input); default: L0 by tenor period
        self._tables = factorLoadingTable
        self._numeraires, self._numeraireDiscountFactors, self._numerairesProcess = {}, {}, None
        self._numerairesAdjusted = {}
        self._zeroBondAverages, self._zeroBondRequests, self._zeroBondLastIndex = None, set(), None
        self._initialState = None
        self._periodLengthSlices = {}
        self._forwardCurveDiscountFactors = None
        self._integratedLIBORCovariance = None

    @classmethod
    def of(cls, liborPeriodDiscretization, analyticModel, forwardRates, discountFactors, randomVariableFactory, covarianceModel, calibrationItems=None,
           properties=None):
        model = cls(liborPeriodDiscretization, forwardRates, discountFactors, randomVariableFactory, covarianceModel, properties)
        if calibrationItems:                                 # LIBORMarketModelFromCovarianceModel.java:296-318: calibrate, if data is given
            if not hasattr(covarianceModel, "getCloneCalibrated"):
                raise TypeError("Calibration restricted to covariance models implementing LIBORCovarianceModelCalibrateable.")   # ClassCastException
            calibrated = covarianceModel.getCloneCalibrated(model, calibrationItems, (properties or {}).get("calibrationParameters"))
            return model.getCloneWithModifiedCovarianceModel(calibrated)
        return model

    def getCovarianceModel(self): return self.covarianceModel

    def getCloneWithModifiedCovarianceModel(self, covarianceModel):                     # :1420-1426: same curves, factory and properties
        return LIBORMarketModelFromCovarianceModel(self.tenor, self.L0, self.discountFactors, self.randomVariableFactory, covarianceModel,
                                                   self._properties)

    def getCloneWithModifiedData(self, dataModified):
        """:1653-1696.  Keys: randomVariableFactory, liborPeriodDiscretization, covarianceModel, forwardRateCurve / discountCurve (here the
        host-side inputs they stand for: forward rates L_j(0) / discount factors on the tenor grid), forwardRateShift (added to the
        forward rates); swaptionMarketData is refused as in the reference.  The clone keeps measure, state space, interpolation, cap."""
        d = dict(dataModified or {})
        if "swaptionMarketData" in d:
            raise RuntimeError("Swaption market data as input for getCloneWithModifiedData not supported.")
        forwardRates = np.asarray(d.get("forwardRateCurve", self.L0), dtype=np.float64)
        if "forwardRateShift" in d:
            forwardRates = forwardRates + np.asarray(d["forwardRateShift"], dtype=np.float64)
        properties = dict(self._properties)
        properties.pop("calibrationParameters", None)
        return LIBORMarketModelFromCovarianceModel.of(d.get("liborPeriodDiscretization", self.tenor), d.get("analyticModel"), forwardRates,
                                                      d.get("discountCurve", self.discountFactors), d.get("randomVariableFactory", self.randomVariableFactory),
                                                      d.get("covarianceModel", self.covarianceModel), None, properties)

    def getReferenceDate(self):                              # :1412-1414: the forward curve's reference date; curves are arrays here
        return None

    def getMeasure(self): return self.measure                # :1530-1550
    def getInterpolationMethod(self): return self.interpolationMethod
    def getSwaptionMarketData(self): return None
    def clone(self): return self.getCloneWithModifiedData(None)

    def getModelParameters(self):
        """:1699-1733: name -> RandomVariable for the initial forward rates (as the process cached by the numeraire sees them at time 0),
        the covariance model's parameters and the numeraire adjustments: with a differentiable factory the keys under which a gradient
        is read.  (Times are formatted by Python's repr, which agrees with Java's Double.toString for the usual grid values.)"""
        process = self._numerairesProcess() if self._numerairesProcess is not None else None
        parameters = {}
        for i in range(self.tenor.getNumberOfTimeSteps()):
            forward = self.getLIBOR(process, 0, i) if process is not None else None
            parameters["FORWARD(%r,%r)" % (self.getLiborPeriod(i), self.getLiborPeriod(i + 1))] = forward
        getParameter = getattr(self.covarianceModel, "getParameter", None)
        if getParameter is not None:
            for i, p in enumerate(getParameter() or []):
                parameters["COVARIANCEMODELPARAMETER(%d)" % i] = p
        for t, adjustment in self.getNumeraireAdjustments().items():
            parameters["NUMERAIREADJUSTMENT(%r)" % t] = adjustment
        return dict(sorted(parameters.items()))             # a TreeMap there

    def getNumeraireAdjustments(self):
        """:1504-1506: tenor time -> forward rate of the discount curve over the period starting there (the numeraire adjustment's
        building blocks); empty until a numeraire has been asked for, like the reference's lazily filled map."""
        if self.discountFactors is None or not self._numeraireDiscountFactors:
            return {}
        c = self.randomVariableFactory.createRandomVariable
        return {self.tenor.getTime(i): c((float(self.discountFactors[i]) / float(self.discountFactors[i + 1]) - 1.0) / self.tenor.getTimeStep(i))
                for i in range(self.tenor.getNumberOfTimeSteps())}

    def getForwardDiscountBond(self, process, time, maturity):                          # :947-952
        if self.discountFactors is None:
            raise ValueError("getForwardDiscountBond needs a discount curve (the reference dereferences it)")
        inverseForwardBondAsOfTime = self.getForwardRate(process, time, time, maturity).mult(maturity - time).add(1.0)
        inverseForwardBondAsOfZero = self.getForwardRate(process, 0.0, time, maturity).mult(maturity - time).add(1.0)
        forwardDiscountBondAsOfZero = self._defaultable_zero_bond_at(process, maturity).div(self._defaultable_zero_bond_at(process, time))
        return forwardDiscountBondAsOfZero.mult(inverseForwardBondAsOfZero).div(inverseForwardBondAsOfTime)

    def getIntegratedLIBORCovariance(self, simulationTimeDiscretization):
        """:1552-1596: [t][i][j] = sum over s <= t of sum_f FL_i,f(s) FL_j,f(s) dt_s for rates not yet fixed at s (zero otherwise), the
        same products in the same order; computed once per model like the reference's lazy field.  (As there, the lower triangle of the
        first time index is left at zero: the symmetrisation runs inside the time integration, which starts at index 1.)"""
        if self._integratedLIBORCovariance is None:
            td, N = simulationTimeDiscretization, self.tenor.getNumberOfTimeSteps()
            T = td.getNumberOfTimeSteps()
            out = np.zeros((T, N, N))
            upper = np.triu(np.ones((N, N), dtype=bool))
            for t in range(T):
                dt = td.getTime(t + 1) - td.getTime(t)
                fl = np.array([[f.doubleValue() for f in self.covarianceModel.getFactorLoading(td.getTime(t), self.tenor.getTime(c), None)] for c in range(N)])
                acc = np.zeros((N, N))
                for f in range(fl.shape[1]):
                    acc = acc + fl[:, f][:, None] * fl[:, f][None, :] * dt
                live = np.array([self.getLiborPeriod(c) > td.getTime(t) for c in range(N)])
                out[t] = np.where(upper & live[:, None], acc, 0.0)
            for t in range(1, T):
                summed = out[t - 1] + out[t]
                out[t] = np.where(upper, summed, 0.0)
                out[t] = out[t] + np.triu(out[t], 1).T
            self._integratedLIBORCovariance = out
        return self._integratedLIBORCovariance

    # ---- ProcessModel callbacks -----------------------------------------------------------------------------------------
    def getNumberOfComponents(self): return self.tenor.getNumberOfTimeSteps()
    def getNumberOfLibors(self): return self.getNumberOfComponents()
    def getNumberOfFactors(self): return self.covarianceModel.getNumberOfFactors()
    def getLiborPeriodDiscretization(self): return self.tenor
    def getLiborPeriod(self, i): return self.tenor.getTime(i)
    def getLiborPeriodIndex(self, time): return self.tenor.getTimeIndex(time)
    def getRandomVariableForConstant(self, v): return self.randomVariableFactory.createRandomVariable(v)

    def _tables_for(self, process):
        if self._tables is None:
            self._tables = self.covarianceModel.getFactorLoadingTable()
        return self._tables

    def getDiscountFactorsFromForwardCurve(self):
        """DiscountCurveFromForwardCurve(forwardRateCurve).getDiscountFactor(T_i) on the tenor grid (:130-142): running product of
        1 / (1 + L_i(0) * (T_{i+1} - T_i)), starting from 1."""
        df = self._forwardCurveDiscountFactors               # (the curve and the tenor grid are fixed at construction; a swaption asks per period)
        if df is None:
            df = [1.0]
            for i in range(self.tenor.getNumberOfTimeSteps()):
                df.append(df[-1] / (1.0 + float(self.L0[i]) * self.tenor.getTimeStep(i)))
            self._forwardCurveDiscountFactors = df
        return df

    def getInitialState(self, process):                      # :1080-1093
        # created once per model: with a differentiable factory these are the independents of the forward-rate sensitivities
        if self._initialState is None:
            c = self.getRandomVariableForConstant
            if self.stateSpace == self.LOGNORMAL:
                self._initialState = [c(math.log(max(r, 0.0)) if max(r, 0.0) > 0 else float("-inf")) for r in self.L0]
            else:
                self._initialState = [c(float(r)) for r in self.L0]
        return list(self._initialState)

    def _first(self, time):
        first = self.getLiborPeriodIndex(time) + 1          # :1127-1130
        if first < 0:
            first = -first - 1 + 1
        return first

    def _sim_index(self, process, time):
        # row of the covariance model's tables for a simulation time: ITS time discretization (AbstractLIBORCovarianceModel.java:70-76)
        td = getattr(self.covarianceModel, "td", None)
        ti = td.getTimeIndex(time) if td is not None else process.getTimeIndex(time)
        return ti if ti >= 0 else -ti - 2

    def getFactorLoading(self, process, timeIndex, componentIndex, realizationAtTimeIndex):
        fl, _ = self._tables_for(process)
        ti = self._sim_index(process, process.getTime(timeIndex))
        return [Scalar(fl[ti, componentIndex, k]) for k in range(fl.shape[2])]

    def getDrift(self, process, timeIndex, x, predictor):    # :1124-1191 on RandomVariables (used by the generic Euler loop)
        time = process.getTime(timeIndex)
        first = self._first(time)
        N, F = self.getNumberOfComponents(), self.getNumberOfFactors()
        fl, var = self._tables_for(process)
        ti = self._sim_index(process, time)
        zero = Scalar(0.0)
        drift = [None] * N
        for c in range(first, N):
            drift[c] = zero
        sums = [zero] * F
        order = range(first, N) if self.measure == self.SPOT else range(N - 1, first - 1, -1)
        for c in order:
            pl = self.tenor.getTimeStep(c)
            fr = x[c]
            ost = Scalar(pl if self.measure == self.SPOT else -pl).discount(fr, pl)
            if self.stateSpace == self.LOGNORMAL:
                ost = ost.mult(fr)
            flc = [Scalar(fl[ti, c, k]) for k in range(F)]
            if self.measure == self.SPOT:
                sums = [sums[k].addProduct(ost, flc[k]) for k in range(F)]
                drift[c] = drift[c].addSumProduct(sums, flc)
            else:
                drift[c] = drift[c].addSumProduct(sums, flc)
                sums = [sums[k].addProduct(ost, flc[k]) for k in range(F)]
        if self.stateSpace == self.LOGNORMAL:
            for c in range(first, N):
                drift[c] = drift[c].addProduct(Scalar(var[ti, c]), -0.5)
        return drift

    def applyStateSpaceTransform(self, process, timeIndex, componentIndex, rv):      # :1199-1212
        v = rv
        if self.stateSpace == self.LOGNORMAL:
            v = v.exp()
        if not math.isinf(self.liborCap):
            v = v.cap(self.liborCap)
        return v

    def applyStateSpaceTransformInverse(self, process, timeIndex, componentIndex, rv):
        return rv.log() if self.stateSpace == self.LOGNORMAL else rv

    def getFusedSpecification(self, process):
        td = process.getTimeDiscretization()
        T, N = td.getNumberOfTimeSteps(), self.getNumberOfComponents()
        fl, var = self._tables_for(process)
        fl, var = np.asarray(fl, dtype=np.float64), np.asarray(var, dtype=np.float64)
        F = process.getNumberOfFactors()
        if fl.ndim != 3 or fl.shape[1] != N or fl.shape[2] != F or var.shape != fl.shape[:2]:
            return None                                      # driver / covariance model disagree on factors or components: generic loop (it raises or handles it per call like the reference)
        # the covariance model has its own time discretization: row of process step t = its index of the step's start time
        # (AbstractLIBORCovarianceModel.java:70-76: getTimeIndex(time), negative -> |index| - 2)
        ctd = getattr(self.covarianceModel, "td", None)
        if ctd is not None and not (ctd.getNumberOfTimeSteps() == T and all(ctd.getTime(t) == td.getTime(t) for t in range(T))):
            rows = []
            for t in range(T):
                ci = ctd.getTimeIndex(td.getTime(t))
                if ci < 0:
                    ci = -ci - 2
                if ci < 0 or ci >= fl.shape[0]:
                    return None
                rows.append(ci)
            fl, var = fl[rows], var[rows]
        elif fl.shape[0] != T:
            return None
        y0 = [s.doubleValue() for s in self.getInitialState(process)]
        x0 = []
        for y in y0:
            x = math.exp(y) if self.stateSpace == self.LOGNORMAL else y
            x0.append(x if math.isinf(self.liborCap) else min(x, self.liborCap))
        return dict(kernel="lmm", measure=self.measure, stateSpace=self.stateSpace, liborCap=self.liborCap, initialState=y0,
                    periodLength=[self.tenor.getTimeStep(j) for j in range(N)], factorLoading=fl, variance=var,
                    firstLive=[self._first(td.getTime(t)) for t in range(T)], initialValues=x0)

    # ---- term-structure functions on RandomVariables (unchanged host logic) -------------------------------------------------
    def getLIBOR(self, process, timeIndex, liborIndex):
        return process.getProcessValue(timeIndex, liborIndex)

    def getForwardRateCurveForward(self, time):
        """getForwardRateCurve().getForward(model, time, paymentOffset) (:1374-1375).  The forward curve is a host-side input given on the
        tenor grid (the curve classes are outside the path): between grid points the forward of the tenor period containing `time` is
        used — for a flat curve, every configuration here, exactly what the reference's ForwardCurveInterpolation returns."""
        if self.forwardCurve is not None:
            return float(self.forwardCurve(time))
        i = self.tenor.getTimeIndex(time)
        if i < 0:
            i = -i - 2
        return float(self.L0[max(0, min(i, self.getNumberOfComponents() - 1))])

    def _onePlusInterpolatedLIBORDt(self, process, timeIndex, periodStartTime, liborPeriodIndex):      # :1324-1395
        tenorPeriodStartTime, tenorPeriodEndTime = self.getLiborPeriod(liborPeriodIndex), self.getLiborPeriod(liborPeriodIndex + 1)
        tenorDt = tenorPeriodEndTime - tenorPeriodStartTime
        if tenorPeriodStartTime < process.getTime(timeIndex):
            timeIndex = min(timeIndex, process.getTimeIndex(tenorPeriodStartTime))       # fixed at the long period's start
            if timeIndex < 0:
                raise ValueError("Tenor discretization not part of time discretization.")
        onePlusLongLIBORDt = self.getLIBOR(process, timeIndex, liborPeriodIndex).mult(tenorDt).add(1.0)
        smallDt = tenorPeriodEndTime - periodStartTime
        alpha = smallDt / tenorDt
        if self.interpolationMethod == "LINEAR":
            onePlusInterpolatedLIBORDt = onePlusLongLIBORDt.mult(alpha).add(1 - alpha)
        elif self.interpolationMethod == "LOG_LINEAR_UNCORRECTED":
            onePlusInterpolatedLIBORDt = onePlusLongLIBORDt.log().mult(alpha).exp()
        else:
            raise NotImplementedError("interpolation method %s (drift-adjusted log-linear interpolation) is outside the hot path" % self.interpolationMethod)
        analyticOnePlusLongLIBORDt = 1 + self.getForwardRateCurveForward(tenorPeriodStartTime) * tenorDt
        analyticOnePlusShortLIBORDt = 1 + self.getForwardRateCurveForward(periodStartTime) * smallDt
        if self.interpolationMethod == "LINEAR":
            analyticOnePlusInterpolatedLIBORDt = analyticOnePlusLongLIBORDt * alpha + (1 - alpha)
        else:
            analyticOnePlusInterpolatedLIBORDt = math.exp(math.log(analyticOnePlusLongLIBORDt) * alpha)
        return onePlusInterpolatedLIBORDt.mult(analyticOnePlusShortLIBORDt / analyticOnePlusInterpolatedLIBORDt)

    def getForwardRate(self, process, time, periodStart, periodEnd):                 # :1237-1305
        ps, pe = self.getLiborPeriodIndex(periodStart), self.getLiborPeriodIndex(periodEnd)
        time = min(time, periodStart)
        ti = process.getTimeIndex(time)
        if ti < 0:
            ti = -ti - 2
            if time - process.getTime(ti) > process.getTime(ti + 1) - time:           # ROUND_NEAREST
                ti += 1
        if pe < 0:                                           # period end between tenor points (:1257-1264)
            previousEndIndex = (-pe - 1) - 1
            nextEndTime = self.getLiborPeriod(previousEndIndex + 1)
            onePlusLongLIBORdt = self.getForwardRate(process, time, periodStart, nextEndTime).mult(nextEndTime - periodStart).add(1.0)
            onePlusInterpolatedLIBORDt = self._onePlusInterpolatedLIBORDt(process, ti, periodEnd, previousEndIndex)
            return onePlusLongLIBORdt.div(onePlusInterpolatedLIBORDt).sub(1.0).div(periodEnd - periodStart)
        if ps < 0:                                           # period start between tenor points (:1267-1279)
            previousStartIndex = (-ps - 1) - 1
            nextStartTime = self.getLiborPeriod(previousStartIndex + 1)
            if nextStartTime > periodEnd:
                raise AssertionError("Interpolation not possible.")
            if nextStartTime == periodEnd:
                return self._onePlusInterpolatedLIBORDt(process, ti, periodStart, previousStartIndex).sub(1.0).div(periodEnd - periodStart)
            onePlusLongLIBORdt = self.getForwardRate(process, time, nextStartTime, periodEnd).mult(periodEnd - nextStartTime).add(1.0)
            onePlusInterpolatedLIBORDt = self._onePlusInterpolatedLIBORDt(process, ti, periodStart, previousStartIndex)
            return onePlusLongLIBORdt.mult(onePlusInterpolatedLIBORDt).sub(1.0).div(periodEnd - periodStart)
        if ps + 1 == pe:
            return self.getLIBOR(process, ti, ps)
        fused = self._forward_rate_from_handles(process, ti, ps, pe, periodEnd - periodStart)
        if fused is not None:
            return fused
        libors = [self.getLIBOR(process, ti, k) for k in range(ps, pe)]
        subs = [self.getLiborPeriod(k + 1) - self.getLiborPeriod(k) for k in range(ps, pe)]
        fused = _accrue_chain(libors, subs, periodEnd - periodStart)
        if fused is not None:
            return fused
        acc = None
        for l, sub in zip(libors, subs):                     # :1288-1302, one pass per period
            acc = l.mult(sub).add(1.0) if acc is None else acc.accrue(l, sub)
        return acc.sub(1.0).div(periodEnd - periodStart)

    def _forward_rate_from_handles(self, process, ti, ps, pe, divisor):
        """The multi-period forward rate of a fused simulation straight from the native handles of L_ps..L_{pe-1} at time index ti (a
        contiguous slice of the handle table): the same kernel on the same vectors as _accrue_chain, without wrapping every rate into a
        RandomVariable first (the Bermudan's basis functions ask for up to 19 rates per call).  None: the caller takes the general way."""
        import ctypes as C
        if pe - ps < 3 or ti == 0:
            return None
        if getattr(process, "_discreteProcess", None) is None:
            process.getProcessValue(ti, ps)                  # (runs the simulation)
        lp = getattr(process, "_discreteProcess", None)
        handles = getattr(lp, "handles", None)
        if handles is None:
            return None
        N = lp.N
        hs = handles[ti * N + ps:ti * N + pe]
        # a deterministic rate (no handle), or every rate frozen before this time index (the result's filtration time is then an earlier one)
        if not hs.all() or not (hs != handles[(ti - 1) * N + ps:(ti - 1) * N + pe]).any():
            return None
        key = (ps, pe)
        ds = self._periodLengthSlices.get(key)
        if ds is None:
            ds = self._periodLengthSlices[key] = np.array([self.getLiborPeriod(k + 1) - self.getLiborPeriod(k) for k in range(ps, pe)], dtype=np.float64)
        out = C.c_uint64()
        nv.check(nv.load().fmb_rv_accrue_chain(pe - ps, nv.hptr(hs), nv.dptr(ds), float(divisor), C.byref(out)))
        return lp.factory.fromDevice(lp.td.getTime(ti), nv.DeviceVector(out.value, lp.P), lp.nPaths)

    def _ensure_cache(self, process):
        # :951-961; a weak reference: the process owns device memory and already references this model (no cycle for the GC to find)
        if self._numerairesProcess is None or self._numerairesProcess() is not process:
            self._numeraires.clear()
            self._numeraireDiscountFactors.clear()
            self._numerairesAdjusted.clear()
            self._zeroBondAverages = None
            self._numerairesProcess = weakref.ref(process)

    def _numeraire_unadjusted_at(self, process, li):                                   # :1017-1074
        self._ensure_cache(process)
        n = self._numeraires.get(li)
        if n is None and self.measure == self.SPOT and li >= 3:
            self._numeraires_by_prefix_accrual(process, li)  # N(T_1) .. N(T_li) from one kernel where possible
            n = self._numeraires.get(li)
        if n is None:
            if self.measure == self.TERMINAL:
                ti = process.getTimeIndex(self.tenor.getTime(li))
                if ti < 0:
                    ti = -ti - 1
                n = self.getRandomVariableForConstant(1.0)
                for k in range(li, self.tenor.getNumberOfTimeSteps()):
                    n = n.discount(self.getLIBOR(process, ti, k), self.tenor.getTimeStep(k))
            else:
                if li != 0:
                    ti = process.getTimeIndex(self.tenor.getTime(li - 1))
                    if ti < 0:
                        ti = -ti - 1
                    n = self._numeraire_unadjusted_at(process, li - 1).accrue(self.getLIBOR(process, ti, li - 1), self.tenor.getTimeStep(li - 1))
                else:
                    n = self.getRandomVariableForConstant(1.0)
            self._numeraires[li] = n
        return n

    def _numeraire_unadjusted(self, process, time):          # :962-1015
        li = self.getLiborPeriodIndex(time)
        if li < 0:                                           # between tenor points (:969-1006)
            upperIndex = -li - 1
            lowerIndex = upperIndex - 1
            if lowerIndex < 0:
                raise ValueError("Numeraire requested for time %r. Unsupported" % time)
            if self.measure == self.TERMINAL:
                n = self.getRandomVariableForConstant(1.0)
                for k in range(upperIndex, self.tenor.getNumberOfTimeSteps()):
                    libor = self.getLIBOR(process, process.getTimeIndex(min(time, self.tenor.getTime(k))), k)
                    n = n.discount(libor, self.tenor.getTimeStep(k))
            else:
                n = self._numeraire_unadjusted(process, self.getLiborPeriod(upperIndex))
            # multiply with the short period bond
            return n.discount(self.getForwardRate(process, time, time, self.getLiborPeriod(upperIndex)), self.getLiborPeriod(upperIndex) - time)
        return self._numeraire_unadjusted_at(process, li)

    def _defaultable_zero_bond_at(self, process, time):      # :886-905: log-linear interpolation of the adjustment between tenor points
        ti = self.tenor.getTimeIndex(time)
        if ti >= 0:
            return self._defaultable_zero_bond(process, ti)
        timeIndexPrev = min(-ti - 2, self.tenor.getNumberOfTimes() - 2)
        timeIndexNext = timeIndexPrev + 1
        timePrev, timeNext = self.tenor.getTime(timeIndexPrev), self.tenor.getTime(timeIndexNext)
        prev, nxt = self._defaultable_zero_bond(process, timeIndexPrev), self._defaultable_zero_bond(process, timeIndexNext)
        return prev.mult(nxt.div(prev).pow((time - timePrev) / (timeNext - timePrev)))

    def _defaultable_zero_bond(self, process, timeIndex):                              # :915-944
        self._ensure_cache(process)
        if not self._numeraireDiscountFactors:
            c = self.randomVariableFactory.createRandomVariable
            adj = c(float(self.discountFactors[0]))
            self._numeraireDiscountFactors[0] = adj
            for i in range(self.tenor.getNumberOfTimeSteps()):
                dfPrev, dfNext, ts = float(self.discountFactors[i]), float(self.discountFactors[i + 1]), self.tenor.getTimeStep(i)
                adj = adj.discount(c((dfPrev / dfNext - 1.0) / ts), ts)
                self._numeraireDiscountFactors[i + 1] = adj
        return self._numeraireDiscountFactors[timeIndex]

    def getNumeraire(self, process, time):                                             # :859-876
        if time < 0:
            raise NotImplementedError("numeraire for negative times is outside the hot path")
        if self.discountFactors is not None:
            # The reference re-evaluates the adjustment (three array passes and a getAverage) on every call; the value is a pure function
            # of (process, time), so the immutable result is kept per time: identical numbers, one reduction (and, sharded, one collective)
            # per distinct date instead of one per call.
            self._ensure_cache(process)
            cached = self._numerairesAdjusted.get(time)
            if cached is not None:
                return cached
        n = self._numeraire_unadjusted(process, time)
        if self.discountFactors is not None:
            dz = self._defaultable_zero_bond_at(process, time)
            nonDefaultableZeroBond = self._zero_bond_averages(process, time)
            if nonDefaultableZeroBond is None:
                nonDefaultableZeroBond = n.invert().mult(self._numeraire_unadjusted(process, 0.0)).getAverage()
            n = n.mult(nonDefaultableZeroBond).div(dz)
            self._numerairesAdjusted[time] = n
        return n

    def _numeraires_by_prefix_accrual(self, process, li):
        """The unadjusted spot-measure numeraires N(T_1) .. N(T_li) that are not cached yet, all from ONE kernel (fmb_rv_accrue_prefix) instead
        of one accrue() pass per date; the same operations in the same order as _numeraire_unadjusted_at (bit-identical).  Does nothing
        where that is not possible (other RandomVariable types, deterministic rates): the per-date route then fills the cache."""
        import ctypes as C
        self._ensure_cache(process)
        first = 1
        while first <= li and first in self._numeraires:
            first += 1
        if first > li:
            return
        start = self._numeraires.get(first - 1) if first > 1 else self._numeraire_unadjusted_at(process, 0)
        if start is None or not start.isDeterministic():
            return                                           # (continuing from a stochastic N would need its vector as the carry: the per-date route)
        while first <= li:                                   # the head of the chain is deterministic (rates fixed at time 0): host scalars
            ti = process.getTimeIndex(self.tenor.getTime(first - 1))
            L = self.getLIBOR(process, ti if ti >= 0 else -ti - 1, first - 1)
            if not L.isDeterministic():
                break
            start = start.accrue(L, self.tenor.getTimeStep(first - 1))
            self._numeraires[first] = start
            first += 1
        if li - first < 2:
            return
        rates, dts, times = [], [], []
        t = start.getFiltrationTime()
        for k in range(first, li + 1):
            ti = process.getTimeIndex(self.tenor.getTime(k - 1))
            if ti < 0:
                ti = -ti - 1
            L = self.getLIBOR(process, ti, k - 1)
            if type(L) is not RandomVariableCuda or L.dv is None or (rates and (L.dv.n != rates[0].dv.n or L.shard is not rates[0].shard)):
                return
            rates.append(L)
            dts.append(self.tenor.getTimeStep(k - 1))
            t = max(t, L.time)
            times.append(t)
        hs = np.array([L.dv.h for L in rates], dtype=np.uint64)
        out = np.zeros(len(rates), dtype=np.uint64)
        nv.check(nv.load().fmb_rv_accrue_prefix(len(rates), start.doubleValue(), nv.hptr(hs), nv.dptr(np.array(dts, dtype=np.float64)), nv.hptr(out)))
        for j, k in enumerate(range(first, li + 1)):
            self._numeraires[k] = RandomVariableCuda(times[j], None, rates[0].shard, _dv=nv.DeviceVector(out[j], rates[0].dv.n), _n=rates[0].nGlobal)

    def _zero_bond_averages(self, process, time):
        """E[N(0) / N(T_k)] for ALL tenor dates up to `time` that have not been averaged yet, in one go.  The reference evaluates them one
        getAverage at a time (one kernel, one host synchronisation and - sharded - one rendezvous of all ranks per date: 20 of the 43 in a
        Bermudan valuation); under the spot measure N(T_k) is a by-product of N(T_i) for every k < i (the accrual chain), so
        fmb_rv_reduce_many sums them all in one launch the first time the latest date is asked for (a backward induction asks for it
        first).  Same operations per path (invert, mult by the deterministic N(0), double-double sum, division by the number of paths).
        Only from the third distinct date on and only for consecutive tenor dates: a product that needs one or two numeraires (a swaption:
        the exercise date) keeps the reference's route, where the accrual chain up to that date is a single fused evaluation and nothing else
        is averaged.
        Returns the average for `time`, or None where this route does not apply (terminal measure, dates between tenor points, other
        RandomVariable types): the caller then takes the reference's."""
        self._ensure_cache(process)
        if self._zeroBondAverages is None:
            self._zeroBondAverages, self._zeroBondRequests, self._zeroBondLastIndex = {}, set(), None
        known = self._zeroBondAverages.get(time)
        li = self.getLiborPeriodIndex(time)
        if known is not None or li < 0 or self.measure != self.SPOT:
            return known
        # a sweep over consecutive tenor dates (a backward induction) is what the batch is for; scattered dates (a portfolio of swaptions
        # with a handful of exercise dates) keep the reference's route
        adjacent = self._zeroBondLastIndex is not None and abs(li - self._zeroBondLastIndex) == 1
        self._zeroBondRequests.add(time)
        self._zeroBondLastIndex = li
        if len(self._zeroBondRequests) < 3 or not adjacent:
            return None
        li = max(li, max(self.getLiborPeriodIndex(t) for t in self._zeroBondRequests))
        n0 = self._numeraire_unadjusted(process, 0.0)
        if not n0.isDeterministic():
            return None
        dates, vectors = [], []
        for k in range(li + 1):
            date = self.tenor.getTime(k)
            if date in self._zeroBondAverages or date in self._numerairesAdjusted:
                continue
            n = self._numeraire_unadjusted_at(process, k)
            if type(n) is not RandomVariableCuda:
                return None
            if n.dv is None:                                  # deterministic (T_0): host scalars, as the type does it
                self._zeroBondAverages[date] = n.invert().mult(n0).getAverage()
                continue
            dates.append(date)
            vectors.append(n)
        if vectors:
            if not all(v.shard is vectors[0].shard and v.dv.n == vectors[0].dv.n for v in vectors):
                return None
            shard, size = vectors[0].shard, vectors[0].size()
            for k in range(0, len(vectors), 64):
                hi, lo = nv.reduce_many(nv.RM_SUM_INVERT_MULT, [v.dv for v in vectors[k:k + 64]], n0.doubleValue())
                hi, lo = shard.sum_dd_many(hi, lo)
                for date, h, l in zip(dates[k:k + 64], hi, lo):
                    self._zeroBondAverages[date] = float("nan") if size == 0 else float(h + l) / size
        return self._zeroBondAverages.get(time)


class LIBORMonteCarloSimulationFromLIBORModel:
    def __init__(self, process_or_model, process=None):
        if process is None:
            self.process, self.model = process_or_model, process_or_model.getModel()
        else:
            self.model, self.process = process_or_model, process

    def getModel(self): return self.model
    def getProcess(self): return self.process
    def getBrownianMotion(self): return self.process.getStochasticDriver()
    def getTimeDiscretization(self): return self.process.getTimeDiscretization()
    def getTime(self, i): return self.process.getTime(i)
    def getTimeIndex(self, t): return self.process.getTimeIndex(t)
    def getNumberOfPaths(self): return self.process.getNumberOfPaths()
    def getNumberOfLibors(self): return self.model.getNumberOfLibors()
    def getLiborPeriodDiscretization(self): return self.model.getLiborPeriodDiscretization()
    def getLiborPeriod(self, i): return self.model.getLiborPeriod(i)
    def getLiborPeriodIndex(self, t): return self.model.getLiborPeriodIndex(t)
    def getLIBOR(self, timeIndex, liborIndex): return self.model.getLIBOR(self.process, timeIndex, liborIndex)
    def getForwardRate(self, time, periodStart, periodEnd): return self.model.getForwardRate(self.process, time, periodStart, periodEnd)
    def getNumeraire(self, time): return self.model.getNumeraire(self.process, time)
    def getRandomVariableForConstant(self, v): return self.model.getRandomVariableForConstant(v)

    def getMonteCarloWeights(self, time):
        ti = self.getTimeIndex(time) if isinstance(time, float) else time
        return self.process.getMonteCarloWeights(ti)

    def getCloneWithModifiedSeed(self, seed):
        return LIBORMonteCarloSimulationFromLIBORModel(self.model, self.process.getCloneWithModifiedSeed(seed))

    def getNumberOfFactors(self): return self.process.getNumberOfFactors()              # LIBORMonteCarloSimulationFromLIBORModel.java:67-69
    def getNumberOfComponents(self): return self.model.getNumberOfComponents()
    def getReferenceDate(self): return self.model.getReferenceDate()                    # :77-79

    def getModelParameters(self): return self.model.getModelParameters()                # :202-205

    def getLIBORs(self, timeIndex):                                                     # :112-120
        return [self.getLIBOR(timeIndex, c) for c in range(self.getNumberOfComponents())]

    def getCloneWithModifiedData(self, dataModified, value=None):                       # :174-200 (map, or key and value)
        d = dict(dataModified) if value is None and not isinstance(dataModified, str) else {dataModified: value}
        modelClone = self.model.getCloneWithModifiedData(d)
        if "discountCurve" in d and len(d) == 1:
            return LIBORMonteCarloSimulationFromLIBORModel(modelClone, self.process)       # the paths do not depend on the discount curve: re-used
        return LIBORMonteCarloSimulationFromLIBORModel(self.process.getCloneWithModifiedModel(modelClone))


# ---- Hull-White ------------------------------------------------------------------------------------------------------------
class ShortRateVolatilityModelAsGiven:
    """J/montecarlo/interestrate/models/covariance/ShortRateVolatilityModelAsGiven.java:20-60 (piecewise constant sigma(t), a(t))."""

    def __init__(self, timeDiscretization, volatility, meanReversion):
        self.timeDiscretization, self.volatility, self.meanReversion = timeDiscretization, list(volatility), list(meanReversion)

    def getTimeDiscretization(self): return self.timeDiscretization
    def getVolatility(self, timeIndex): return Scalar(self.volatility[timeIndex])
    def getMeanReversion(self, timeIndex): return Scalar(self.meanReversion[timeIndex])


class HullWhiteModel:
    """J/montecarlo/interestrate/models/HullWhiteModel.java — the process part (:277-424) with the closed forms for piecewise
    constant coefficients (:584-795).  All coefficients are deterministic Scalars evaluated on the host in the reference's
    operation order; the fused kernel receives them as per-step tables (drift multipliers of x0, four factor loadings)."""

    def __init__(self, randomVariableFactory, liborPeriodDiscretization, volatilityModel, properties=None, discountFactors=None,
                 discountFactorsFromForwardCurve=None):
        """discountFactors[i] = discountCurve.getDiscountFactor(T_i) (or None: no discount curve), discountFactorsFromForwardCurve[i] =
        DiscountCurveFromForwardCurve(forwardRateCurve).getDiscountFactor(T_i), both on the tenor grid (curves are host-side inputs)."""
        self.randomVariableFactory = randomVariableFactory if randomVariableFactory is not None else RandomVariableCudaFactory()
        self.liborPeriodDiscretization = liborPeriodDiscretization
        self.volatilityModel = volatilityModel
        self._properties = properties
        self.dfDiscount = None if discountFactors is None else np.asarray(discountFactors, dtype=np.float64)
        self.dfForward = None if discountFactorsFromForwardCurve is None else np.asarray(discountFactorsFromForwardCurve, dtype=np.float64)
        self._numeraireDiscountFactors, self._dfFromForwardCache, self._forwardRateCache = [], [], []
        self._mrTimeCache = {}
        self._fusedSpecCache = {}                            # time discretization -> per-step coefficient tables

    def getReferenceDate(self): return None

    def getCloneWithModifiedData(self, dataModified):        # :478-487: randomVariableFactory, volatilityModel
        d = dict(dataModified or {})
        return HullWhiteModel(d.get("randomVariableFactory", self.randomVariableFactory), self.liborPeriodDiscretization,
                              d.get("volatilityModel", self.volatilityModel), self._properties, self.dfDiscount, self.dfForward)

    def getVolatilityModel(self): return self.volatilityModel                                      # :805-812

    def getCloneWithModifiedVolatilityModel(self, volatilityModel):                                # :801-804
        return self.getCloneWithModifiedData({"volatilityModel": volatilityModel})

    def getForwardDiscountBond(self, process, time, maturity):                                     # :360-365
        if self.dfDiscount is None:
            raise ValueError("getForwardDiscountBond needs a discount curve (the reference dereferences it)")
        inverseForwardBondAsOfTime = self.getForwardRate(process, time, time, maturity).mult(maturity - time).add(1.0)
        inverseForwardBondAsOfZero = self.getForwardRate(process, 0.0, time, maturity).mult(maturity - time).add(1.0)
        forwardDiscountBondAsOfZero = self._curve_interpolated(maturity, self._discount_factor_at).div(self._curve_interpolated(time, self._discount_factor_at))
        return forwardDiscountBondAsOfZero.mult(inverseForwardBondAsOfZero).div(inverseForwardBondAsOfTime)

    def getIntegratedBondSquaredVolatility(self, time, maturity):                                  # :797-799
        return self.getShortRateConditionalVariance(0, time).mult(self.getB(time, maturity).squared())

    def getNumberOfComponents(self): return 2
    def getNumberOfFactors(self): return 1                                                        # :287-290 (sic; the driver's count is used)
    def getLiborPeriodDiscretization(self): return self.liborPeriodDiscretization
    def getRandomVariableForConstant(self, v): return self.randomVariableFactory.createRandomVariable(v)
    def applyStateSpaceTransform(self, process, timeIndex, componentIndex, rv): return rv
    def applyStateSpaceTransformInverse(self, process, timeIndex, componentIndex, rv): return rv

    def getInitialState(self, process):
        zero = self.getRandomVariableForConstant(0.0)
        return [zero, zero]

    def _vol_index(self, t):
        i = self.volatilityModel.getTimeDiscretization().getTimeIndex(t)
        return i if i >= 0 else -i - 2

    def getMRTime(self, time, maturity):                                                          # :584-607
        # (memoised: the integrals below ask for the same (time, maturity) pairs again and again; the values are immutable Scalars)
        key = (time, maturity)
        hit = self._mrTimeCache.get(key)
        if hit is None:
            hit = self._mrTimeCache[key] = self._getMRTime(time, maturity)
        return hit

    def _getMRTime(self, time, maturity):
        vm, td = self.volatilityModel, self.volatilityModel.getTimeDiscretization()
        i0, i1 = self._vol_index(time), self._vol_index(maturity)
        integral, timePrev = Scalar(0.0), time
        for ti in range(i0 + 1, i1 + 1):
            timeNext = td.getTime(ti)
            integral = integral.add(vm.getMeanReversion(ti - 1).mult(timeNext - timePrev))
            timePrev = timeNext
        return integral.add(vm.getMeanReversion(i1).mult(maturity - timePrev))

    def getB(self, time, maturity):                                                               # :609-640
        vm, td = self.volatilityModel, self.volatilityModel.getTimeDiscretization()
        i0, i1 = self._vol_index(time), self._vol_index(maturity)
        integral, timePrev = Scalar(0.0), time
        for ti in range(i0 + 1, i1 + 1):
            timeNext = td.getTime(ti)
            integral = integral.add(self.getMRTime(timeNext, maturity).mult(-1.0).exp().sub(
                self.getMRTime(timePrev, maturity).mult(-1.0).exp()).div(vm.getMeanReversion(ti - 1)))
            timePrev = timeNext
        return integral.add(self.getMRTime(maturity, maturity).mult(-1.0).exp().sub(
            self.getMRTime(timePrev, maturity).mult(-1.0).exp()).div(vm.getMeanReversion(i1)))

    def _segments(self, time, maturity):
        vm, td = self.volatilityModel, self.volatilityModel.getTimeDiscretization()
        i0, i1 = self._vol_index(time), self._vol_index(maturity)
        timePrev = time
        for ti in range(i0 + 1, i1 + 1):
            timeNext = td.getTime(ti)
            yield timePrev, timeNext, vm.getMeanReversion(ti - 1), vm.getVolatility(ti - 1)
            timePrev = timeNext
        yield timePrev, maturity, vm.getMeanReversion(i1), vm.getVolatility(i1)

    def getV(self, time, maturity):                                                               # :642-690
        if time == maturity:
            return Scalar(0.0)
        integral = Scalar(0.0)
        ePrev = self.getMRTime(time, maturity).mult(-1).exp()
        for timePrev, timeNext, m, v in self._segments(time, maturity):
            v2 = v.squared().div(m.squared())
            eNext = self.getMRTime(timeNext, maturity).mult(-1).exp()
            integral = integral.add(v2.mult(eNext.sub(ePrev).mult(-2).div(m).add(eNext.squared().sub(ePrev.squared()).div(m).div(2.0)).add(timeNext - timePrev)))
            ePrev = eNext
        return integral

    def getDV(self, time, maturity):                                                              # :692-738
        if time == maturity:
            return Scalar(0.0)
        integral = Scalar(0.0)
        ePrev = self.getMRTime(time, maturity).mult(-1).exp()
        for timePrev, timeNext, m, v in self._segments(time, maturity):
            v2 = v.squared().div(m.squared())
            eNext = self.getMRTime(timeNext, maturity).mult(-1).exp()
            integral = integral.add(v2.mult(eNext.sub(ePrev).add(eNext.squared().sub(ePrev.squared()).div(-2.0))))
            ePrev = eNext
        return integral

    def _drift_coefficients(self, process, timeIndex):                                            # :367-387
        time, timeNext = process.getTime(timeIndex), process.getTime(timeIndex + 1)
        m = self.volatilityModel.getMeanReversion(self._vol_index(time))
        B = self.getB(time, timeNext)
        return m.mult(B.div(-1 * (timeNext - time))), B.div(timeNext - time)

    def getDrift(self, process, timeIndex, x, predictor):
        if process.getTime(timeIndex + 1) == process.getTime(timeIndex):
            return [None, None]
        c0, c1 = self._drift_coefficients(process, timeIndex)
        return [x[0].mult(c0), x[0].mult(c1)]

    def getFactorLoading(self, process, timeIndex, componentIndex, x):                            # :389-424
        time, timeNext = process.getTime(timeIndex), process.getTime(timeIndex + 1)
        vi = self._vol_index(time)
        m = self.volatilityModel.getMeanReversion(vi)
        mrt = m.mult(-2.0 * (timeNext - time))
        scaling = mrt.exp().sub(1.0).div(mrt).sqrt()
        volEff = scaling.mult(self.volatilityModel.getVolatility(vi))
        if componentIndex == 0:
            return [volEff, Scalar(0.0)]
        volLogNum = self.getV(time, timeNext).div(timeNext - time).sqrt()
        rho = self.getDV(time, timeNext).div(timeNext - time).div(volEff.mult(volLogNum))
        return [volLogNum.mult(rho), volLogNum.mult(rho.squared().sub(1).mult(-1).sqrt())]

    def getFusedSpecification(self, process):
        if process.getScheme() in (Scheme.PREDICTOR_CORRECTOR, Scheme.PREDICTOR_CORRECTOR_FUNCTIONAL):
            return None                                       # the corrector is not fused for this model: generic device loop
        td = process.getTimeDiscretization()
        T = td.getNumberOfTimeSteps()
        # the per-step coefficients depend on the (immutable) model and the time grid only: a model that is simulated again - another
        # seed, another path count - does not evaluate its closed forms again (they are ~100 Scalar operations per step on the host)
        spec = self._fusedSpecCache.get(td)
        if spec is None:
            d0, d1, fl = [], [], []
            for t in range(T):
                c0, c1 = self._drift_coefficients(process, t)
                f0, f1 = self.getFactorLoading(process, t, 0, None), self.getFactorLoading(process, t, 1, None)
                d0.append(c0.doubleValue())
                d1.append(c1.doubleValue())
                fl.append([f0[0].doubleValue(), f0[1].doubleValue(), f1[0].doubleValue(), f1[1].doubleValue()])
            spec = self._fusedSpecCache[td] = dict(kernel="hull_white", drift0=d0, drift1=d1, factorLoadings=np.array(fl), initialValues=[0.0, 0.0])
        return spec

    # ---- term structure functions of the Hull-White model (:305-357, :431-582, :797-950); deterministic parts are Scalars -------------
    def getLiborPeriod(self, i): return self.liborPeriodDiscretization.getTime(i)
    def getLiborPeriodIndex(self, t): return self.liborPeriodDiscretization.getTimeIndex(t)
    def getNumberOfLibors(self): return self.liborPeriodDiscretization.getNumberOfTimeSteps()
    def getModel(self): return self

    def getShortRateConditionalVariance(self, time, maturity):                                     # :740-775
        integral = Scalar(0.0)
        ePrev = self.getMRTime(time, maturity).mult(-2).exp()
        for timePrev, timeNext, m, v in self._segments(time, maturity):
            eNext = self.getMRTime(timeNext, maturity).mult(-2).exp()
            integral = integral.add(v.squared().div(m).mult(eNext.sub(ePrev).div(2)))
            ePrev = eNext
        return integral

    def _forward_rate_initial_value(self, i):                                                      # :935-950
        td = self.liborPeriodDiscretization
        while len(self._forwardRateCache) <= i:
            k = len(self._forwardRateCache)
            self._forwardRateCache.append(self.getRandomVariableForConstant((self.dfForward[k] / self.dfForward[k + 1] - 1.0) / td.getTimeStep(k)))
        return self._forwardRateCache[i]

    def _df_from_forward_curve_at(self, timeIndex):                                                # :912-933
        td = self.liborPeriodDiscretization
        while len(self._dfFromForwardCache) <= timeIndex:
            i = len(self._dfFromForwardCache)
            if i == 0:
                self._dfFromForwardCache.append(self.getRandomVariableForConstant(float(self.dfForward[0])))
            else:
                self._dfFromForwardCache.append(self._dfFromForwardCache[i - 1].div(self._forward_rate_initial_value(i - 1).mult(td.getTimeStep(i - 1)).add(1.0)))
        return self._dfFromForwardCache[timeIndex]

    def _curve_interpolated(self, time, at):                                                       # :845-860, :896-910
        td = self.liborPeriodDiscretization
        ti = td.getTimeIndex(time)
        if ti >= 0:
            return at(ti)
        prev = min(-ti - 2, td.getNumberOfTimes() - 2)
        tp, tn = td.getTime(prev), td.getTime(prev + 1)
        a, b = at(prev), at(prev + 1)
        return a.mult(b.div(a).pow((time - tp) / (tn - tp)))

    def _df_from_forward_curve(self, time): return self._curve_interpolated(time, self._df_from_forward_curve_at)

    def _discount_factor_at(self, timeIndex):                                                      # :862-889
        td = self.liborPeriodDiscretization
        if not self._numeraireDiscountFactors:
            adj = self.getRandomVariableForConstant(float(self.dfDiscount[0]))
            self._numeraireDiscountFactors.append(adj)
            for i in range(td.getNumberOfTimeSteps()):
                ts = td.getTimeStep(i)
                adj = adj.discount(self.getRandomVariableForConstant((self.dfDiscount[i] / self.dfDiscount[i + 1] - 1.0) / ts), ts)
                self._numeraireDiscountFactors.append(adj)
        return self._numeraireDiscountFactors[timeIndex]

    def _zero_rate_from_forward_curve(self, time):                                                 # :891-902 (same index twice: sic)
        td = self.liborPeriodDiscretization
        ti = td.getTimeIndex(time)
        if ti < 0:
            ti = min(-ti - 2, td.getNumberOfTimes() - 2)
        d = self._df_from_forward_curve_at(ti)
        return d.div(d).log().div(td.getTimeStep(ti))

    def _short_rate(self, process, timeIndex):                                                     # :493-510
        time = process.getTime(timeIndex)
        value = process.getProcessValue(timeIndex, 0).add(self.getDV(0, time))
        return value.add(self._zero_rate_from_forward_curve(time))

    def _A(self, process, time, maturity):                                                         # :543-557
        zeroRate = self._zero_rate_from_forward_curve(time)
        forwardBond = self._df_from_forward_curve(maturity).div(self._df_from_forward_curve(time)).log()
        B = self.getB(time, maturity)
        return B.mult(zeroRate).sub(B.squared().mult(self.getShortRateConditionalVariance(0, time).div(2))).add(forwardBond).exp()

    def getZeroCouponBond(self, process, time, maturity):                                          # :512-523
        ti = process.getTimeIndex(time)
        if ti < 0:
            timeLo = process.getTime(-ti - 1 - 1)
            return self.getZeroCouponBond(process, timeLo, maturity).div(self.getZeroCouponBond(process, timeLo, time))
        return self._short_rate(process, ti).mult(self.getB(time, maturity).mult(-1)).exp().mult(self._A(process, time, maturity))

    def getForwardRate(self, process, time, periodStart, periodEnd):                               # :431-435
        return self.getZeroCouponBond(process, time, periodStart).div(self.getZeroCouponBond(process, time, periodEnd)).sub(1.0).div(periodEnd - periodStart)

    def getLIBOR(self, process, timeIndex, liborIndex):                                            # :437-440
        t = process.getTime(timeIndex)
        return self.getZeroCouponBond(process, t, self.getLiborPeriod(liborIndex)).div(self.getZeroCouponBond(process, t, self.getLiborPeriod(liborIndex + 1))) \
            .sub(1.0).div(self.liborPeriodDiscretization.getTimeStep(liborIndex))

    def getNumeraire(self, process, time):                                                         # :305-357
        if time == process.getTime(0):
            return self.getRandomVariableForConstant(1.0)
        ti = process.getTimeIndex(time)
        if ti < 0:                                           # between simulation times: log-linear interpolation (:317-333)
            previousTimeIndex = (-ti - 1) - 1
            previousTime, nextTime = process.getTime(previousTimeIndex), process.getTime(previousTimeIndex + 1)
            return self.getNumeraire(process, previousTime).log().mult(nextTime - time) \
                .add(self.getNumeraire(process, nextTime).log().mult(time - previousTime)).div(nextTime - previousTime).exp()
        numeraireNormalized = process.getProcessValue(ti, 1).add(self.getV(0, time).mult(0.5)).exp()
        numeraireNormalized = numeraireNormalized.mult(numeraireNormalized.invert().getAverage())   # control variate on the zero bond
        fromForward = self._df_from_forward_curve(time)
        if self.dfDiscount is not None:
            discountFactor = self._curve_interpolated(time, self._discount_factor_at).div(fromForward.getAverage()).mult(fromForward)
        else:
            discountFactor = fromForward
        return numeraireNormalized.div(discountFactor)
