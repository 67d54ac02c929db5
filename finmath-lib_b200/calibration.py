"""Calibration loops over the device simulation — SURVEY.md §8f rank 4 (second half): the optimiser and the generic calibration of a
parametric LIBOR covariance model.  Host-side control logic around the hot path: every objective evaluation re-simulates the model with
the FUSED Euler kernel on Brownian increments that stay resident in HBM for the whole calibration (generated once), and values the
calibration products with device RandomVariable operations; only the K averaged residuals come back per evaluation.

* ``LevenbergMarquardt``  — J/optimizer/LevenbergMarquardt.java (defaults :146-169, finite-difference derivatives :553-616, termination
  :621-638, main loop :640-726, mean squared error :728-735, trial step :737-812, clone with new targets :826-850)
* ``CalibrationProduct``  — J/montecarlo/interestrate/CalibrationProduct.java:18-100
* ``getCloneCalibrated``  — J/montecarlo/interestrate/models/covariance/AbstractLIBORCovarianceModelParametric.java:333-470 (the
  double-valued optimiser path: parameter names, defaults and the objective  (value_i - target_i) · weight_i  against zero)

The reference evaluates the finite-difference columns and the products on thread pools; here the device is the parallel resource, so
evaluations are issued one after the other on the library's stream (the Python host code between launches is the serial part).
"""
import copy
import math

import numpy as np

from . import native as nv
from .stochastic import Scalar


class SolverException(Exception):
    pass


class RegularizationMethod:
    LEVENBERG = "LEVENBERG"                                  # H + lambda I
    LEVENBERG_MARQUARDT = "LEVENBERG_MARQUARDT"              # H + lambda diag(H)


def solveLinearEquationSVD(A, b):
    """LinearAlgebra.solveLinearEquationSVD (J/functions/LinearAlgebra.java:256-265): commons-math SingularValueDecomposition solver =
    pseudo-inverse with singular values below max(rows, cols) · ulp(s_max) dropped.  A handful of unknowns: host arithmetic, as in the
    reference."""
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if not (np.all(np.isfinite(A)) and np.all(np.isfinite(b))):
        raise SolverException("matrix contains non-finite entries")
    U, s, Vt = np.linalg.svd(A)
    tol = max(max(A.shape) * np.spacing(s[0]), math.sqrt(np.finfo(np.float64).tiny)) if s.size else 0.0
    inv = np.array([1.0 / x if x > tol else 0.0 for x in s])
    return Vt.T @ (inv * (U.T @ b))


class LevenbergMarquardt:
    """Minimises  sum_k w_k (f_k(x) - y_k)^2 / n.  Either subclass and override ``setValues(parameters, values)`` (and optionally
    ``setDerivatives(parameters, derivatives)``, derivatives[parameter][value]) like the reference's anonymous classes, or pass
    ``objectiveFunction=callable(parameters, values)``."""

    def __init__(self, initialParameters=None, targetValues=None, maxIteration=100, numberOfThreads=1, regularizationMethod=None,
                 objectiveFunction=None):
        self.regularizationMethod = regularizationMethod if regularizationMethod is not None else RegularizationMethod.LEVENBERG_MARQUARDT
        self.initialParameters = None if initialParameters is None else np.array(initialParameters, dtype=np.float64)
        self.targetValues = None if targetValues is None else np.array(targetValues, dtype=np.float64)
        self.weights = None if targetValues is None else np.ones(len(self.targetValues))
        self.parameterSteps = None
        self.maxIteration = int(maxIteration)
        self.numberOfThreads = numberOfThreads               # accepted for signature parity; evaluations are stream-ordered on the device
        self.lambda_ = 0.001
        self.lambdaDivisor = 3.0
        self.lambdaMultiplicator = 2.0
        self.errorRootMeanSquaredTolerance = 0.0
        self._objective = objectiveFunction
        self._reset()

    def _reset(self):
        self.iteration = 0
        self.parameterTest = self.parameterCurrent = self.valueTest = self.valueCurrent = self.derivativeCurrent = None
        self.parameterIncrement = None
        self.errorMeanSquaredCurrent = float("inf")
        self.errorRootMeanSquaredChange = float("inf")
        self.isParameterCurrentDerivativeValid = False
        self.numberOfEvaluations = 0

    # ---- configuration (:385-472): not after the solver has run ---------------------------------------------------------------------
    def _modifiable(self):
        if self.done():
            raise NotImplementedError("Solver cannot be modified after it has run.")          # UnsupportedOperationException

    def setInitialParameters(self, v):
        self._modifiable()
        self.initialParameters = np.array(v, dtype=np.float64)
        return self

    def setParameterSteps(self, v):
        self._modifiable()
        self.parameterSteps = None if v is None else np.array(v, dtype=np.float64)
        return self

    def setTargetValues(self, v):
        self._modifiable()
        self.targetValues = np.array(v, dtype=np.float64)
        if self.weights is None or len(self.weights) != len(self.targetValues):
            self.weights = np.ones(len(self.targetValues))
        return self

    def setMaxIteration(self, n):
        self._modifiable()
        self.maxIteration = int(n)
        return self

    def setWeights(self, w):
        self._modifiable()
        self.weights = np.array(w, dtype=np.float64)
        return self

    def setErrorTolerance(self, tol):
        self._modifiable()
        self.errorRootMeanSquaredTolerance = float(tol)
        return self

    def getLambda(self): return self.lambda_

    def setLambda(self, v):
        self.lambda_ = float(v)
        return self

    def getLambdaMultiplicator(self): return self.lambdaMultiplicator

    def setLambdaMultiplicator(self, v):
        if v <= 1.0:
            raise ValueError("Parameter lambdaMultiplicator is required to be > 1.")
        self.lambdaMultiplicator = float(v)

    def getLambdaDivisor(self): return self.lambdaDivisor

    def setLambdaDivisor(self, v):
        if v <= 1.0:
            raise ValueError("Parameter lambdaDivisor is required to be > 1.")
        self.lambdaDivisor = float(v)

    # ---- results ---------------------------------------------------------------------------------------------------------------------
    def getBestFitParameters(self): return self.parameterCurrent
    def getRootMeanSquaredError(self): return math.sqrt(self.errorMeanSquaredCurrent)
    def getIterations(self): return self.iteration

    # ---- the objective ---------------------------------------------------------------------------------------------------------------
    def setValues(self, parameters, values):
        if self._objective is None:
            raise NotImplementedError("override setValues or pass objectiveFunction")
        self._objective(parameters, values)

    def setDerivatives(self, parameters, derivatives):
        """One-sided finite differences against valueCurrent (called with parameters == parameterCurrent only) :553-616."""
        for p in range(len(self.parameterCurrent)):
            shifted = np.array(parameters, dtype=np.float64)
            step = self.parameterSteps[p] if self.parameterSteps is not None else (abs(shifted[p]) + 1) * 1e-8
            shifted[p] += step
            column = derivatives[p]
            try:
                self._evaluate(shifted, column)
            except Exception:
                column[:] = float("nan")                     # "we signal an exception to calculate the derivative as NaN" -> 0 below
            column -= self.valueCurrent
            column /= step
            column[np.isnan(column)] = 0.0

    def _evaluate(self, parameters, values):
        self.numberOfEvaluations += 1
        self.setValues(parameters, values)

    def done(self):
        return (self.iteration > self.maxIteration or self.errorRootMeanSquaredChange <= self.errorRootMeanSquaredTolerance
                or math.isinf(self.lambda_))

    def getMeanSquaredError(self, value):
        error = 0.0
        for k in range(len(value)):
            deviation = value[k] - self.targetValues[k]
            error += self.weights[k] * deviation * deviation
        return error / len(value)

    def run(self):
        P, V = len(self.initialParameters), len(self.targetValues)
        self.parameterTest = np.array(self.initialParameters, dtype=np.float64)
        self.parameterIncrement = np.zeros(P)
        self.parameterCurrent = np.zeros(P)
        self.valueTest, self.valueCurrent = np.zeros(V), np.zeros(V)
        self.derivativeCurrent = np.zeros((P, V))
        self.iteration = 0
        while True:
            self.iteration += 1
            self._evaluate(self.parameterTest, self.valueTest)
            errorMeanSquaredTest = self.getMeanSquaredError(self.valueTest)
            # accept the point if it is better than the current one (a NaN error is never better)
            if errorMeanSquaredTest < self.errorMeanSquaredCurrent:
                self.errorRootMeanSquaredChange = math.sqrt(self.errorMeanSquaredCurrent) - math.sqrt(errorMeanSquaredTest)
                self.parameterCurrent[:] = self.parameterTest
                self.valueCurrent[:] = self.valueTest
                self.errorMeanSquaredCurrent = errorMeanSquaredTest
                self.isParameterCurrentDerivativeValid = False
                self.lambda_ /= self.lambdaDivisor              # move faster
            else:
                self.errorRootMeanSquaredChange = math.sqrt(errorMeanSquaredTest) - math.sqrt(self.errorMeanSquaredCurrent)
                self.lambda_ *= self.lambdaMultiplicator         # reject: move slower
            if self.done():
                break
            self._updateParameterTest()

    def _updateParameterTest(self):
        if not self.isParameterCurrentDerivativeValid:
            self.setDerivatives(self.parameterCurrent, self.derivativeCurrent)
            self.isParameterCurrentDerivativeValid = True
        J, w = self.derivativeCurrent, self.weights
        P = len(self.parameterCurrent)
        hessianInvalid = True
        while hessianInvalid and math.isfinite(self.lambda_):
            hessianInvalid = False
            H = np.zeros((P, P))
            for i in range(P):
                for j in range(i, P):
                    h = float(np.sum(w * J[i] * J[j]))
                    if i == j:
                        if self.regularizationMethod == RegularizationMethod.LEVENBERG:
                            h += self.lambda_
                        elif h == 0.0:
                            h = self.lambda_
                        else:
                            h *= 1 + self.lambda_
                    H[i, j] = H[j, i] = h
            beta = np.array([float(np.sum(w * (self.targetValues - self.valueCurrent) * J[i])) for i in range(P)])
            try:
                self.parameterIncrement = solveLinearEquationSVD(H, beta)
            except Exception:
                hessianInvalid = True                         # not invertible: increase lambda
                self.lambda_ *= 16
        self.parameterTest = self.parameterCurrent + self.parameterIncrement

    def clone(self):
        c = copy.copy(self)
        c._reset()
        return c

    def getCloneWithModifiedTargetValues(self, newTargetValues, newWeights, isUseBestParametersAsInitialParameters):
        c = self.clone()
        c.targetValues = np.array(newTargetValues, dtype=np.float64)
        c.weights = np.array(newWeights, dtype=np.float64)
        if isUseBestParametersAsInitialParameters and self.done():
            c.initialParameters = np.array(self.getBestFitParameters(), dtype=np.float64)
        return c


class OptimizerFactoryLevenbergMarquardt:
    """J/optimizer/OptimizerFactoryLevenbergMarquardt.java: (maxIterations, errorTolerance, maxThreads) -> optimiser with weights 1."""

    def __init__(self, maxIterations=100, errorTolerance=0.0, maxThreads=1, regularizationMethod=None):
        self.maxIterations, self.errorTolerance, self.maxThreads, self.regularizationMethod = maxIterations, errorTolerance, maxThreads, regularizationMethod

    def getOptimizer(self, objectiveFunction, initialParameters, lowerBound=None, upperBound=None, parameterSteps=None, targetValues=None):
        o = LevenbergMarquardt(initialParameters, targetValues, self.maxIterations, self.maxThreads, self.regularizationMethod, objectiveFunction)
        o.setErrorTolerance(self.errorTolerance)
        if parameterSteps is not None:
            o.setParameterSteps(parameterSteps)
        return o


class CalibrationProduct:
    def __init__(self, product, targetValue, weight, name=None, priority=0):
        self.name, self.product, self.weight, self.priority = name, product, float(weight), priority
        self.targetValue = targetValue if hasattr(targetValue, "getTypePriority") else Scalar(targetValue)

    def getName(self): return self.name if self.name is not None else repr(self.product)
    def getProduct(self): return self.product
    def getTargetValue(self): return self.targetValue
    def getWeight(self): return self.weight
    def getPriority(self): return self.priority

    def __repr__(self):
        return "CalibrationProduct [product=%r, targetValue=%r, weight=%r]" % (self.product, self.targetValue, self.weight)


def getCloneCalibrated(covarianceModel, calibrationModel, calibrationProducts, calibrationParameters=None):
    """AbstractLIBORCovarianceModelParametric.getCloneCalibratedLegazy :333-470.

    calibrationParameters: numberOfPaths (2000), seed (31415), maxIterations (400), parameterStep (1e-4), accuracy (1e-7),
    brownianMotion, optimizerFactory — the reference's keys and defaults; plus ``shard`` (paths of the calibration simulation sharded
    over ranks: every rank sees the same averaged residuals, so all ranks take identical steps).
    After the call ``covarianceModel.lastCalibration`` holds {iterations, evaluations, rootMeanSquaredError, bestParameters}."""
    from .models import LIBORMonteCarloSimulationFromLIBORModel
    from .montecarlo import BrownianMotionCuda, EulerSchemeFromProcessModel
    p = dict(calibrationParameters or {})
    initialParameters = np.array(covarianceModel.getParameterAsDouble(), dtype=np.float64)
    K = len(calibrationProducts)
    parameterStep = np.full(len(initialParameters), float(p.get("parameterStep", 1e-4)))
    numberOfPaths, seed = int(p.get("numberOfPaths", 2000)), int(p.get("seed", 31415))
    maxIterations, accuracy = int(p.get("maxIterations", 400)), float(p.get("accuracy", 1e-7))
    brownianMotion = p.get("brownianMotion")
    if brownianMotion is None:
        brownianMotion = BrownianMotionCuda(covarianceModel.getTimeDiscretization(), covarianceModel.getNumberOfFactors(), numberOfPaths, seed,
                                            shard=p.get("shard"))
    optimizerFactory = p.get("optimizerFactory") or OptimizerFactoryLevenbergMarquardt(maxIterations, accuracy, 2)

    def calibrationError(parameters, values):
        candidate = covarianceModel.getCloneWithModifiedParameters(parameters)
        model = calibrationModel.getCloneWithModifiedCovarianceModel(candidate)
        simulation = LIBORMonteCarloSimulationFromLIBORModel(EulerSchemeFromProcessModel(model, brownianMotion))
        residuals = []
        for item in calibrationProducts:
            try:
                residuals.append(item.getProduct().getValueRV(0.0, simulation).sub(item.getTargetValue()).mult(item.getWeight()))
            except (nv.NoDeviceError, nv.FmbError, MemoryError):
                raise                                        # a failing device is not a "non-working product"
            except Exception:
                residuals.append(None)                       # "automatically exclude non-working calibration products" (:395-398)
        for k, r in enumerate(residuals):                    # the averages are read after all products are queued
            values[k] = r.getAverage() if r is not None else 0.0

    optimizer = optimizerFactory.getOptimizer(calibrationError, initialParameters, None, None, parameterStep, np.zeros(K))
    optimizer.run()
    best = optimizer.getBestFitParameters()
    calibrated = covarianceModel.getCloneWithModifiedParameters(best)
    calibrated.lastCalibration = dict(iterations=optimizer.getIterations(), evaluations=getattr(optimizer, "numberOfEvaluations", None),
                                      rootMeanSquaredError=optimizer.getRootMeanSquaredError(), bestParameters=np.array(best))
    return calibrated
