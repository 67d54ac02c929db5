"""Product consumers — pure RandomVariable algebra, exactly the reference's call sequence, so that they exercise the device
type the way the unchanged Java classes would (none of them ever touches realizations on the host).

* EuropeanOption      J/montecarlo/assetderivativevaluation/products/EuropeanOption.java:172-193
* DigitalOption       J/montecarlo/assetderivativevaluation/products/DigitalOption.java:73-93
* Caplet              J/montecarlo/interestrate/products/Caplet.java:114-160
* Swaption            J/montecarlo/interestrate/products/Swaption.java:137-200
* BermudanSwaption    J/montecarlo/interestrate/products/BermudanSwaption.java:90-252
"""
import bisect

from .montecarlo import MonteCarloConditionalExpectationRegression
from .stochastic import RandomVariableFromDoubleArray, Scalar


class AbstractMonteCarloProduct:
    def getValue(self, *args):                               # AbstractMonteCarloProduct.java:81-84: getValue(model) = getValue(0.0, model).getAverage()
        if len(args) == 1:
            return self.getValueRV(0.0, args[0]).getAverage()
        return self.getValueRV(*args)

    def getValues(self, *args):                              # :87-104, :132-135: {"value": average, "error": standard error}
        evaluationTime, model = (0.0, args[0]) if len(args) == 1 else args
        values = self.getValueRV(evaluationTime, model)
        if values is None:
            return None
        return {"value": values.getAverage(), "error": values.getStandardError()}

    def getValuesForModifiedData(self, *args):
        """:110-156: (evaluationTime, model, map) | (evaluationTime, model, key, value) | (model, map) | (model, key, value): the values on
        model.getCloneWithModifiedData(...) — a bump-and-revalue on the same random numbers."""
        if not isinstance(args[0], (int, float)):
            args = (0.0,) + args
        evaluationTime, model = args[0], args[1]
        dataModified = args[2] if len(args) == 3 else {args[2]: args[3]}
        return self.getValues(evaluationTime, model.getCloneWithModifiedData(dataModified))

    def getCurrency(self):                                   # :159-161 (no currency given)
        return None


class EuropeanOption(AbstractMonteCarloProduct):
    def __init__(self, maturity, strike, callOrPutSign=1.0, underlyingIndex=0):
        self.maturity, self.strike, self.sign, self.underlyingIndex = maturity, strike, float(callOrPutSign), underlyingIndex

    def getMaturity(self): return self.maturity              # EuropeanOption.java:195-240
    def getStrike(self): return self.strike
    def getCallOrPut(self): return self.sign
    def getUnderlyingIndex(self): return self.underlyingIndex
    def getNameOfUnderliyng(self): return None               # (sic; assets are addressed by index here)

    def getValueRV(self, evaluationTime, model):
        underlyingAtMaturity = model.getAssetValue(float(self.maturity), self.underlyingIndex)
        values = underlyingAtMaturity.sub(self.strike).mult(self.sign).floor(0.0)
        numeraireAtMaturity = model.getNumeraire(float(self.maturity))
        monteCarloWeights = model.getMonteCarloWeights(float(self.maturity))
        values = values.div(numeraireAtMaturity).mult(monteCarloWeights)
        numeraireAtEvalTime = model.getNumeraire(float(evaluationTime))
        monteCarloWeightsAtEvalTime = model.getMonteCarloWeights(float(evaluationTime))
        return values.mult(numeraireAtEvalTime).div(monteCarloWeightsAtEvalTime)


class DigitalOption(AbstractMonteCarloProduct):
    """J/montecarlo/assetderivativevaluation/products/DigitalOption.java:73-93 — the indicator payoff 1(S(T) - K >= 0) (a choose: the
    operation whose derivative autodiff.py approximates)."""

    def __init__(self, maturity, strike, underlyingIndex=0):
        self.maturity, self.strike, self.underlyingIndex = maturity, strike, underlyingIndex

    def getValueRV(self, evaluationTime, model):
        underlyingAtMaturity = model.getAssetValue(float(self.maturity), self.underlyingIndex)
        values = underlyingAtMaturity.sub(self.strike).choose(Scalar(1.0), Scalar(0.0))
        values = values.div(model.getNumeraire(float(self.maturity))).mult(model.getMonteCarloWeights(float(self.maturity)))
        return values.mult(model.getNumeraire(float(evaluationTime))).div(model.getMonteCarloWeights(float(evaluationTime)))


class Caplet(AbstractMonteCarloProduct):
    def __init__(self, maturity, periodLength, strike, daycountFraction=None, isFloorlet=False):
        self.maturity, self.periodLength, self.strike = maturity, periodLength, strike
        self.daycountFraction = periodLength if daycountFraction is None else daycountFraction
        self.isFloorlet = isFloorlet

    def getValueRV(self, evaluationTime, model):
        paymentDate = self.maturity + self.periodLength
        forwardRate = model.getForwardRate(self.maturity, self.maturity, self.maturity + self.periodLength)
        numeraire = model.getNumeraire(paymentDate)
        monteCarloProbabilities = model.getMonteCarloWeights(paymentDate)
        if not self.isFloorlet:
            values = forwardRate.sub(self.strike).floor(0.0).mult(self.daycountFraction)
        else:
            values = forwardRate.sub(self.strike).cap(0.0).mult(-1.0 * self.daycountFraction)
        values = values.div(numeraire).mult(monteCarloProbabilities)
        return values.mult(model.getNumeraire(float(evaluationTime))).div(model.getMonteCarloWeights(float(evaluationTime)))


class Swaption(AbstractMonteCarloProduct):
    def __init__(self, exerciseDate, fixingDates, paymentDates, swaprates, notional=1.0, periodLengths=None):
        self.exerciseDate, self.fixingDates, self.paymentDates, self.swaprates = exerciseDate, list(fixingDates), list(paymentDates), list(swaprates)
        self.notional, self.periodLengths = notional, periodLengths

    def getExerciseDate(self): return self.exerciseDate      # Swaption.java:241-275
    def getFixingDates(self): return self.fixingDates
    def getPaymentDates(self): return self.paymentDates
    def getPeriodLengths(self): return self.periodLengths
    def getSwaprates(self): return self.swaprates
    def getNotional(self): return self.notional

    def getExerciseIndicator(self, model):                   # :241-243: 1 where the swaption is exercised (value at the exercise date > 0)
        return self.getValueRV(self.exerciseDate, model).mult(-1.0).choose(Scalar(0.0), Scalar(1.0))

    def getValueRV(self, evaluationTime, model):
        value = model.getRandomVariableForConstant(0.0)
        for period in range(len(self.fixingDates) - 1, -1, -1):
            fixingDate, paymentDate, swaprate = self.fixingDates[period], self.paymentDates[period], self.swaprates[period]
            if paymentDate <= evaluationTime:
                break
            periodLength = self.periodLengths[period] if self.periodLengths is not None else paymentDate - fixingDate
            libor = model.getForwardRate(self.exerciseDate, fixingDate, paymentDate)
            payoff = libor.sub(swaprate).mult(periodLength).mult(self.notional)
            discountingDate = max(fixingDate, self.exerciseDate)
            discountingAdjustment = self._discounting_adjustment(model, discountingDate, paymentDate)
            value = value.add(payoff)
            value = value.discount(libor, paymentDate - discountingDate).mult(discountingAdjustment)
        values = value.floor(0.0)
        values = values.div(model.getNumeraire(float(self.exerciseDate))).mult(model.getMonteCarloWeights(float(self.exerciseDate)))
        return values.mult(model.getNumeraire(float(evaluationTime))).div(model.getMonteCarloWeights(float(evaluationTime)))


    @staticmethod
    def _discounting_adjustment(model, discountingDate, paymentDate):
        """:160-171 — forwardBondOnForwardCurve / forwardBondOnDiscountCurve, 1.0 without a discount curve.  The curves are host-side
        inputs given on the tenor grid (model.discountFactors; the forward-implied curve is DiscountCurveFromForwardCurve :130-142)."""
        m = model.getModel() if hasattr(model, "getModel") else None
        discountFactors = getattr(m, "discountFactors", None)
        if m is None or discountFactors is None:
            return 1.0
        i0, i1 = m.getLiborPeriodIndex(discountingDate), m.getLiborPeriodIndex(paymentDate)
        if i0 < 0 or i1 < 0:
            raise NotImplementedError("discounting adjustment for dates off the tenor grid (curve interpolation) is outside the hot path")
        implied = m.getDiscountFactorsFromForwardCurve()
        forwardBondOnForwardCurve = implied[i0] / implied[i1]
        forwardBondOnDiscountCurve = float(discountFactors[i0]) / float(discountFactors[i1])
        return forwardBondOnForwardCurve / forwardBondOnDiscountCurve


class BermudanSwaption(AbstractMonteCarloProduct):
    def __init__(self, isPeriodStartDateExerciseDate, fixingDates, periodLengths, paymentDates, periodNotionals, swaprates, isCallable=True,
                 regressionBasisFunctionsProvider=None):
        self.isExercise, self.fixingDates, self.periodLengths = list(isPeriodStartDateExerciseDate), list(fixingDates), list(periodLengths)
        self.paymentDates, self.periodNotionals, self.swaprates = list(paymentDates), list(periodNotionals), list(swaprates)
        self.isCallable, self.regressionBasisFunctionsProvider = isCallable, regressionBasisFunctionsProvider
        self.lastRegressions = []

    def getFixingDates(self): return self.fixingDates        # BermudanSwaption.java:254-311
    def getPeriodLengths(self): return self.periodLengths
    def getPaymentDates(self): return self.paymentDates
    def getPeriodNotionals(self): return self.periodNotionals
    def getSwapRates(self): return self.swaprates
    def getIsCallable(self): return self.isCallable
    def getFinalMaturity(self): return self.paymentDates[-1]
    def getExerciseTimes(self): return [t for t, e in zip(self.fixingDates, self.isExercise) if e]

    def getValues(self, evaluationTime, model):
        self.lastRegressions = []
        values = model.getRandomVariableForConstant(0.0)
        valuesUnderlying = model.getRandomVariableForConstant(0.0)
        exerciseTime = model.getRandomVariableForConstant(float("inf"))
        for period in range(len(self.fixingDates) - 1, -1, -1):
            fixingDate = self.fixingDates[period]
            exerciseDate = fixingDate
            periodLength, paymentDate = self.periodLengths[period], self.paymentDates[period]
            notional, swaprate = self.periodNotionals[period], self.swaprates[period]
            libor = model.getForwardRate(fixingDate, fixingDate, fixingDate + periodLength)
            payoff = libor.sub(swaprate).mult(periodLength).mult(notional)
            numeraire = model.getNumeraire(float(paymentDate))
            monteCarloProbabilities = model.getMonteCarloWeights(float(paymentDate))
            payoff = payoff.div(numeraire).mult(monteCarloProbabilities)
            if self.isCallable:
                valuesUnderlying = valuesUnderlying.add(payoff)
            else:
                values = values.add(payoff)
            if self.isExercise[period]:
                triggerValuesDiscounted = values.sub(valuesUnderlying)
                estimator = self.getConditionalExpectationEstimator(fixingDate, model)
                triggerValues = triggerValuesDiscounted.getConditionalExpectation(estimator)
                self.lastRegressions.append(estimator)
                values = triggerValues.choose(values, valuesUnderlying)
                exerciseTime = triggerValues.choose(exerciseTime, Scalar(exerciseDate))
        numeraireAtZero = model.getNumeraire(float(evaluationTime))
        monteCarloProbabilitiesAtZero = model.getMonteCarloWeights(float(evaluationTime))
        values = values.mult(numeraireAtZero).div(monteCarloProbabilitiesAtZero)
        return {"value": values, "error": values.getStandardError(), "exerciseTime": exerciseTime}

    def getValueRV(self, evaluationTime, model):
        return self.getValues(evaluationTime, model)["value"]

    def getConditionalExpectationEstimator(self, fixingDate, model):          # :182-187
        basis = (self.regressionBasisFunctionsProvider.getBasisFunctions(fixingDate, model) if self.regressionBasisFunctionsProvider is not None
                 else self.getBasisFunctions(fixingDate, model))
        return MonteCarloConditionalExpectationRegression(basis)

    def getBasisFunctions(self, fixingDate, model):                           # :215-252
        basisFunctions = [RandomVariableFromDoubleArray(1.0)]                  # :220 — the CPU type, on purpose
        i = bisect.bisect_left(self.fixingDates, fixingDate)
        fixingDateIndex = i if (i < len(self.fixingDates) and self.fixingDates[i] == fixingDate) else -(i + 1)
        if fixingDateIndex < 0:
            fixingDateIndex = -fixingDateIndex                                 # :224-226 (sic)
        if fixingDateIndex >= len(self.fixingDates):
            fixingDateIndex = len(self.fixingDates) - 1
        rateShort = model.getForwardRate(fixingDate, fixingDate, self.paymentDates[fixingDateIndex])
        discountShort = rateShort.mult(self.paymentDates[fixingDateIndex] - fixingDate).add(1.0).invert()
        basisFunctions.append(discountShort)
        basisFunctions.append(discountShort.pow(2.0))
        rateLong = model.getForwardRate(fixingDate, self.fixingDates[fixingDateIndex], self.paymentDates[-1])
        discountLong = rateLong.mult(self.paymentDates[-1] - self.fixingDates[fixingDateIndex]).add(1.0).invert()
        basisFunctions.append(discountLong)
        basisFunctions.append(discountLong.pow(2.0))
        numeraire = model.getNumeraire(float(fixingDate)).invert()
        basisFunctions.append(numeraire)
        return basisFunctions


class BermudanOption(AbstractMonteCarloProduct):
    """Asset Bermudan option, J/montecarlo/assetderivativevaluation/products/BermudanOption.java:150-330, exercise method
    ESTIMATE_COND_EXPECTATION (lower bound).  Like the reference it pulls the underlying to the host for its basis functions
    (:298, :314 `new RandomVariableFromDoubleArray(0.0, underlying.getRealizations())`); the CPU-typed basis functions are handed
    back to the device by type priority, so powers, regression and exercise decisions still run as kernels."""

    ESTIMATE_COND_EXPECTATION, UPPER_BOUND_METHOD = 0, 1

    def __init__(self, exerciseDates, notionals, strikes, exerciseMethod=0, numberOfBasisFunctions=5, intrinsicValueAsBasisFunction=False, useBinning=False):
        if numberOfBasisFunctions <= 0:
            raise ValueError("The vaue of numberOfBasisFunctions must be larger or equal 1. %s" % numberOfBasisFunctions)
        if exerciseMethod != self.ESTIMATE_COND_EXPECTATION:
            raise NotImplementedError("UPPER_BOUND_METHOD (golden-section search over a martingale) is outside the hot path")
        self.exerciseDates, self.notionals, self.strikes = list(exerciseDates), list(notionals), list(strikes)
        self.numberOfBasisFunctions, self.intrinsicValueAsBasisFunction, self.useBinning = numberOfBasisFunctions, intrinsicValueAsBasisFunction, useBinning
        self.lastValuationExerciseTime = None
        self.lastRegressions = []

    def getExerciseDates(self): return self.exerciseDates    # BermudanOption.java:332-360
    def getNotionals(self): return self.notionals
    def getStrikes(self): return self.strikes
    def getLastValuationExerciseTime(self): return self.lastValuationExerciseTime

    def getValueRV(self, evaluationTime, model):
        self.lastRegressions = []
        value = model.getRandomVariableForConstant(0.0)
        exerciseTime = model.getRandomVariableForConstant(self.exerciseDates[-1] + 1)
        for e in range(len(self.exerciseDates) - 1, -1, -1):
            exerciseDate, notional, strike = float(self.exerciseDates[e]), self.notionals[e], self.strikes[e]
            underlyingAtExercise = model.getAssetValue(exerciseDate, 0)
            numeraireAtPayment = model.getNumeraire(exerciseDate)
            monteCarloWeights = model.getMonteCarloWeights(exerciseDate)
            valueOfPaymentsIfExercised = underlyingAtExercise.sub(strike).mult(notional).div(numeraireAtPayment).mult(monteCarloWeights)
            basisUnderlying = underlyingAtExercise.sub(strike).floor(0.0) if self.intrinsicValueAsBasisFunction else underlyingAtExercise
            basis = self._binning(basisUnderlying) if self.useBinning else self._polynomials(basisUnderlying)
            estimator = MonteCarloConditionalExpectationRegression(basis)
            valueIfNotExcercisedEstimated = value.getConditionalExpectation(estimator)
            self.lastRegressions.append(estimator)
            exerciseCriteria = valueIfNotExcercisedEstimated.sub(valueOfPaymentsIfExercised)
            value = exerciseCriteria.choose(value, valueOfPaymentsIfExercised)
            exerciseTime = exerciseCriteria.choose(exerciseTime, Scalar(exerciseDate))
        self.lastValuationExerciseTime = exerciseTime
        return value.mult(model.getNumeraire(float(evaluationTime))).div(model.getMonteCarloWeights(float(evaluationTime)))

    def _polynomials(self, underlying):                       # :292-306
        u = RandomVariableFromDoubleArray(0.0, underlying.getRealizations())
        return [u.pow(float(k)) for k in range(self.numberOfBasisFunctions)]

    def _binning(self, underlying):                           # :308-326
        import numpy as np
        u = RandomVariableFromDoubleArray(0.0, underlying.getRealizations())
        values = np.sort(u.getRealizations())
        n = self.numberOfBasisFunctions
        out = []
        for i in range(n):
            binLeft = float(values[int(float(i) / float(n) * values.size)])
            out.append(u.sub(binLeft).choose(RandomVariableFromDoubleArray(1.0), RandomVariableFromDoubleArray(0.0)))
        return out
