"""Host-side mirror of the reference's RandomVariable seam on top of the C ABI.

* ``RandomVariable``            — the interface, J/stochastic/RandomVariable.java:28-689 (Java method names kept).
* ``Scalar``                    — J/stochastic/Scalar.java:21-379, type priority 0 (host only).
* ``RandomVariableFromDoubleArray`` — the reference's CPU type, priority 1.  Only what unchanged product code needs when
  it instantiates it directly (BermudanSwaption.java:220 builds ``new RandomVariableFromDoubleArray(1.0)``): a
  deterministic value or a host array that is handed to the GPU type.  It performs NO vector arithmetic itself.
* ``RandomVariableCuda``        — type priority 2: wins every mixed operation against the two CPU types and stays below
  the AAD wrapper (priority 3), RandomVariable.java:45-51.  Deterministic values stay host scalars; stochastic values are
  device vectors behind an fmb handle; every operation is one CUDA kernel (include/finmath_b200.h).
* ``RandomVariableCudaFactory`` — J/montecarlo/RandomVariableFactory.java:30-95.

Operation semantics (deterministic branches, rounding order, filtration time = max of operands) follow
J/montecarlo/RandomVariableFromDoubleArray.java line by line; citations are on the methods.
"""
import math

import numpy as np

from . import native as nv
from .sharding import LOCAL

NEG_INF = float("-inf")


def _jmin(a, b):
    """Java Math.min (NaN-propagating, -0.0 < +0.0)."""
    if a != a:
        return a
    if a == 0.0 and b == 0.0 and math.copysign(1.0, b) < 0:
        return b
    return a if a <= b else b


def _jmax(a, b):
    if a != a:
        return a
    if a == 0.0 and b == 0.0 and math.copysign(1.0, a) < 0:
        return b
    return a if a >= b else b


def _jdiv(a, b):
    """IEEE division like Java (no ZeroDivisionError)."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return float("nan")
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return float("-inf") if neg else float("inf")


def _jpow(a, b):
    try:
        return math.pow(a, b)
    except (OverflowError, ValueError):
        return float(np.power(np.float64(a), np.float64(b)))


def _jlog(a):
    if a > 0:
        return math.log(a)
    return float("-inf") if a == 0 else float("nan")


def _jexp(a):
    try:
        return math.exp(a)
    except OverflowError:
        return float("inf")


_UNARY_HOST = {
    nv.U_SQUARED: lambda x, a: x * x, nv.U_SQRT: lambda x, a: math.sqrt(x) if x >= 0 else float("nan"), nv.U_EXP: lambda x, a: _jexp(x),
    nv.U_LOG: lambda x, a: _jlog(x), nv.U_SIN: lambda x, a: math.sin(x), nv.U_COS: lambda x, a: math.cos(x),
    nv.U_INVERT: lambda x, a: _jdiv(1.0, x), nv.U_ABS: lambda x, a: abs(x), nv.U_ISNAN: lambda x, a: 1.0 if x != x else 0.0,
    nv.U_EXPM1: lambda x, a: math.expm1(x), nv.U_ADD: lambda x, a: x + a, nv.U_SUB: lambda x, a: x - a, nv.U_BUS: lambda x, a: a - x,
    nv.U_MULT: lambda x, a: x * a, nv.U_DIV: lambda x, a: _jdiv(x, a), nv.U_VID: lambda x, a: _jdiv(a, x), nv.U_CAP: _jmin,
    nv.U_FLOOR: _jmax, nv.U_POW: _jpow,
}


class RandomVariable:
    """Interface marker (J/stochastic/RandomVariable.java:28)."""

    def getTypePriority(self):
        raise NotImplementedError

    # default methods of the interface (RandomVariable.java:520-689)
    def bus_double(self, value):
        return self.mult(-1.0).add(value)

    def addSumProduct(self, factor1, factor2):
        result = self
        for f1, f2 in zip(factor1, factor2):
            result = result.addProduct(f1, f2)
        return result

    def getConditionalExpectation(self, estimator):
        return estimator.getConditionalExpectation(self)

    def appy(self, operator):                                # RandomVariable.java:316-318 (sic): a RandomVariable -> RandomVariable function
        return operator(self)

    def expm1(self):                                         # RandomVariable.java:502-504
        return self.exp().sub(1.0)

    def getValues(self):                                     # RandomVariable.java:62-64: "this" unless the type wraps inner values (AAD)
        return self

    def expectation(self):                                   # RandomVariable.java:466-468
        return self.average()

    def variance(self):
        return self.squared().average().sub(self.average().squared())

    def covariance(self, other):
        return self.mult(other).average().sub(self.average().mult(other.average()))


def _is_number(x):
    t = type(x)
    return t is float or t is int or isinstance(x, (int, float, np.floating, np.integer))


class Scalar(RandomVariable):
    """J/stochastic/Scalar.java — deterministic, priority 0, filtration time -inf."""
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = float(value)

    def getTypePriority(self):
        return 0

    def getFiltrationTime(self):
        return NEG_INF

    def isDeterministic(self):
        return True

    def size(self):
        return 1

    def get(self, i):
        return self.value

    def doubleValue(self):
        return self.value

    def getOperator(self):                                   # Scalar.java:88-95: both null
        return None

    def getRealizationsStream(self):
        return None

    def getHistogram(self, *args):                           # Scalar.java:168-175: UnsupportedOperationException
        raise NotImplementedError("getHistogram of a Scalar")

    def apply(self, operator, *arguments):                   # Scalar.java:183-197: the unary form only; the others return null
        return Scalar(operator(self.value)) if not arguments else None

    def getRealizations(self):
        return None                                       # Scalar.java:83-85

    def equals(self, rv):
        return rv.isDeterministic() and self.value == rv.get(0)

    def getMin(self):
        return self.value

    def getMax(self):
        return self.value

    def getAverage(self, probabilities=None):
        return self.value if probabilities is None else self.value * probabilities.getAverage()

    def getVariance(self, probabilities=None):
        return 0.0

    def getSampleVariance(self):
        return 0.0

    def getStandardDeviation(self, probabilities=None):
        return 0.0

    def getStandardError(self, probabilities=None):
        return 0.0

    def getQuantile(self, quantile, probabilities=None):
        return self.value

    def getQuantileExpectation(self, a, b):
        return self.value

    def cache(self):
        return self

    def average(self):
        return self

    def _u(self, op, a=0.0):
        return Scalar(_UNARY_HOST[op](self.value, a))

    def squared(self): return self._u(nv.U_SQUARED)
    def sqrt(self): return self._u(nv.U_SQRT)
    def exp(self): return self._u(nv.U_EXP)
    def expm1(self): return self._u(nv.U_EXPM1)
    def log(self): return self._u(nv.U_LOG)
    def sin(self): return self._u(nv.U_SIN)
    def cos(self): return self._u(nv.U_COS)
    def invert(self): return self._u(nv.U_INVERT)
    def abs(self): return self._u(nv.U_ABS)
    def isNaN(self): return self._u(nv.U_ISNAN)
    def pow(self, e): return self._u(nv.U_POW, e)

    # binary: Scalar.java:276-312 — delegate to the argument with the re-ordered arithmetic
    # (the Scalar-with-Scalar shortcuts below evaluate exactly the expression the delegation would: same operations, same order)
    def add(self, x):
        if type(x) is Scalar:
            return Scalar(x.value + self.value)
        return Scalar(self.value + x) if _is_number(x) else x.add(self.value)

    def sub(self, x):
        if type(x) is Scalar:
            return Scalar((x.value - self.value) * -1.0)
        return Scalar(self.value - x) if _is_number(x) else x.sub(self.value).mult(-1.0)

    def bus(self, x):
        return Scalar(x - self.value) if _is_number(x) else x.sub(self.value)

    def mult(self, x):
        if type(x) is Scalar:
            return Scalar(x.value * self.value)
        return Scalar(self.value * x) if _is_number(x) else x.mult(self.value)

    def div(self, x):
        return Scalar(_jdiv(self.value, x)) if _is_number(x) else x.invert().mult(self.value)

    def vid(self, x):
        return Scalar(_jdiv(x, self.value)) if _is_number(x) else x.div(self.value)

    def cap(self, x):
        return Scalar(_jmin(self.value, x)) if _is_number(x) else x.cap(self.value)

    def floor(self, x):
        return Scalar(_jmax(self.value, x)) if _is_number(x) else x.floor(self.value)

    def accrue(self, rate, periodLength):                 # :315-318
        return rate.mult(periodLength * self.value).add(self.value)

    def discount(self, rate, periodLength):               # :320-328
        if self.value == 0:
            return rate.mult(0.0)
        return rate.mult(_jdiv(periodLength, self.value)).add(_jdiv(1.0, self.value)).invert()

    def choose(self, valueIfTriggerNonNegative, valueIfTriggerNegative):
        return valueIfTriggerNonNegative if self.value >= 0 else valueIfTriggerNegative

    def addProduct(self, factor1, factor2):               # :349-357
        return factor1.mult(factor2).add(self.value)

    def addRatio(self, numerator, denominator):
        return numerator.div(denominator).add(self.value)

    def subRatio(self, numerator, denominator):
        return numerator.div(denominator).sub(self.value).mult(-1.0)

    def __repr__(self):
        return "Scalar(%r)" % self.value


class RandomVariableFromDoubleArray(RandomVariable):
    """The reference's CPU type (priority 1) as far as unchanged callers construct it directly; see the module docstring."""

    def __init__(self, time_or_value, value=None):
        if value is None:
            self.time, value = NEG_INF, time_or_value       # RandomVariableFromDoubleArray(double) :76-78
        else:
            self.time = float(time_or_value)
        if _is_number(value):
            self.valueIfNonStochastic, self.realizations = float(value), None
        else:
            self.valueIfNonStochastic, self.realizations = float("nan"), np.ascontiguousarray(value, dtype=np.float64)

    def getTypePriority(self):
        return 1

    def getFiltrationTime(self):
        return self.time

    def isDeterministic(self):
        return self.realizations is None

    def size(self):
        return 1 if self.realizations is None else self.realizations.size

    def get(self, i):
        return self.valueIfNonStochastic if self.realizations is None else float(self.realizations[i])

    def doubleValue(self):
        if self.realizations is None:
            return self.valueIfNonStochastic
        raise NotImplementedError("The random variable is non-deterministic")

    def getRealizations(self):
        return np.array([self.valueIfNonStochastic]) if self.realizations is None else self.realizations.copy()

    def getAverage(self, probabilities=None):
        if self.realizations is None:
            return self.valueIfNonStochastic if probabilities is None else self.valueIfNonStochastic * probabilities.getAverage()
        return self._gpu().getAverage(probabilities)

    def _gpu(self):
        g = self.__dict__.get("_gpu_twin")
        if g is None:                                       # uploaded once, then reused (the array is immutable by contract)
            g = RandomVariableCuda(self.time, self.valueIfNonStochastic if self.realizations is None else self.realizations)
            self.__dict__["_gpu_twin"] = g
        return g

    def __getattr__(self, name):
        # any arithmetic on the CPU type is routed to the GPU type (priority 2 wins; there is no CPU arithmetic here)
        if name.startswith("_"):
            raise AttributeError(name)
        gpu = self._gpu()
        return getattr(gpu, name)


def _unwrap(rv):
    """The innermost values of a wrapping RandomVariable (getValues() chain; the identity for plain types)."""
    inner = rv.getValues()
    while inner is not rv:
        rv, inner = inner, inner.getValues()
    return rv


class RandomVariableCuda(nv._F.RV, RandomVariable):
    """Device-resident RandomVariable, type priority 2.

    The five attributes (time, valueIfNonStochastic, shard, nGlobal, dv) live in the C base type, which also implements the common
    cases of the hot operations (_fast_unary / _fast_binary / _fast_ternary: stochastic op number, stochastic op stochastic, with the
    same filtration-time rules as the methods below); they return NotImplemented for everything else and the method carries on.
    ``dv`` is a DeviceVector or a LazyVector (element-wise operations recorded for one fused evaluation: native.py); both offer h and n."""

    def __init__(self, time, value, shard=None, _dv=None, _n=None):
        self.time = float(time)
        self.shard = shard if shard is not None else LOCAL
        self.nGlobal = _n                                  # logical (all-shard) number of paths; known without a collective
        if _dv is not None:
            self.dv, self.valueIfNonStochastic = _dv, float("nan")
            if _n is None:
                self.nGlobal = self.shard.global_count(_dv.n)
        elif _is_number(value):
            self.dv, self.valueIfNonStochastic = None, float(value)
        else:
            a = np.ascontiguousarray(value, dtype=np.float64)
            lo, hi = self.shard.local_range(a.size)
            self.dv, self.valueIfNonStochastic = nv.DeviceVector.upload(a[lo:hi]), float("nan")
            self.nGlobal = a.size

    # ---- accessors -------------------------------------------------------------------------------------------
    def getTypePriority(self):
        return 2

    def getFiltrationTime(self):
        return self.time

    def isDeterministic(self):
        return self.dv is None

    def size(self):                                        # :246-252
        return 1 if self.dv is None else self.nGlobal

    def get(self, i):
        if self.dv is None:
            return self.valueIfNonStochastic
        return self.shard.get_element(self.dv, int(i))

    def getRealizations(self):                             # :602-611 — a copy
        if self.dv is None:
            return np.array([self.valueIfNonStochastic])
        return self.shard.gather(self.dv.download())

    def getLocalRealizations(self):
        return np.array([self.valueIfNonStochastic]) if self.dv is None else self.dv.download()

    def doubleValue(self):                                 # :613-624
        if self.dv is None:
            return self.valueIfNonStochastic
        if self.size() == 1:
            return self.getAverage()
        raise NotImplementedError("The random variable is non-deterministic")     # UnsupportedOperationException

    def equals(self, rv):                                  # :188-208
        if self.time != rv.getFiltrationTime():
            return False
        if self.isDeterministic() and rv.isDeterministic():
            return self.valueIfNonStochastic == rv.doubleValue()
        if self.isDeterministic() != rv.isDeterministic():
            return False
        return bool(np.array_equal(self.getRealizations(), rv.getRealizations()))

    def cache(self):
        return self

    def apply(self, *args):                                # :646-730 take Java lambdas; T/montecarlo/RandomVariableTest.java:247-249 accepts this
        raise NotImplementedError("apply(lambda) cannot run on the device")

    def getOperator(self):
        raise NotImplementedError("getOperator() is not available for device-resident values")

    def getRealizationsStream(self):
        raise NotImplementedError("getRealizationsStream() is not available for device-resident values")

    # ---- helpers ----------------------------------------------------------------------------------------------
    def _new(self, time, dv, n=None):
        return RandomVariableCuda(time, None, self.shard, _dv=dv, _n=n if n is not None else self._n_of(dv))

    def _n_of(self, dv):
        """Logical length of a result: all stochastic operands of one operation are shards of equally long logical vectors."""
        if self.nGlobal is not None and self.dv is not None and self.dv.n == dv.n:
            return self.nGlobal
        return self._pending_n

    def _det(self, time, v):
        return RandomVariableCuda(time, v, self.shard)

    _pending_n = None

    def _operand(self, rv):
        """(device vector or None, scalar value) of any RandomVariable."""
        if rv.getTypePriority() > 2:                        # a wrapper (AAD) met where no delegation applies (choose): its inner values, no copy
            rv = _unwrap(rv)
        if isinstance(rv, RandomVariableCuda):
            if rv.dv is not None:
                self._pending_n = rv.nGlobal
            return rv.dv, rv.valueIfNonStochastic
        if rv.isDeterministic():
            return None, rv.doubleValue()
        up = RandomVariableCuda(rv.getFiltrationTime(), rv.getRealizations(), self.shard)
        self._pending_n = up.nGlobal
        return up.dv, float("nan")

    def _map1(self, op, a=0.0):
        if self.dv is None:
            return self._det(self.time, _UNARY_HOST[op](self.valueIfNonStochastic, a))
        return self._new(self.time, nv.unary(op, self.dv, a))

    def _tmax(self, *rvs):
        t = self.time
        for r in rvs:
            t = max(t, r.getFiltrationTime())
        return t

    # ---- unary and rv∘double: :742-1020 -------------------------------------------------------------------------
    def _map1f(self, op, a=0.0):
        r = self._fast_unary(op, a)
        return r if r is not NotImplemented else self._map1(op, a)

    def squared(self): return self._map1f(nv.U_SQUARED)
    def sqrt(self): return self._map1f(nv.U_SQRT)
    def exp(self): return self._map1f(nv.U_EXP)
    def expm1(self): return self._map1f(nv.U_EXPM1)
    def log(self): return self._map1f(nv.U_LOG)
    def sin(self): return self._map1f(nv.U_SIN)
    def cos(self): return self._map1f(nv.U_COS)
    def invert(self): return self._map1f(nv.U_INVERT)
    def abs(self): return self._map1f(nv.U_ABS)
    def isNaN(self): return self._map1f(nv.U_ISNAN)
    def pow(self, exponent): return self._map1f(nv.U_POW, exponent)

    def average(self):                                     # :877-880
        return RandomVariableCuda(NEG_INF, self.getAverage(), self.shard)

    # ---- binary: :1027-1276 ------------------------------------------------------------------------------------
    def _bin(self, op, rv, host, b_det_shortcut):
        t = self._tmax(rv)
        ydv, yv = self._operand(rv)
        if self.dv is None and ydv is None:
            return self._det(t, host(self.valueIfNonStochastic, yv))
        if ydv is None and b_det_shortcut is not None:
            return self._map1(b_det_shortcut, yv)          # "return this.add(randomVariable.doubleValue())": keeps this.time
        return self._new(t, nv.binary(op, self.dv, self.valueIfNonStochastic, ydv, yv))

    def add(self, x):
        r = self._fast_binary(nv.B_ADD, nv.U_ADD, x, True)
        if r is not NotImplemented:
            return r
        if _is_number(x):
            return self._map1(nv.U_ADD, x)
        if x.getTypePriority() > 2:
            return x.add(self)
        return self._bin(nv.B_ADD, x, lambda a, b: a + b, nv.U_ADD)

    def sub(self, x):
        r = self._fast_binary(nv.B_SUB, nv.U_SUB, x, True)
        if r is not NotImplemented:
            return r
        if _is_number(x):
            return self._map1(nv.U_SUB, x)
        if x.getTypePriority() > 2:
            return x.bus(self)
        return self._bin(nv.B_SUB, x, lambda a, b: a - b, nv.U_SUB)

    def bus(self, x):                                      # x - this
        if _is_number(x):
            return self._map1(nv.U_BUS, x)
        if x.getTypePriority() > 2:
            return x.sub(self)
        t = self._tmax(x)
        ydv, yv = self._operand(x)
        if self.dv is None and ydv is None:
            return self._det(t, yv - self.valueIfNonStochastic)
        return self._new(t, nv.binary(nv.B_SUB, ydv, yv, self.dv, self.valueIfNonStochastic))

    def mult(self, x):
        r = self._fast_binary(nv.B_MULT, nv.U_MULT, x, True)
        if r is not NotImplemented:
            return r
        if _is_number(x):
            return self._map1(nv.U_MULT, x)
        if x.getTypePriority() > 2:
            return x.mult(self)
        return self._bin(nv.B_MULT, x, lambda a, b: a * b, nv.U_MULT)

    def div(self, x):
        r = self._fast_binary(nv.B_DIV, nv.U_DIV, x, False)
        if r is not NotImplemented:
            return r
        if _is_number(x):
            return self._map1(nv.U_DIV, x)
        if x.getTypePriority() > 2:
            return x.vid(self)
        return self._bin(nv.B_DIV, x, _jdiv, None)

    def vid(self, x):                                      # x / this
        if _is_number(x):
            return self._map1(nv.U_VID, x)
        if x.getTypePriority() > 2:
            return x.div(self)
        t = self._tmax(x)
        ydv, yv = self._operand(x)
        if self.dv is None and ydv is None:
            return self._det(t, _jdiv(yv, self.valueIfNonStochastic))
        return self._new(t, nv.binary(nv.B_DIV, ydv, yv, self.dv, self.valueIfNonStochastic))

    def cap(self, x):
        r = self._fast_binary(nv.B_CAP, nv.U_CAP, x, False)
        if r is not NotImplemented:
            return r
        if _is_number(x):
            return self._map1(nv.U_CAP, x)
        if x.getTypePriority() > 2:
            return x.cap(self)
        return self._bin(nv.B_CAP, x, _jmin, None)

    def floor(self, x):
        r = self._fast_binary(nv.B_FLOOR, nv.U_FLOOR, x, True)
        if r is not NotImplemented:
            return r
        if _is_number(x):
            return self._map1(nv.U_FLOOR, x)
        if x.getTypePriority() > 2:
            return x.floor(self)
        return self._bin(nv.B_FLOOR, x, _jmax, nv.U_FLOOR if self.dv is not None else None)

    # ---- ternary: :1278-1479 -----------------------------------------------------------------------------------
    def accrue(self, rate, periodLength):
        r = self._fast_ternary(nv.T_ACCRUE, rate, None, periodLength)
        if r is not NotImplemented:
            return r
        if rate.getTypePriority() > 2:
            return rate.mult(periodLength).add(1.0).mult(self)
        rdv, rval = self._operand(rate)
        if rdv is None:
            return self.mult(1.0 + rval * periodLength)
        return self._new(self._tmax(rate), nv.ternary(nv.T_ACCRUE, self.dv, self.valueIfNonStochastic, rdv, rval, None, 0.0, periodLength))

    def discount(self, rate, periodLength):
        r = self._fast_ternary(nv.T_DISCOUNT, rate, None, periodLength)
        if r is not NotImplemented:
            return r
        if rate.getTypePriority() > 2:
            return rate.mult(periodLength).add(1.0).invert().mult(self)
        rdv, rval = self._operand(rate)
        if rdv is None:
            return self.div(1.0 + rval * periodLength)
        return self._new(self._tmax(rate), nv.ternary(nv.T_DISCOUNT, self.dv, self.valueIfNonStochastic, rdv, rval, None, 0.0, periodLength))

    def choose(self, valueIfTriggerNonNegative, valueIfTriggerNegative):
        r = self._fast_ternary(nv.T_CHOOSE, valueIfTriggerNonNegative, valueIfTriggerNegative, 0.0)
        if r is not NotImplemented:
            return r
        if self.dv is None:
            return valueIfTriggerNonNegative if self.valueIfNonStochastic >= 0 else valueIfTriggerNegative
        t = self._tmax(valueIfTriggerNonNegative, valueIfTriggerNegative)
        adv, av = self._operand(valueIfTriggerNonNegative)
        bdv, bv = self._operand(valueIfTriggerNegative)
        return self._new(t, nv.ternary(nv.T_CHOOSE, self.dv, 0.0, adv, av, bdv, bv))

    def addProduct(self, factor1, factor2):
        r = self._fast_ternary(nv.T_ADD_PRODUCT_D, factor1, None, factor2) if type(factor2) is float else self._fast_ternary(nv.T_ADD_PRODUCT, factor1, factor2, 0.0)
        if r is not NotImplemented:
            return r
        if _is_number(factor2):                            # addProduct(RandomVariable, double) :1365-1391
            if factor1.getTypePriority() > 2:
                return factor1.mult(factor2).add(self)
            fdv, fv = self._operand(factor1)
            if fdv is None:
                return self.add(fv * factor2)
            return self._new(self._tmax(factor1), nv.ternary(nv.T_ADD_PRODUCT_D, self.dv, self.valueIfNonStochastic, fdv, fv, None, 0.0, factor2))
        if factor1.getTypePriority() > 2 or factor2.getTypePriority() > 2:
            return factor1.mult(factor2).add(self)
        t = self._tmax(factor1, factor2)
        d1, d2 = factor1.isDeterministic(), factor2.isDeterministic()
        if self.dv is None and d1 and d2:
            return self._det(t, self.valueIfNonStochastic + (factor1.doubleValue() * factor2.doubleValue()))
        if d1 and d2:
            return self.add(factor1.doubleValue() * factor2.doubleValue())
        if d2:
            return self.addProduct(factor1, factor2.doubleValue())
        if d1:
            return self.addProduct(factor2, factor1.doubleValue())
        if self.dv is not None:
            f1, _ = self._operand(factor1)
            f2, _ = self._operand(factor2)
            return self._new(t, nv.ternary(nv.T_ADD_PRODUCT, self.dv, 0.0, f1, 0.0, f2, 0.0))
        return self.add(factor1.mult(factor2))

    def addRatio(self, numerator, denominator):
        if numerator.getTypePriority() > 2 or denominator.getTypePriority() > 2:
            return numerator.div(denominator).add(self)
        t = self._tmax(numerator, denominator)
        ndv, nval = self._operand(numerator)
        ddv, dval = self._operand(denominator)
        if self.dv is None and ndv is None and ddv is None:
            return self._det(t, self.valueIfNonStochastic + _jdiv(nval, dval))
        return self._new(t, nv.ternary(nv.T_ADD_RATIO, self.dv, self.valueIfNonStochastic, ndv, nval, ddv, dval))

    def subRatio(self, numerator, denominator):
        if numerator.getTypePriority() > 2 or denominator.getTypePriority() > 2:
            return numerator.div(denominator).mult(-1).add(self)
        t = self._tmax(numerator, denominator)
        ndv, nval = self._operand(numerator)
        ddv, dval = self._operand(denominator)
        if self.dv is None and ndv is None and ddv is None:
            return self._det(t, self.valueIfNonStochastic - _jdiv(nval, dval))
        return self._new(t, nv.ternary(nv.T_SUB_RATIO, self.dv, self.valueIfNonStochastic, ndv, nval, ddv, dval))

    # ---- reductions: :262-575 ----------------------------------------------------------------------------------
    def _sum(self, op, w=None, a=0.0):
        hi, lo = nv.reduce(op, self.dv, w, a)
        hi, lo = self.shard.sum_dd(hi, lo)
        return hi + lo

    def getAverage(self, probabilities=None):
        if probabilities is not None:
            if self.dv is None:
                return self.valueIfNonStochastic * probabilities.getAverage()
            if self.size() == 0:
                return float("nan")
            wdv, wv = self._operand(probabilities)
            if wdv is None:
                return self.mult(wv).getAverage()
            return self._sum(nv.R_SUM_PRODUCT, wdv) / self.size()
        if self.dv is None:
            return self.valueIfNonStochastic
        n = self.size()
        if n == 0:
            return float("nan")
        return self._sum(nv.R_SUM) / n

    def getVariance(self, probabilities=None):
        if probabilities is not None:                      # :357-380 — quirk: not divided by n
            if self.dv is None:
                return 0.0
            if self.size() == 0:
                return float("nan")
            average = self.getAverage(probabilities)
            wdv, wv = self._operand(probabilities)
            if wdv is None:
                return self.sub(average).squared().mult(wv)._sum(nv.R_SUM)
            return self._sum(nv.R_CENTERED_M2_W, wdv, average)
        if self.dv is None or self.size() == 1:
            return 0.0
        if self.size() == 0:
            return float("nan")
        average = self.getAverage()
        return self._sum(nv.R_CENTERED_M2, None, average) / self.size()

    def getSampleVariance(self):
        if self.dv is None or self.size() == 1:
            return 0.0
        if self.size() == 0:
            return float("nan")
        return self.getVariance() * self.size() / (self.size() - 1)

    def getStandardDeviation(self, probabilities=None):
        if self.dv is None:
            return 0.0
        if self.size() == 0:
            return float("nan")
        return math.sqrt(self.getVariance(probabilities))

    def getStandardError(self, probabilities=None):
        if self.dv is None:
            return 0.0
        if self.size() == 0:
            return float("nan")
        return self.getStandardDeviation(probabilities) / math.sqrt(self.size())

    def getMin(self):
        if self.dv is None:
            return self.valueIfNonStochastic
        if self.dv.n == 0 and self.size() == 0:
            return float(np.finfo(np.float64).max)         # Double.MAX_VALUE :262-264
        return self.shard.min(nv.reduce(nv.R_MIN, self.dv)[0], self.dv.n > 0)

    def getMax(self):
        if self.dv is None:
            return self.valueIfNonStochastic
        if self.dv.n == 0 and self.size() == 0:
            return -float(np.finfo(np.float64).max)
        return self.shard.max(nv.reduce(nv.R_MAX, self.dv)[0], self.dv.n > 0)

    @staticmethod
    def _quantile_index(n, q):
        return min(max(int(math.floor((n + 1) * q - 1 + 0.5)), 0), n - 1)      # Math.round, :454

    def _select(self, rank):
        """Element of the given rank of the sorted logical vector: radix select on the device (no sort; with shards only 256-bin
        histograms are exchanged, or - host-side exchange - the shards are gathered)."""
        import ctypes as C
        dv = self.dv if (self.shard.world == 1 or self.shard.native_comm) else nv.DeviceVector.upload(self.getRealizations())
        out = C.c_double()
        nv.check(nv.load().fmb_rv_select(dv.h, int(rank), C.byref(out)))
        return out.value

    def getQuantile(self, quantile, probabilities=None):
        if self.dv is None:
            return self.valueIfNonStochastic
        if self.size() == 0:
            return float("nan")
        if probabilities is not None:
            raise RuntimeError("Method not implemented.")   # :471
        return self._select(self._quantile_index(self.size(), quantile))

    def getQuantileExpectation(self, quantileStart, quantileEnd):
        if self.dv is None:
            return self.valueIfNonStochastic
        if self.size() == 0:
            return float("nan")
        if quantileStart > quantileEnd:
            return self.getQuantileExpectation(quantileEnd, quantileStart)
        n = self.size()
        i0, i1 = self._quantile_index(n, quantileStart), self._quantile_index(n, quantileEnd)
        # mean of sorted[i0 .. i1] (:484-497) from two selected values and ONE pass: the elements strictly between them, plus the copies of
        # the two end values that fall inside the index range (the reference adds the sorted values one by one; here the sum is double-double)
        v0, v1 = self._select(i0), self._select(i1)
        if v0 == v1 or (v0 != v0 and v1 != v1):
            return v0
        import ctypes as C
        dv = self.dv if (self.shard.world == 1 or self.shard.native_comm) else nv.DeviceVector.upload(self.getRealizations())
        out = (C.c_double * 4)()
        nv.check(nv.load().fmb_rv_range_sum(dv.h, v0, v1, out))
        inner, count_le_v0, count_lt_v1 = out[0] + out[1], int(out[2]), int(out[3])
        copies0 = count_le_v0 - i0                            # copies of v0 at ranks >= i0
        copies1 = i1 - count_lt_v1 + 1                        # copies of v1 at ranks <= i1
        return (inner + copies0 * v0 + copies1 * v1) / (i1 - i0 + 1)

    def getHistogram(self, intervalPoints, standardDeviations=None):
        import ctypes as C
        if standardDeviations is not None:                 # getHistogram(int, double) :553-575
            numberOfPoints = int(intervalPoints)
            pts = np.empty(numberOfPoints)
            anchors = np.empty(numberOfPoints + 1)
            center = self.getAverage()
            radius = standardDeviations * self.getStandardDeviation()
            stepSize = (numberOfPoints - 1) / 2.0
            for i in range(numberOfPoints):
                alpha = (-(numberOfPoints - 1) / 2.0 + i) / stepSize
                pts[i] = center + alpha * radius
                anchors[i] = center + alpha * radius - radius / (2 * stepSize)
            anchors[numberOfPoints] = center + 1 * radius + radius / (2 * stepSize)
            return [anchors, self.getHistogram(pts)]
        pts = nv.as_f64(intervalPoints)
        h = np.zeros(pts.size + 1)
        if self.dv is None:                                # :505-517
            for k in range(pts.size):
                if self.valueIfNonStochastic > pts[k]:
                    h[k] = 1.0
                    break
            h[pts.size] = 1.0
            return h
        n = self.size()
        counts = np.zeros(pts.size, dtype=np.uint64)
        native_all = self.shard.world == 1 or self.shard.native_comm
        for k0 in range(0, pts.size, 500):                  # (at most 511 thresholds per native call)
            chunk = np.ascontiguousarray(pts[k0:k0 + 500])
            part = np.zeros(chunk.size, dtype=np.uint64)
            nv.check(nv.load().fmb_rv_count_le(self.dv.h, nv.dptr(chunk), chunk.size, part.ctypes.data_as(nv.c_hp)))
            if not native_all:
                part = np.array([int(v) for v in self.shard._all_gather([float(v) for v in part]).sum(axis=0)], dtype=np.uint64)
            counts[k0:k0 + 500] = part
        prev = 0
        for k in range(pts.size):                          # left-open / right-closed buckets in sorted order :528-550
            c = max(int(counts[k]), prev)
            h[k] = c - prev
            prev = c
        h[pts.size] = n - prev
        if n > 0:
            h /= n
        return h

    def __repr__(self):
        if self.dv is None:
            return "RandomVariableCuda(time=%r, value=%r)" % (self.time, self.valueIfNonStochastic)
        return "RandomVariableCuda(time=%r, paths=%d)" % (self.time, self.size())


class RandomVariableCudaFactory:
    """RandomVariableFactory (J/montecarlo/RandomVariableFactory.java:30-95) producing RandomVariableCuda."""

    def __init__(self, shard=None):
        self.shard = shard if shard is not None else LOCAL

    def createRandomVariable(self, *args):
        if len(args) == 1:
            v = args[0]
            if _is_number(v):
                return RandomVariableCuda(NEG_INF, v, self.shard)
            return RandomVariableCuda(NEG_INF, v, self.shard)     # createRandomVariable(double[]) -> time -inf
        time, v = args
        return RandomVariableCuda(time, v, self.shard)

    def createRandomVariableArray(self, values):
        return [self.createRandomVariable(NEG_INF, float(v)) for v in values]

    def createRandomVariableMatrix(self, values):
        return [[self.createRandomVariable(NEG_INF, float(v)) for v in row] for row in values]

    def fromDevice(self, time, dv, n=None):
        return RandomVariableCuda(time, None, self.shard, _dv=dv, _n=n)

    @staticmethod
    def getRandomVariableOrDefault(randomVariableFactory, value, defaultValue):    # RandomVariableFactory.java:32-48
        if value is None:
            return defaultValue
        if isinstance(value, RandomVariable):
            return value
        if _is_number(value):
            if randomVariableFactory is None:
                raise TypeError("Object value of type Number but nor randomVariableFactory given.")          # NullPointerException there
            return randomVariableFactory.createRandomVariable(float(value))
        raise ValueError("Object value must be of type Number or RandomVariable.")


nv._bind_fast()       # (native.load() may have run before this module existed: give the accelerator the class it instantiates)
