"""Backward-mode (adjoint) stochastic automatic differentiation ON TOP of the device RandomVariable — SURVEY.md §8f rank 4.

Mirrors the reference's consumer of the factory seam:

* ``RandomVariableDifferentiableAAD``        — J/montecarlo/automaticdifferentiation/backward/RandomVariableDifferentiableAAD.java
  (type priority 3, :63; operator tree node :79-160; adjoint update rule :165-215; partial derivatives :226-434; Dirac-delta
  regression :436-541; backward sweep ``getGradient`` :660-706; the differentiable operations :835-1290)
* ``RandomVariableDifferentiableAADFactory`` — …/backward/RandomVariableDifferentiableAADFactory.java:25-120 (property names and defaults)
* ``RandomVariableDifferentiable``           — …/RandomVariableDifferentiable.java:22-105 (getID / getGradient / getTangents / getCloneIndependent)

What the wrapper adds is BOOK-KEEPING only: every value, every partial derivative and every adjoint is an inner RandomVariable — with
``RandomVariableDifferentiableAADFactory(RandomVariableCudaFactory())`` a device vector produced by the library's own kernels — so the
forward evaluation and the backward sweep both run on the GPU and nothing is copied to the host (the adjoint update
``D_i += D_m · ∂f_m/∂x_i`` is ONE fused add-product kernel per edge).  The operator tree is a table here: one row per operator with
(which argument values the adjoint needs, the partial derivative per argument index) instead of the reference's switch statements.

A process model whose parameters come from a differentiable factory does not describe itself to the fused Euler kernels
(``getFusedSpecification`` sees non-plain parameters and declines, models.py), so ``EulerSchemeFromProcessModel`` evolves it with the
reference's generic recipe on RandomVariable operations — each of them recorded here.
"""
import heapq
import itertools

from .stochastic import RandomVariable, RandomVariableCudaFactory, Scalar, _is_number

_nextID = itertools.count()                                  # AtomicLong indexOfNextRandomVariable (:65)

_ZERO, _ONE, _MINUS_ONE = Scalar(0.0), Scalar(1.0), Scalar(-1.0)


class DiracDeltaApproximationMethod:
    """…AADFactory.java:30-36 (the spelling of the fourth name is the reference's)."""
    DISCRETE_DELTA = "DISCRETE_DELTA"
    REGRESSION_ON_DENSITY = "REGRESSION_ON_DENSITY"
    REGRESSION_ON_DISTRIBUITON = "REGRESSION_ON_DISTRIBUITON"
    ONE = "ONE"
    ZERO = "ZERO"
    _ALL = (DISCRETE_DELTA, REGRESSION_ON_DENSITY, REGRESSION_ON_DISTRIBUITON, ONE, ZERO)

    @classmethod
    def valueOf(cls, name):
        if name not in cls._ALL:
            raise ValueError("No enum constant DiracDeltaApproximationMethod." + str(name))       # IllegalArgumentException of Enum.valueOf
        return name


# ---- the operator table ---------------------------------------------------------------------------------------------------------
# retain(args, vals): drop the argument values the adjoint never reads (:100-157 — "the simple modification which reduces memory
# requirements"; on the device every dropped value is a freed vector of 8·paths bytes).
def _keepAll(args, vals):
    return vals


def _keepNone(args, vals):
    return None


def _keepProductFactors(args, vals):
    if args[0] is None:
        vals[1] = None
    if args[1] is None:
        vals[0] = None
    return vals


def _keepQuotient(args, vals):
    if args[1] is None:
        vals[0] = None
    return vals


def _keepAddProduct(args, vals):
    vals[0] = None
    if args[1] is None:
        vals[2] = None
    if args[2] is None:
        vals[1] = None
    return vals


def _keepAccrue(args, vals):
    if args[1] is None and args[2] is None:
        vals[0] = None
    if args[0] is None and args[1] is None:
        vals[1] = None
    if args[0] is None and args[2] is None:
        vals[2] = None
    return vals


def _keepChoose(args, vals):
    if args[0] is None:
        vals[1] = None
        vals[2] = None
    return vals


def _centredSecondMoment(X, scale):
    # d/dx_i of the (biased / sample) variance as the reference writes it (:262-283)
    n = X.size()
    return X.sub(X.getAverage() * (2.0 * n - 1.0) / n).mult(scale)


def _weightedMoment(A, B):
    n = A.size()
    return B.mult(2.0).mult(A.mult(B.add(A.getAverage(B) * (n - 1)).sub(A.getAverage(B))))


def _indicatorOfExtreme(X, extreme):
    # :268-287 apply(x -> x == extreme ? 1 : 0) — without a host lambda: 1 - sign(|x - extreme|) as two chooses
    d = X.sub(extreme)
    return d.choose(d.mult(-1.0).choose(_ONE, _ZERO), _ZERO)


def _diracDelta(node, X, Y, Z):
    """∂ choose(X, Y, Z) / ∂X = (Y - Z) · δ(X) under the factory's approximation of δ (:386-425)."""
    method = node.factory.getDiracDeltaApproximationMethod()
    if method == DiracDeltaApproximationMethod.ZERO:
        return _ZERO
    if method != DiracDeltaApproximationMethod.DISCRETE_DELTA:
        return Y.sub(Z)                                     # ONE, and the two regression methods (their density enters in _propagate)
    epsilon = node.factory.getDiracDeltaApproximationWidthPerStdDev() * X.getStandardDeviation()
    if epsilon == float("inf"):
        return Y.sub(Z)
    if epsilon > 0:
        d = Y.sub(Z)
        d = d.mult(X.add(epsilon / 2).choose(_ONE, _ZERO))
        d = d.mult(X.sub(epsilon / 2).choose(_ZERO, _ONE))
        return d.div(epsilon)
    return _ZERO


# name -> (retain rule, partial derivative f(node, argumentIndex, X, Y, Z))
_OPERATORS = {
    "SQUARED":  (_keepAll, lambda n, i, X, Y, Z: X.mult(2.0)),
    "SQRT":     (_keepAll, lambda n, i, X, Y, Z: X.sqrt().invert().mult(0.5)),
    "EXP":      (_keepAll, lambda n, i, X, Y, Z: X.exp()),
    "LOG":      (_keepAll, lambda n, i, X, Y, Z: X.invert()),
    "SIN":      (_keepAll, lambda n, i, X, Y, Z: X.cos()),
    "COS":      (_keepAll, lambda n, i, X, Y, Z: X.sin().mult(-1.0)),
    "INVERT":   (_keepAll, lambda n, i, X, Y, Z: X.invert().squared().mult(-1)),
    "ABS":      (_keepAll, lambda n, i, X, Y, Z: X.choose(_ONE, _MINUS_ONE)),
    "AVERAGE":  (_keepNone, lambda n, i, X, Y, Z: _ONE),
    "CONDITIONAL_EXPECTATION": (_keepAll, lambda n, i, X, Y, Z: _ONE),
    "VARIANCE": (_keepAll, lambda n, i, X, Y, Z: _centredSecondMoment(X, 2.0 / X.size())),
    "SVARIANCE": (_keepAll, lambda n, i, X, Y, Z: _centredSecondMoment(X, 2.0 / (X.size() - 1))),
    "STDEV":    (_keepAll, lambda n, i, X, Y, Z: _centredSecondMoment(X, 2.0 / X.size()).mult(0.5).div(X.getVariance() ** 0.5)),
    "STDERROR": (_keepAll, lambda n, i, X, Y, Z: _centredSecondMoment(X, 2.0 / X.size()).mult(0.5).div((X.getVariance() * X.size()) ** 0.5)),
    "MIN":      (_keepAll, lambda n, i, X, Y, Z: _indicatorOfExtreme(X, X.getMin())),
    "MAX":      (_keepAll, lambda n, i, X, Y, Z: _indicatorOfExtreme(X, X.getMax())),
    "ADD":      (_keepNone, lambda n, i, X, Y, Z: _ONE),
    "SUB":      (_keepNone, lambda n, i, X, Y, Z: _ONE if i == 0 else _MINUS_ONE),
    "MULT":     (_keepProductFactors, lambda n, i, X, Y, Z: Y if i == 0 else X),
    "DIV":      (_keepQuotient, lambda n, i, X, Y, Z: Y.invert() if i == 0 else X.div(Y.squared()).mult(-1)),
    "CAP":      (_keepAll, lambda n, i, X, Y, Z: X.sub(Y).choose(_ZERO, _ONE) if i == 0 else X.sub(Y).choose(_ONE, _ZERO)),
    "FLOOR":    (_keepAll, lambda n, i, X, Y, Z: X.sub(Y).choose(_ONE, _ZERO) if i == 0 else X.sub(Y).choose(_ZERO, _ONE)),
    "AVERAGE2": (_keepAll, lambda n, i, X, Y, Z: Y if i == 0 else X),
    "VARIANCE2": (_keepAll, lambda n, i, X, Y, Z: _weightedMoment(X, Y) if i == 0 else _weightedMoment(Y, X)),
    "STDEV2":   (_keepAll, lambda n, i, X, Y, Z: _weightedMoment(X, Y).div(X.getVariance(Y) ** 0.5) if i == 0
                 else _weightedMoment(Y, X).div(Y.getVariance(X) ** 0.5)),
    "STDERROR2": (_keepAll, lambda n, i, X, Y, Z: _weightedMoment(X, Y).div((X.getVariance(Y) * X.size()) ** 0.5) if i == 0
                  else _weightedMoment(Y, X).div((Y.getVariance(X) * Y.size()) ** 0.5)),
    # the exponent is a constant (:346-348)
    "POW":      (_keepAll, lambda n, i, X, Y, Z: X.pow(Y.doubleValue() - 1.0).mult(Y) if i == 0 else _ZERO),
    "ADDPRODUCT": (_keepAddProduct, lambda n, i, X, Y, Z: _ONE if i == 0 else (Z if i == 1 else Y)),
    "ADDRATIO": (_keepAll, lambda n, i, X, Y, Z: _ONE if i == 0 else (Z.invert() if i == 1 else Y.div(Z.squared()).mult(-1.0))),
    "SUBRATIO": (_keepAll, lambda n, i, X, Y, Z: _ONE if i == 0 else (Z.invert().mult(-1.0) if i == 1 else Y.div(Z.squared()))),
    "ACCRUE":   (_keepAccrue, lambda n, i, X, Y, Z: Y.mult(Z).add(1.0) if i == 0 else (X.mult(Z) if i == 1 else X.mult(Y))),
    "DISCOUNT": (_keepAll, lambda n, i, X, Y, Z: Y.mult(Z).add(1.0).invert() if i == 0
                 else (X.mult(Z) if i == 1 else X.mult(Y)).div(Y.mult(Z).add(1.0).squared()).mult(-1.0)),
    "CHOOSE":   (_keepChoose, lambda n, i, X, Y, Z: _diracDelta(n, X, Y, Z) if i == 0
                 else (X.choose(_ONE, _ZERO) if i == 1 else X.choose(_ZERO, _ONE))),
}


class _OperatorTreeNode:
    """A vertex of the operator tree: id m, operator f_m, argument vertices and (only the needed) argument values (:79-160)."""
    __slots__ = ("id", "operatorType", "arguments", "argumentValues", "operator", "factory")

    def __init__(self, operatorType, arguments, argumentValues, operator, factory):
        self.id = next(_nextID)
        self.operatorType = operatorType
        self.arguments = arguments
        self.operator = operator
        self.factory = factory
        if operatorType is not None and argumentValues is not None:
            argumentValues = _OPERATORS[operatorType][0](arguments, list(argumentValues))
        self.argumentValues = argumentValues

    @staticmethod
    def of(randomVariable):
        return randomVariable.getOperatorTreeNode() if isinstance(randomVariable, RandomVariableDifferentiableAAD) else None

    def partialDerivative(self, argumentIndex):
        v = self.argumentValues
        n = len(self.arguments)
        X = v[0] if v is not None and n > 0 else None
        Y = v[1] if v is not None and n > 1 else None
        Z = v[2] if v is not None and n > 2 else None
        return _OPERATORS[self.operatorType][1](self, argumentIndex, X, Y, Z)

    def propagate(self, derivatives):
        """D_i += D_m · ∂f_m/∂x_i for the arguments i of this vertex m (:165-215)."""
        for argumentIndex, argument in enumerate(self.arguments):
            if argument is None:
                continue                                    # a (non-differentiable) constant
            partial = self.partialDerivative(argumentIndex)
            derivative = derivatives[self.id]
            op = self.operatorType
            if op == "AVERAGE":                             # the adjoint of E is E (ssrn 2995695)
                derivative = derivative.average()
            elif op == "CONDITIONAL_EXPECTATION":           # the adjoint of E(·|F) is E(·|F) with the same estimator
                derivative = self.operator.getConditionalExpectation(derivative)
            elif op == "CHOOSE" and argumentIndex == 0 and self.factory.getDiracDeltaApproximationMethod() in (
                    DiracDeltaApproximationMethod.REGRESSION_ON_DENSITY, DiracDeltaApproximationMethod.REGRESSION_ON_DISTRIBUITON):
                derivative = self._diracDeltaRegression(derivative, self.argumentValues[0])
            previous = derivatives.get(argument.id)
            derivatives[argument.id] = derivative.mult(partial) if previous is None else previous.addProduct(partial, derivative)

    # ---- indicator functions by regression (ssrn 3282667; :436-541) ------------------------------------------------------------
    def _diracDeltaRegression(self, derivative, indicator):
        epsilon = self.factory.getDiracDeltaApproximationWidthPerStdDev() * indicator.getStandardDeviation()
        localizedOne = indicator.add(epsilon / 2).choose(_ONE, _ZERO).mult(indicator.sub(epsilon / 2).choose(_ZERO, _ONE))
        derivative = derivative.mult(localizedOne).div(localizedOne.getAverage())
        return derivative.mult(self._densityRegression(indicator))

    def _densityRegression(self, indicator):
        from .montecarlo import LinearRegression
        method = self.factory.getDiracDeltaApproximationMethod()
        half = 50                                            # numberOfSamplePointsHalf (:449)
        step = self.factory.getDiracDeltaApproximationDensityRegressionWidthPerStdDev() / 2 * indicator.getStandardDeviation() / half
        xs, ys = [0.0] * (2 * half), [0.0] * (2 * half)
        positive, negative = indicator.choose(_ONE, _ZERO), indicator.choose(_ZERO, _ONE)
        onDensity = method == DiracDeltaApproximationMethod.REGRESSION_ON_DENSITY
        width = step
        for i in range(0, 2 * half, 2):
            width += step
            below = indicator.add(width).choose(_ONE, _ZERO).mult(negative).getAverage()
            above = indicator.sub(width).choose(_ZERO, _ONE).mult(positive).getAverage()
            xs[i], xs[i + 1] = -width, width
            ys[i], ys[i + 1] = (below / width, above / width) if onDensity else (-below, above)
        make = self.factory.createRandomVariableNonDifferentiable
        x, y = make(0.0, xs), make(0.0, ys)
        basis = [x.mult(0.0).add(1.0), x, x.squared()] if onDensity else [x, x.squared(), x.pow(3)]
        return LinearRegression(basis).getRegressionCoefficients(y)[0]


class RandomVariableDifferentiable(RandomVariable):
    """…/RandomVariableDifferentiable.java:22-105."""

    def getID(self):
        raise NotImplementedError

    def getGradient(self, independentIDs=None):
        raise NotImplementedError

    def getTangents(self, dependentIDs=None):
        raise NotImplementedError

    def getCloneIndependent(self):
        raise NotImplementedError("Cloning not supported.")


def _valuesOf(rv):
    return rv.getValues() if rv is not None else None


class RandomVariableDifferentiableAAD(RandomVariableDifferentiable):
    typePriorityDefault = 3

    def __init__(self, values, factory=None, _operator=None, _arguments=None, _argumentValues=None, _estimator=None):
        """Public form: (value or RandomVariable [, factory]) creates an independent (a leaf)  — :573-590.
        The differentiable operations below use the private keywords (vertex with arguments)."""
        if _is_number(values):
            values = Scalar(values)
        if factory is None:
            factory = values.getFactory() if isinstance(values, RandomVariableDifferentiableAAD) else RandomVariableDifferentiableAADFactory()
        self.values = values
        self.factory = factory
        self.operatorTreeNode = _OperatorTreeNode(_operator, _arguments, _argumentValues, _estimator, factory)

    @staticmethod
    def of(value):
        return RandomVariableDifferentiableAAD(value)

    # ---- vertex construction ----------------------------------------------------------------------------------------------------
    def _vertex(self, values, operator, arguments, estimator=None):
        """All arguments are RandomVariables; their vertices and values are extracted (:596-602)."""
        return RandomVariableDifferentiableAAD(values, self.factory, operator, [_OperatorTreeNode.of(a) for a in arguments],
                                               [_valuesOf(a) for a in arguments], estimator)

    def _vertexOf(self, values, operator, nodes, argumentValues):
        return RandomVariableDifferentiableAAD(values, self.factory, operator, nodes, argumentValues, None)

    def getOperatorTreeNode(self):
        return self.operatorTreeNode

    def getValues(self):
        return self.values

    def getFactory(self):
        return self.factory

    def getID(self):
        return self.operatorTreeNode.id

    def getTypePriority(self):
        return self.typePriorityDefault

    # ---- the backward sweep (:660-706) ------------------------------------------------------------------------------------------
    def getGradient(self, independentIDs=None):
        """Map id -> d this / d (vertex id), for the ids in ``independentIDs`` (all leaves — or, with
        isGradientRetainsLeafNodesOnly = false, all vertices — if None).

        Vertices are processed in descending id order (an argument always has a smaller id than the vertex using it), so every adjoint is
        complete when it is propagated; adjoints of inner vertices are dropped as soon as they are used — on the device that releases
        their vectors during the sweep."""
        leavesOnly = self.factory is not None and self.factory.isGradientRetainsLeafNodesOnly()
        wanted = None if independentIDs is None else set(independentIDs)
        root = self.operatorTreeNode
        derivatives = {root.id: _ONE}
        pending = {root.id: root}
        order = [-root.id]
        while order:
            vertexID = -heapq.heappop(order)
            vertex = pending.pop(vertexID)
            if vertex.arguments:
                vertex.propagate(derivatives)
                if leavesOnly:
                    derivatives.pop(vertexID, None)
                for argument in vertex.arguments:
                    if argument is not None and argument.id not in pending:
                        pending[argument.id] = argument
                        heapq.heappush(order, -argument.id)
            if wanted is not None and vertexID not in wanted:
                derivatives.pop(vertexID, None)
        return derivatives

    def getTangents(self, dependentIDs=None):
        raise NotImplementedError("getTangents")            # UnsupportedOperationException (:713-715)

    def getCloneIndependent(self):
        return RandomVariableDifferentiableAAD(self.getValues())

    # ---- end points: plain numbers, not differentiable (:722-833, :1234-1262) -------------------------------------------------
    def equals(self, randomVariable):
        return self.values.equals(randomVariable)

    def cache(self):
        self.values = self.values.cache()
        return self

    def isNaN(self):
        return self.values.isNaN()

    def apply(self, *args):
        raise NotImplementedError("Applying functions is not supported.")                         # :1249-1262

    # ---- differentiable operations ---------------------------------------------------------------------------------------------
    def _unary(self, name, operator, *a):
        return self._vertex(getattr(self.values, name)(*a), operator, [self])

    def squared(self): return self._unary("squared", "SQUARED")
    def sqrt(self): return self._unary("sqrt", "SQRT")
    def exp(self): return self._unary("exp", "EXP")
    def log(self): return self._unary("log", "LOG")
    def sin(self): return self._unary("sin", "SIN")
    def cos(self): return self._unary("cos", "COS")
    def invert(self): return self._unary("invert", "INVERT")
    def abs(self): return self._unary("abs", "ABS")
    def average(self): return self._unary("average", "AVERAGE")

    def expectation(self):
        return self.average()

    def pow(self, exponent):
        return self._vertex(self.values.pow(exponent), "POW", [self, Scalar(exponent)])

    def getConditionalExpectation(self, estimator):
        return self._vertex(self.values.getConditionalExpectation(estimator), "CONDITIONAL_EXPECTATION", [self], estimator)

    def _higher(self, *rvs):
        p = self.getTypePriority()
        for r in rvs:
            if r.getTypePriority() > p:
                return True
        return False

    def _withConstant(self, name, operator, value, retained):
        # op(double): the constant is argument 1 and carries no vertex (:849-920)
        return self._vertexOf(getattr(self.values, name)(value), operator, [self.operatorTreeNode, None], retained)

    def add(self, x):
        if _is_number(x):
            return self._withConstant("add", "ADD", x, [None, None])
        if self._higher(x):
            return x.add(self)
        return self._vertexOf(self.values.add(x.getValues()), "ADD", [self.operatorTreeNode, _OperatorTreeNode.of(x)], [None, None])

    def sub(self, x):
        if _is_number(x):
            return self._withConstant("sub", "SUB", x, [None, None])
        if self._higher(x):
            return x.bus(self)
        return self._vertexOf(self.values.sub(x.getValues()), "SUB", [self.operatorTreeNode, _OperatorTreeNode.of(x)], [None, None])

    def bus(self, x):                                        # x - this: SUB with swapped arguments (:1047-1061)
        if _is_number(x):
            return self.mult(-1.0).add(x)                    # RandomVariable.bus(double) default (J/stochastic/RandomVariable.java:528-530)
        if self._higher(x):
            return x.sub(self)
        return self._vertexOf(self.values.bus(x.getValues()), "SUB", [_OperatorTreeNode.of(x), self.operatorTreeNode], [None, None])

    def mult(self, x):
        if _is_number(x):
            return self._withConstant("mult", "MULT", x, [None, Scalar(x)])
        if self._higher(x):
            return x.mult(self)
        return self._vertex(self.values.mult(x.getValues()), "MULT", [self, x])

    def div(self, x):
        if _is_number(x):
            return self._withConstant("div", "DIV", x, [None, Scalar(x)])
        if self._higher(x):
            return x.vid(self)
        return self._vertex(self.values.div(x.getValues()), "DIV", [self, x])

    def vid(self, x):                                        # x / this: DIV with swapped arguments (:1093-1107)
        if _is_number(x):
            return self.invert().mult(x)                     # RandomVariable.vid(double) default (RandomVariable.java:558-560)
        if self._higher(x):
            return x.div(self)
        return self._vertex(self.values.vid(x.getValues()), "DIV", [x, self])

    def cap(self, x):
        if _is_number(x):
            return self._withConstant("cap", "CAP", x, [self.values, Scalar(x)])
        if self._higher(x):
            return x.cap(self)
        return self._vertex(self.values.cap(x.getValues()), "CAP", [self, x])

    def floor(self, x):
        if _is_number(x):
            return self._withConstant("floor", "FLOOR", x, [self.values, Scalar(x)])
        if self._higher(x):
            return x.floor(self)
        return self._vertex(self.values.floor(x.getValues()), "FLOOR", [self, x])

    def accrue(self, rate, periodLength):
        if self._higher(rate):
            return rate.mult(periodLength).add(1.0).mult(self)
        return self._vertex(self.values.accrue(rate.getValues(), periodLength), "ACCRUE", [self, rate, Scalar(periodLength)])

    def discount(self, rate, periodLength):
        if self._higher(rate):
            return rate.mult(periodLength).add(1.0).invert().mult(self)
        return self._vertex(self.values.discount(rate.getValues(), periodLength), "DISCOUNT", [self, rate, Scalar(periodLength)])

    def choose(self, valueIfTriggerNonNegative, valueIfTriggerNegative):
        a, b = valueIfTriggerNonNegative, valueIfTriggerNegative
        return self._vertex(self.values.choose(a.getValues(), b.getValues()), "CHOOSE", [self, a, b])

    def addProduct(self, factor1, factor2):
        if _is_number(factor2):                              # :1165-1181
            if self._higher(factor1):
                return factor1.mult(factor2).add(self)
            return self._vertexOf(self.values.addProduct(factor1.getValues(), factor2), "ADDPRODUCT",
                                  [self.operatorTreeNode, _OperatorTreeNode.of(factor1), None],
                                  [self.values, factor1.getValues(), Scalar(factor2)])
        if self._higher(factor1, factor2):
            return factor1.mult(factor2).add(self)
        return self._vertex(self.values.addProduct(factor1.getValues(), factor2.getValues()), "ADDPRODUCT", [self, factor1, factor2])

    def addRatio(self, numerator, denominator):
        if self._higher(numerator, denominator):
            return numerator.div(denominator).add(self)
        return self._vertex(self.values.addRatio(numerator.getValues(), denominator.getValues()), "ADDRATIO", [self, numerator, denominator])

    def subRatio(self, numerator, denominator):
        if self._higher(numerator, denominator):
            return numerator.div(denominator).mult(-1).add(self)
        return self._vertex(self.values.subRatio(numerator.getValues(), denominator.getValues()), "SUBRATIO", [self, numerator, denominator])

    # ---- differentiable statistics: deterministic vertices (:1264-1318) ---------------------------------------------------------
    def _statistic(self, value, operator):
        return self._vertex(Scalar(value), operator, [self])

    def getVarianceAsRandomVariableAAD(self): return self._statistic(self.getVariance(), "VARIANCE")
    def getSampleVarianceAsRandomVariableAAD(self): return self._statistic(self.getSampleVariance(), "SVARIANCE")
    def getStandardDeviationAsRandomVariableAAD(self): return self._statistic(self.getStandardDeviation(), "STDEV")
    def getStandardErrorAsRandomVariableAAD(self): return self._statistic(self.getStandardError(), "STDERROR")
    def getMinAsRandomVariableAAD(self): return self._statistic(self.getMin(), "MIN")
    def getMaxAsRandomVariableAAD(self): return self._statistic(self.getMax(), "MAX")

    def __repr__(self):
        return "RandomVariableDifferentiableAAD [values=%r,\n ID=%d]" % (self.values, self.getID())


def _endPoint(name):
    def method(self, *args):
        return getattr(self.values, name)(*args)
    method.__name__ = name
    method.__doc__ = "End point (a plain number / array of the inner values): not differentiable."
    return method


for _name in ("getFiltrationTime", "get", "size", "isDeterministic", "getRealizations", "doubleValue", "getMin", "getMax", "getAverage",
              "getVariance", "getSampleVariance", "getStandardDeviation", "getStandardError", "getQuantile", "getQuantileExpectation",
              "getHistogram", "getOperator", "getRealizationsStream"):
    setattr(RandomVariableDifferentiableAAD, _name, _endPoint(_name))
del _name


class RandomVariableDifferentiableAADFactory:
    """…/backward/RandomVariableDifferentiableAADFactory.java:25-120 over AbstractRandomVariableDifferentiableFactory.java:19-60.

    The inner factory defaults to the device factory (the reference defaults to its CPU array factory; this package has no CPU
    arithmetic)."""

    def __init__(self, randomVariableFactoryForNonDifferentiable=None, properties=None):
        if properties is None and isinstance(randomVariableFactoryForNonDifferentiable, dict):   # the (properties) constructor :64-66
            randomVariableFactoryForNonDifferentiable, properties = None, randomVariableFactoryForNonDifferentiable
        p = properties if properties is not None else {}
        self.randomVariableFactoryForNonDifferentiable = (randomVariableFactoryForNonDifferentiable
                                                          if randomVariableFactoryForNonDifferentiable is not None else RandomVariableCudaFactory())
        self.diracDeltaApproximationMethod = DiracDeltaApproximationMethod.valueOf(
            p.get("diracDeltaApproximationMethod", DiracDeltaApproximationMethod.DISCRETE_DELTA))
        self.diracDeltaApproximationWidthPerStdDev = float(p.get("diracDeltaApproximationWidthPerStdDev", p.get("barrierDiracWidth", 0.05)))
        self.diracDeltaApproximationDensityRegressionWidthPerStdDev = float(p.get("diracDeltaApproximationDensityRegressionWidthPerStdDev", 0.5))
        self._isGradientRetainsLeafNodesOnly = bool(p.get("isGradientRetainsLeafNodesOnly", True))

    def createRandomVariable(self, *args):
        """(value) | (time, value) | (time, values)."""
        return RandomVariableDifferentiableAAD(self.createRandomVariableNonDifferentiable(*args), self)

    def createRandomVariableNonDifferentiable(self, *args):
        if len(args) == 1:
            return self.randomVariableFactoryForNonDifferentiable.createRandomVariable(0.0, args[0])   # createRandomVariable(value) = (0.0, value)
        return self.randomVariableFactoryForNonDifferentiable.createRandomVariable(*args)

    def createRandomVariableArray(self, values):
        return [self.createRandomVariable(v) for v in values]

    # the two hooks the device Brownian motion uses on its factory (montecarlo.py): increments become independents without a copy
    @property
    def shard(self):
        return self.randomVariableFactoryForNonDifferentiable.shard

    def fromDevice(self, time, dv, n=None):
        return RandomVariableDifferentiableAAD(self.randomVariableFactoryForNonDifferentiable.fromDevice(time, dv, n), self)

    def createRandomVariableMatrix(self, values):
        return [self.createRandomVariableArray(row) for row in values]

    def getDiracDeltaApproximationMethod(self): return self.diracDeltaApproximationMethod
    def getDiracDeltaApproximationWidthPerStdDev(self): return self.diracDeltaApproximationWidthPerStdDev
    def getDiracDeltaApproximationDensityRegressionWidthPerStdDev(self): return self.diracDeltaApproximationDensityRegressionWidthPerStdDev
    def getBarrierDiracWidth(self): return self.diracDeltaApproximationWidthPerStdDev
    def isGradientRetainsLeafNodesOnly(self): return self._isGradientRetainsLeafNodesOnly

    def __repr__(self):
        return ("RandomVariableDifferentiableAADFactory [diracDeltaApproximationMethod=%s, diracDeltaApproximationWidthPerStdDev=%r, "
                "diracDeltaApproximationDensityRegressionWidthPerStdDev=%r, isGradientRetainsLeafNodesOnly=%r, inner=%r]" % (
                    self.diracDeltaApproximationMethod, self.diracDeltaApproximationWidthPerStdDev,
                    self.diracDeltaApproximationDensityRegressionWidthPerStdDev, self._isGradientRetainsLeafNodesOnly,
                    self.randomVariableFactoryForNonDifferentiable))
