"""TEST INFRASTRUCTURE ONLY: a host RandomVariable on numpy arrays (type priority 1), so that the book-keeping of autodiff.py (operator
tree, retention rules, backward sweep, type-priority delegation) can be exercised by ``-m "not gpu"`` without a device.  The product has
no CPU arithmetic; nothing under finmath-lib_b200/ imports this file."""
import numpy as np


def _quiet(f):
    with np.errstate(divide="ignore", invalid="ignore"):     # IEEE results like the JVM, no warnings
        return f()


def make(pkg):
    Scalar, RandomVariable = pkg.Scalar, pkg.RandomVariable

    def val(x):
        if isinstance(x, (int, float, np.floating, np.integer)):
            return float(x), float("-inf")
        if isinstance(x, NumpyRandomVariable):
            return x.v, x.time
        if x.isDeterministic():
            return x.doubleValue(), x.getFiltrationTime()
        return np.asarray(x.getRealizations()), x.getFiltrationTime()

    class NumpyRandomVariable(RandomVariable):
        def __init__(self, time, value):
            self.time = float(time)
            self.v = float(value) if np.ndim(value) == 0 else np.array(value, dtype=np.float64)

        # --- accessors
        def getTypePriority(self): return 1
        def getFiltrationTime(self): return self.time
        def isDeterministic(self): return np.ndim(self.v) == 0
        def size(self): return 1 if self.isDeterministic() else self.v.size
        def get(self, i): return self.v if self.isDeterministic() else float(self.v[i])
        def doubleValue(self): return float(self.v)
        def getRealizations(self): return np.array([self.v]) if self.isDeterministic() else self.v.copy()
        def cache(self): return self
        def equals(self, other): return bool(np.array_equal(np.broadcast_to(self.v, np.shape(val(other)[0]) or (1,)), np.atleast_1d(val(other)[0])))

        def getAverage(self, w=None):
            if w is None:
                return float(np.mean(self.v))
            return float(np.sum(self.v * val(w)[0] * np.ones(self.size())))

        def getVariance(self, w=None):
            if w is None:
                return float(np.mean(self.v * self.v) - np.mean(self.v) ** 2) if not self.isDeterministic() else 0.0
            m = self.getAverage(w)
            return float(np.sum((self.v - m) ** 2 * val(w)[0]))

        def getSampleVariance(self): return self.getVariance() * self.size() / (self.size() - 1)
        def getStandardDeviation(self, w=None): return float(np.sqrt(self.getVariance(w)))
        def getStandardError(self, w=None): return self.getStandardDeviation(w) / float(np.sqrt(self.size()))
        def getMin(self): return float(np.min(self.v))
        def getMax(self): return float(np.max(self.v))
        def average(self): return NumpyRandomVariable(float("-inf"), self.getAverage())

        # --- algebra
        def _higher(self, *xs):
            return any(not isinstance(x, (int, float, np.floating, np.integer)) and x.getTypePriority() > 1 for x in xs)

        def _new(self, v, *others):
            v = v() if callable(v) else v
            t = self.time
            for o in others:
                t = max(t, val(o)[1])
            return NumpyRandomVariable(t, v)

        def _u(self, f): return NumpyRandomVariable(self.time, f(self.v))
        def squared(self): return self._u(lambda x: x * x)
        def sqrt(self): return self._u(np.sqrt)
        def exp(self): return self._u(np.exp)
        def log(self): return self._u(np.log)
        def sin(self): return self._u(np.sin)
        def cos(self): return self._u(np.cos)
        def invert(self): return self._u(lambda x: 1.0 / x)
        def abs(self): return self._u(np.abs)
        def pow(self, e): return self._u(lambda x: np.power(x, e))
        def isNaN(self): return self._u(lambda x: np.where(np.isnan(x), 1.0, 0.0))

        def add(self, x): return x.add(self) if self._higher(x) else self._new(self.v + val(x)[0], x)
        def sub(self, x): return x.bus(self) if self._higher(x) else self._new(self.v - val(x)[0], x)
        def bus(self, x): return x.sub(self) if self._higher(x) else self._new(val(x)[0] - self.v, x)
        def mult(self, x): return x.mult(self) if self._higher(x) else self._new(self.v * val(x)[0], x)
        def div(self, x): return x.vid(self) if self._higher(x) else self._new(_quiet(lambda: self.v / val(x)[0]), x)
        def vid(self, x): return x.div(self) if self._higher(x) else self._new(_quiet(lambda: val(x)[0] / self.v), x)
        def cap(self, x): return x.cap(self) if self._higher(x) else self._new(np.minimum(self.v, val(x)[0]), x)
        def floor(self, x): return x.floor(self) if self._higher(x) else self._new(np.maximum(self.v, val(x)[0]), x)

        def accrue(self, rate, dt):
            return rate.mult(dt).add(1.0).mult(self) if self._higher(rate) else self._new(self.v * (1.0 + val(rate)[0] * dt), rate)

        def discount(self, rate, dt):
            return rate.mult(dt).add(1.0).invert().mult(self) if self._higher(rate) else self._new(self.v / (1.0 + val(rate)[0] * dt), rate)

        def choose(self, a, b):
            return self._new(np.where(self.v >= 0, val(a)[0], val(b)[0]), a, b)

        def addProduct(self, f1, f2):
            if self._higher(f1, f2):
                return f1.mult(f2).add(self)
            return self._new(self.v + val(f1)[0] * val(f2)[0], f1, f2)

        def addRatio(self, n, d):
            return n.div(d).add(self) if self._higher(n, d) else self._new(self.v + val(n)[0] / val(d)[0], n, d)

        def subRatio(self, n, d):
            return n.div(d).mult(-1).add(self) if self._higher(n, d) else self._new(self.v - val(n)[0] / val(d)[0], n, d)

        def __repr__(self):
            return "NumpyRandomVariable(%r, %r)" % (self.time, self.v)

    class NumpyRandomVariableFactory:
        shard = None

        def createRandomVariable(self, *args):
            if len(args) == 1:
                return NumpyRandomVariable(float("-inf"), args[0])
            return NumpyRandomVariable(*args)

    class NumpyBrownianMotion:
        """Stand-in for the device Brownian motion (generic Euler recipe only): numpy normals, same interface."""

        def __init__(self, timeDiscretization, numberOfFactors, numberOfPaths, seed, randomVariableFactory=None):
            self.timeDiscretization, self.numberOfFactors, self.numberOfPaths, self.seed = timeDiscretization, numberOfFactors, numberOfPaths, seed
            self.randomVariableFactory = randomVariableFactory if randomVariableFactory is not None else NumpyRandomVariableFactory()
            rng = np.random.default_rng(seed)
            T = timeDiscretization.getNumberOfTimeSteps()
            z = rng.standard_normal((T, numberOfFactors, numberOfPaths))
            self.increments = [[self.randomVariableFactory.createRandomVariable(timeDiscretization.getTime(t + 1),
                                                                                z[t, f] * np.sqrt(timeDiscretization.getTimeStep(t)))
                                for f in range(numberOfFactors)] for t in range(T)]

        def getCloneWithModifiedSeed(self, seed):
            return NumpyBrownianMotion(self.timeDiscretization, self.numberOfFactors, self.numberOfPaths, seed, self.randomVariableFactory)

        def getTimeDiscretization(self): return self.timeDiscretization
        def getNumberOfFactors(self): return self.numberOfFactors
        def getNumberOfPaths(self): return self.numberOfPaths
        def getBrownianIncrement(self, t, f): return self.increments[t][f]
        def getIncrement(self, t, f=None): return list(self.increments[t]) if f is None else self.increments[t][f]
        def getRandomVariableForConstant(self, v): return self.randomVariableFactory.createRandomVariable(v)

    return NumpyRandomVariable, NumpyRandomVariableFactory, NumpyBrownianMotion
