"""Monte-Carlo core on the device: time grid, Brownian driver, Euler scheme, regression estimator.

Mirrors (Java method names kept so that tests read like the reference's):
* ``TimeDiscretizationFromArray``   J/time/TimeDiscretizationFromArray.java:36-389 (host side; defines every dt)
* ``BrownianMotionCuda``            J/montecarlo/BrownianMotionFromMersenneRandomNumbers.java:41-258 (interfaces
                                    BrownianMotion.java:23-117, IndependentIncrements.java:24-112)
* ``EulerSchemeFromProcessModel``   J/montecarlo/process/EulerSchemeFromProcessModel.java:60-403
* ``MonteCarloConditionalExpectationRegression``  J/montecarlo/conditionalexpectation/MonteCarloConditionalExpectationRegression.java:33-180
"""
import ctypes as C
import math
import threading

import bisect

import numpy as np

from . import native as nv
from .sharding import LOCAL
from .stochastic import RandomVariableCuda, RandomVariableCudaFactory, Scalar, _unwrap

TIME_TICK_SIZE = 1.0 / (365.0 * 24.0)                       # TimeDiscretizationFromArray.java:39


class TimeDiscretizationFromArray:
    SHORT_PERIOD_AT_START, SHORT_PERIOD_AT_END = 0, 1       # ShortPeriodLocation :44-47

    def __init__(self, *args, tickSize=TIME_TICK_SIZE):
        self.tick = tickSize
        if len(args) == 3:                                  # (initial, numberOfTimeSteps, deltaT) :210-217
            initial, n, dt = args
            times = [initial + i * dt for i in range(int(n) + 1)]
        elif len(args) == 4:                                # (initial, last, deltaT, shortPeriodLocation) :227-249
            initial, last, dt, where = args
            count = int(math.ceil((last - initial) / dt)) + 1
            if where == self.SHORT_PERIOD_AT_END:
                times = [min(last, initial + i * dt) for i in range(count)]
            else:
                times = [max(initial, last - i * dt) for i in range(count)]
        else:
            times = list(args[0])
        rounded = sorted(set(self._round(t) for t in times))      # :57-64 round, distinct, sorted
        self.times = np.array(rounded, dtype=np.float64)
        self._list = rounded                                # the same values as Python floats (scalar lookups without numpy overhead)
        self._index = {t: i for i, t in enumerate(rounded)}
        self._asked = {}

    def _round(self, t):
        try:
            return float(round(t / self.tick)) * self.tick         # Math.rint (half-even) :387-389; Python's round() is half-even too
        except (OverflowError, ValueError):
            return float(np.rint(t / self.tick) * self.tick)

    def getNumberOfTimes(self):
        return self.times.size

    def getNumberOfTimeSteps(self):
        return self.times.size - 1

    def getTime(self, i):
        return self._list[i]

    def getTimeStep(self, i):
        return self._list[i + 1] - self._list[i]

    def getTimeIndex(self, time):                           # Arrays.binarySearch :272-274
        i = self._asked.get(time)                           # (products ask for the same few dates over and over)
        if i is None:
            key = self._round(time)
            i = self._index.get(key)
            if i is None:
                i = -(bisect.bisect_left(self._list, key) + 1)
            if len(self._asked) < 4096:
                self._asked[time] = i
        return i

    def getTimeIndexNearestLessOrEqual(self, time):
        i = self.getTimeIndex(time)
        return i if i >= 0 else -i - 2

    def getTimeIndexNearestGreaterOrEqual(self, time):     # :286-292 (binary search on the unrounded time, as in the reference)
        i = bisect.bisect_left(self._list, time)
        return i                                            # found: its index; not found: the insertion point (-index-1 of a negative result)

    def getAsDoubleArray(self):
        return self.times.copy()

    def getAsArrayList(self):                               # :301-307
        return list(self._list)

    def getFirstTime(self):                                 # TimeDiscretization.java:84-95
        return self._list[0]

    def getLastTime(self):
        return self._list[-1]

    def getTickSize(self):                                  # :345-347
        return self.tick

    def doubleStream(self):                                 # TimeDiscretization.java:118-120 (an iterator stands in for the DoubleStream)
        return iter(self._list)

    def __iter__(self):                                     # :350-352
        return iter(self._list)

    def getTimeShiftedTimeDiscretization(self, timeShift):  # :310-318: shifted times rounded with THIS tick size, new grid with the default one
        return TimeDiscretizationFromArray([self._round(t + timeShift) for t in self._list])

    def filter(self, timesToKeep):                          # :321-323
        return TimeDiscretizationFromArray([t for t in self._list if timesToKeep(t)], tickSize=self.tick)

    def union(self, that):                                  # :330-334: the finer tick size
        return TimeDiscretizationFromArray(self._list + list(that.getAsDoubleArray()), tickSize=min(self.tick, that.getTickSize()))

    def intersect(self, that):                              # :337-342: exact matches only, the coarser tick size
        other = set(float(t) for t in that.getAsDoubleArray())
        return TimeDiscretizationFromArray([t for t in self._list if t in other], tickSize=max(self.tick, that.getTickSize()))

    def __repr__(self):
        return "TimeDiscretizationFromArray [timeDiscretizationFromArray=%s, timeTickSize=%r]" % (self._list, self.tick)

    def __eq__(self, other):
        return isinstance(other, TimeDiscretizationFromArray) and np.array_equal(self.times, other.times) and self.tick == other.tick

    def __hash__(self):
        return hash((self.times.tobytes(), self.tick))


class BrownianMotionCuda:
    """BrownianMotionFromMersenneRandomNumbers on the device (counter-addressable MT19937 + AS241).

    State = (timeDiscretization, numberOfFactors, numberOfPaths, seed, factory) as in the reference (:45-51); increments are
    generated lazily on first use (:122-136), all T*F at once, into one device slab [t][f][path].  With a ShardContext of
    world > 1, ``numberOfPaths`` is the global count and this rank generates its contiguous block by jump-ahead.
    """

    def __init__(self, timeDiscretization, numberOfFactors, numberOfPaths, seed, randomVariableFactory=None, shard=None):
        self.timeDiscretization = timeDiscretization
        self.numberOfFactors = int(numberOfFactors)
        self.numberOfPaths = int(numberOfPaths)
        self.seed = int(np.int32(seed))
        self.shard = shard if shard is not None else (randomVariableFactory.shard if randomVariableFactory is not None else LOCAL)
        self.randomVariableFactory = randomVariableFactory if randomVariableFactory is not None else RandomVariableCudaFactory(self.shard)
        self._lock = threading.Lock()
        self._increments = None
        self._handles = None

    # ---- sharding --------------------------------------------------------------------------------------------
    def getPathRange(self):
        return self.shard.local_range(self.numberOfPaths)

    def getNumberOfLocalPaths(self):
        lo, hi = self.getPathRange()
        return hi - lo

    # ---- generation ------------------------------------------------------------------------------------------
    def _generate(self):
        T = self.timeDiscretization.getNumberOfTimeSteps()
        F = self.numberOfFactors
        lo, hi = self.getPathRange()
        sqrt_dt = np.array([math.sqrt(self.timeDiscretization.getTimeStep(t)) for t in range(T)], dtype=np.float64)   # :153-156
        out = np.zeros(T * F, dtype=np.uint64)
        nv.check(nv.load().fmb_bm_generate(self.seed, T, F, hi - lo, lo, nv.dptr(sqrt_dt), nv.hptr(out)))
        self._handles = out
        incs = []
        for t in range(T):
            time = self.timeDiscretization.getTime(t + 1)    # filtration time t_{i+1} :184-190
            incs.append([self.randomVariableFactory.fromDevice(time, nv.DeviceVector(out[t * F + f], hi - lo), self.numberOfPaths) for f in range(F)])
        self._increments = incs

    def _ensure(self):
        with self._lock:
            if self._increments is None:
                self._generate()

    def getBrownianIncrement(self, timeIndex, factor):
        self._ensure()
        return self._increments[timeIndex][factor]

    def getIncrement(self, timeIndex, factor=None):
        self._ensure()
        if factor is None:
            return list(self._increments[timeIndex])
        return self._increments[timeIndex][factor]

    def getIncrementHandles(self):
        self._ensure()
        return self._handles

    def getTimeDiscretization(self):
        return self.timeDiscretization

    def getNumberOfFactors(self):
        return self.numberOfFactors

    def getNumberOfPaths(self):
        return self.numberOfPaths

    def getSeed(self):
        return self.seed

    def getRandomVariableForConstant(self, value):
        return self.randomVariableFactory.createRandomVariable(value)

    def getCloneWithModifiedSeed(self, seed):
        return BrownianMotionCuda(self.timeDiscretization, self.numberOfFactors, self.numberOfPaths, seed, self.randomVariableFactory, self.shard)

    def getCloneWithModifiedTimeDiscretization(self, newTimeDiscretization):
        return BrownianMotionCuda(newTimeDiscretization, self.numberOfFactors, self.numberOfPaths, self.seed, self.randomVariableFactory, self.shard)

    def __eq__(self, other):                                # :227-251
        return (isinstance(other, BrownianMotionCuda) and self.timeDiscretization == other.timeDiscretization
                and self.numberOfFactors == other.numberOfFactors and self.numberOfPaths == other.numberOfPaths and self.seed == other.seed)

    def __hash__(self):
        return hash((self.timeDiscretization, self.numberOfFactors, self.numberOfPaths, self.seed))



class MersenneTwister:
    """J/randomnumbers/MersenneTwister.java:19-58 — the 1-D generator `new MersenneTwister(long seed)`.  On this backend it is a
    description of a stream (seed + position); the numbers themselves are produced on the device by jump-ahead."""

    def __init__(self, seed):
        self.seed = int(seed)
        self._position = 0                                   # uniforms handed out so far through nextDouble()

    def getDimension(self):
        return 1

    def nextDouble(self):                                    # sequential host access (tests, small uses): one device round trip per call
        u = np.empty(1)
        nv.check(nv.load().fmb_mt_uniforms(self.seed, self._position, 1, nv.dptr(u)))
        self._position += 1
        return float(u[0])

    def getNext(self):
        return [self.nextDouble()]


class RandomNumberGeneratorFrom1D:
    """J/randomnumbers/RandomNumberGeneratorFrom1D.java — `dimension` consecutive draws of a 1-D generator form one vector."""

    def __init__(self, randomNumberGenerator1D, dimension):
        self.generator, self.dimension = randomNumberGenerator1D, int(dimension)

    def getDimension(self):
        return self.dimension

    def getNext(self):
        return [self.generator.nextDouble() for _ in range(self.dimension)]


class _IncrementsFromUniforms:
    """Common part of the two classes below: uniforms U[t][f][path] in the reference's draw order (path -> time -> factor), generated on
    the device when the source is a MersenneTwister stream (bit-exact with the sequential generator at any path offset, so shards work),
    taken from the host generator path by path otherwise."""

    def _device_uniforms(self, seed):
        T, F = self.timeDiscretization.getNumberOfTimeSteps(), self.numberOfFactors
        lo, hi = self.shard.local_range(self.numberOfPaths)
        out = np.zeros(T * F, dtype=np.uint64)
        nv.check(nv.load().fmb_uniforms_generate(int(seed), T, F, hi - lo, lo, nv.hptr(out)))
        return [[nv.DeviceVector(out[t * F + f], hi - lo) for f in range(F)] for t in range(T)]

    def _host_uniforms(self, generator):
        T, F, P = self.timeDiscretization.getNumberOfTimeSteps(), self.numberOfFactors, self.numberOfPaths
        if generator.getDimension() < T * F:
            raise ValueError("The dimension of the random number generator is smaller than timeSteps * factors.")
        u = np.empty((P, T * F))
        for path in range(P):                                # the generator is a host object: sequential by contract (:160-161)
            u[path] = generator.getNext()[:T * F]
        lo, hi = self.shard.local_range(P)
        return [[nv.DeviceVector.upload(u[lo:hi, t * F + f]) for f in range(F)] for t in range(T)]

    def _wrap(self, timeIndex, dv):
        return self.randomVariableFactory.fromDevice(self.timeDiscretization.getTime(timeIndex + 1), dv, self.numberOfPaths)

    def getTimeDiscretization(self): return self.timeDiscretization
    def getNumberOfFactors(self): return self.numberOfFactors
    def getNumberOfPaths(self): return self.numberOfPaths
    def getRandomVariableForConstant(self, value): return self.randomVariableFactory.createRandomVariable(value)


class IndependentIncrementsFromICDF(_IncrementsFromUniforms):
    """J/montecarlo/IndependentIncrementsFromICDF.java:41-240 — Z_j(t_i) = ICDF_{i,j}(U_{i,j}), uniforms from MersenneTwister(seed) in
    the order path -> time -> factor (:173-206).  The uniforms are generated on the device.  An inverse distribution function is applied
    on the device when it accepts a RandomVariable (it is then written with RandomVariable operations, e.g.
    ``lambda u: u.mult(-1.0).add(1.0).log().mult(-1.0 / lam)``; NORMAL_ICDF is the built-in AS241 transform); a plain double -> double
    callable is applied on the host (download, map, upload), like `apply(DoubleUnaryOperator)` of the reference type would."""

    @staticmethod
    def NORMAL_ICDF(u):
        """NormalDistribution.inverseCumulativeDistribution as a device operation on a RandomVariableCuda."""
        return u._new(u.time, nv.unary(nv.U_ICDF_NORMAL, u.dv))

    def __init__(self, timeDiscretization, numberOfFactors, numberOfPaths, seed, inverseCumulativeDistributionFunctions, randomVariableFactory=None,
                 shard=None):
        self.timeDiscretization, self.numberOfFactors, self.numberOfPaths = timeDiscretization, int(numberOfFactors), int(numberOfPaths)
        self.seed = int(np.int32(seed))
        self.inverseCumulativeDistributionFunctions = inverseCumulativeDistributionFunctions
        self.shard = shard if shard is not None else (randomVariableFactory.shard if randomVariableFactory is not None else LOCAL)
        self.randomVariableFactory = randomVariableFactory if randomVariableFactory is not None else RandomVariableCudaFactory(self.shard)
        self._lock = threading.Lock()
        self._increments = None

    def _generate(self):
        uniforms = self._device_uniforms(self.seed)
        T, F = self.timeDiscretization.getNumberOfTimeSteps(), self.numberOfFactors
        incs = []
        for t in range(T):
            row = []
            for f in range(F):
                icdf = self.inverseCumulativeDistributionFunctions(t)(f)
                u = self._wrap(t, uniforms[t][f])
                try:
                    z = icdf(u)
                    if not isinstance(z, RandomVariableCuda):
                        raise TypeError
                    z = z if z.time == u.time else z._new(u.time, z.dv)
                except (TypeError, AttributeError):
                    values = u.getLocalRealizations()
                    z = self._wrap(t, nv.DeviceVector.upload(np.array([icdf(float(v)) for v in values], dtype=np.float64)))
                row.append(z)
            incs.append(row)
        self._increments = incs

    def getIncrement(self, timeIndex, factor):
        with self._lock:
            if self._increments is None:
                self._generate()
        return self._increments[timeIndex][factor]

    def getSeed(self): return self.seed

    def getCloneWithModifiedSeed(self, seed):
        return IndependentIncrementsFromICDF(self.timeDiscretization, self.numberOfFactors, self.numberOfPaths, seed, self.inverseCumulativeDistributionFunctions,
                                             self.randomVariableFactory, self.shard)

    def getCloneWithModifiedTimeDiscretization(self, newTimeDiscretization):
        return IndependentIncrementsFromICDF(newTimeDiscretization, self.numberOfFactors, self.numberOfPaths, self.seed, self.inverseCumulativeDistributionFunctions,
                                             self.randomVariableFactory, self.shard)


class BrownianMotionFromRandomNumberGenerator(_IncrementsFromUniforms):
    """J/montecarlo/BrownianMotionFromRandomNumberGenerator.java:41-220 — increments ICDF(u_{t*F+f}) * sqrt(dt_t) from the vectors
    `randomNumberGenerator.getNext()`, one per path (:160-170).  A MersenneTwister-based generator (RandomNumberGeneratorFrom1D over
    MersenneTwister) never leaves the device: the same jump-ahead stream as BrownianMotionFromMersenneRandomNumbers; any other generator
    object is asked path by path on the host and its numbers are uploaded; ICDF and scaling always run on the device."""

    def __init__(self, timeDiscretization, numberOfFactors, numberOfPaths, randomNumberGenerator, randomVariableFactory=None, shard=None):
        self.timeDiscretization, self.numberOfFactors, self.numberOfPaths = timeDiscretization, int(numberOfFactors), int(numberOfPaths)
        self.randomNumberGenerator = randomNumberGenerator
        self.shard = shard if shard is not None else (randomVariableFactory.shard if randomVariableFactory is not None else LOCAL)
        self.randomVariableFactory = randomVariableFactory if randomVariableFactory is not None else RandomVariableCudaFactory(self.shard)
        self._lock = threading.Lock()
        self._increments = None

    def _generate(self):
        g = self.randomNumberGenerator
        T, F = self.timeDiscretization.getNumberOfTimeSteps(), self.numberOfFactors
        if isinstance(g, RandomNumberGeneratorFrom1D) and isinstance(g.generator, MersenneTwister) and g.generator._position == 0 and g.getDimension() == T * F:
            uniforms = self._device_uniforms(g.generator.seed)
        else:
            uniforms = self._host_uniforms(g)
        incs = []
        for t in range(T):
            sqrtDeltaT = math.sqrt(self.timeDiscretization.getTimeStep(t))                  # :146-149
            incs.append([self._wrap(t, nv.unary(nv.U_MULT, nv.unary(nv.U_ICDF_NORMAL, uniforms[t][f]), sqrtDeltaT)) for f in range(F)])
        self._increments = incs

    def getBrownianIncrement(self, timeIndex, factor):
        with self._lock:
            if self._increments is None:
                self._generate()
        return self._increments[timeIndex][factor]

    def getIncrement(self, timeIndex, factor=None):
        if factor is None:
            return [self.getBrownianIncrement(timeIndex, f) for f in range(self.numberOfFactors)]
        return self.getBrownianIncrement(timeIndex, factor)

    def getCloneWithModifiedSeed(self, seed):
        raise NotImplementedError("a clone with a modified seed is not defined for a generic random number generator")    # :206-209 returns null

    def getCloneWithModifiedTimeDiscretization(self, newTimeDiscretization):
        return BrownianMotionFromRandomNumberGenerator(newTimeDiscretization, self.numberOfFactors, self.numberOfPaths, self.randomNumberGenerator,
                                                       self.randomVariableFactory, self.shard)


class _BrownianMotionDecorator:
    """Shared plumbing of the two wrappers below: they are not BrownianMotionCuda instances, so an Euler scheme on top of them
    runs the generic device loop (one kernel per RandomVariable operation) — the fused kernels need the raw increment slab."""

    def getIncrement(self, timeIndex, factor=None):
        if factor is None:
            return [self.getBrownianIncrement(timeIndex, f) for f in range(self.getNumberOfFactors())]
        return self.getBrownianIncrement(timeIndex, factor)

    def getTimeDiscretization(self): return self.brownianMotion.getTimeDiscretization()
    def getNumberOfPaths(self): return self.brownianMotion.getNumberOfPaths()
    def getRandomVariableForConstant(self, value): return self.brownianMotion.getRandomVariableForConstant(value)


class BrownianMotionView(_BrownianMotionDecorator):
    """J/montecarlo/BrownianMotionView.java:27-87 — a selection of factors of another Brownian motion (same device vectors)."""

    def __init__(self, brownianMotion, factors):
        self.brownianMotion, self.factors = brownianMotion, list(factors)

    def getBrownianIncrement(self, timeIndex, factor): return self.brownianMotion.getBrownianIncrement(timeIndex, self.factors[factor])
    def getNumberOfFactors(self): return len(self.factors)
    def getCloneWithModifiedSeed(self, seed): return BrownianMotionView(self.brownianMotion.getCloneWithModifiedSeed(seed), self.factors)

    def getCloneWithModifiedTimeDiscretization(self, td):
        return BrownianMotionView(self.brownianMotion.getCloneWithModifiedTimeDiscretization(td), self.factors)


class CorrelatedBrownianMotion(_BrownianMotionDecorator):
    """J/montecarlo/CorrelatedBrownianMotion.java:45-110 — dW_factor = sum_k factorLoadings[factor][k] * dU_k, accumulated with
    addProduct from a RandomVariableFromDoubleArray(0.0) in factor order, zero loadings skipped (:63-72)."""

    def __init__(self, uncollelatedFactors, factorLoadings):
        self.brownianMotion, self.factorLoadings = uncollelatedFactors, [list(map(float, row)) for row in factorLoadings]

    def getBrownianIncrement(self, timeIndex, factor):
        from .stochastic import RandomVariableFromDoubleArray
        increment = RandomVariableFromDoubleArray(0.0)
        for k, loading in enumerate(self.factorLoadings[factor]):
            if loading != 0:
                increment = increment.addProduct(self.brownianMotion.getBrownianIncrement(timeIndex, k), loading)
        return increment

    def getNumberOfFactors(self): return len(self.factorLoadings)
    def getCloneWithModifiedSeed(self, seed): return CorrelatedBrownianMotion(self.brownianMotion.getCloneWithModifiedSeed(seed), self.factorLoadings)

    def getCloneWithModifiedTimeDiscretization(self, td):
        return CorrelatedBrownianMotion(self.brownianMotion.getCloneWithModifiedTimeDiscretization(td), self.factorLoadings)

class Scheme:
    EULER, PREDICTOR_CORRECTOR, EULER_FUNCTIONAL, PREDICTOR_CORRECTOR_FUNCTIONAL = range(4)


class _LazyRow:
    __slots__ = ("store", "t")

    def __init__(self, store, t):
        self.store, self.t = store, t

    def __getitem__(self, c):
        return self.store.get(self.t, c)


class _LazyProcessValues:
    """discreteProcess[t][c] of a fused simulation: native handles wrapped into RandomVariableCuda objects on first use.
    A frozen component returns the very object of the earlier time index (aliasing of EulerSchemeFromProcessModel.java:285)."""

    def __init__(self, handles, T, N, P, bm, td, initialValues):
        self.handles, self.T, self.N, self.P = handles, T, N, P
        self.factory, self.nPaths, self.td, self.initialValues = bm.randomVariableFactory, bm.getNumberOfPaths(), td, initialValues
        self.cache = {}
        self.owned = set()                                   # indices whose handle reference has been handed to a DeviceVector

    def __getitem__(self, t):
        return _LazyRow(self, t)

    def get(self, t, c):
        i = t * self.N + c
        h = int(self.handles[i])
        if h == 0:
            key = ("det", c)
            if key not in self.cache:
                self.cache[key] = self.factory.createRandomVariable(self.initialValues[c])
            return self.cache[key]
        # first time index carrying this handle (later ones are aliases holding an extra native reference)
        t0 = t
        while t0 > 0 and int(self.handles[(t0 - 1) * self.N + c]) == h:
            t0 -= 1
        i0 = t0 * self.N + c
        rv = self.cache.get(i0)
        if rv is None:
            rv = self.factory.fromDevice(self.td.getTime(t0), nv.DeviceVector(h, self.P), self.nPaths)
            self.cache[i0] = rv
            self.owned.add(i0)
        return rv

    def __del__(self):
        try:
            hs = self.handles.copy()                         # one native reference per returned entry (aliases included) ...
            if self.owned:
                hs[list(self.owned)] = 0                     # ... except those handed to a DeviceVector, which releases its own
            nv.load().fmb_rv_free_many(nv.hptr(hs), hs.size)
        except Exception:
            pass


class EulerSchemeFromProcessModel:
    """Euler scheme on the device.

    If the model describes itself through ``getFusedSpecification`` (the four models of the path do), the whole time loop
    runs in ONE fused kernel (fmb_euler_*).  Any other ProcessModel is evolved by the generic loop below — the reference's
    own recipe (:202-318) on device RandomVariables, one kernel per operation — so arbitrary models keep working.
    """

    def __init__(self, model, stochasticDriver, scheme=None, forceGeneric=False):
        self.model = model
        self.stochasticDriver = stochasticDriver
        self.timeDiscretization = stochasticDriver.getTimeDiscretization()
        if scheme is None:                                   # :108-123
            scheme = Scheme.EULER_FUNCTIONAL
            try:
                model.applyStateSpaceTransformInverse(None, 0, 0, None)
            except NotImplementedError:
                scheme = Scheme.EULER
            except Exception:
                pass
        self.scheme = scheme
        self.forceGeneric = forceGeneric
        self._lock = threading.Lock()
        self._discreteProcess = None
        self._weights = None
        self.usedFusedKernel = None

    # ---- MonteCarloProcess / Process interface -------------------------------------------------------------------
    def getModel(self): return self.model
    def getStochasticDriver(self): return self.stochasticDriver
    def getBrownianMotion(self): return self.stochasticDriver
    def getScheme(self): return self.scheme
    def getTimeDiscretization(self): return self.timeDiscretization
    def getTime(self, timeIndex):
        if timeIndex < 0 or timeIndex >= self.timeDiscretization.getNumberOfTimes():
            raise IndexError("Index %d for process time discretization out of bounds." % timeIndex)
        return self.timeDiscretization.getTime(timeIndex)
    def getTimeIndex(self, time): return self.timeDiscretization.getTimeIndex(time)
    def getNumberOfPaths(self): return self.stochasticDriver.getNumberOfPaths()
    def getNumberOfFactors(self): return self.stochasticDriver.getNumberOfFactors()
    def getNumberOfComponents(self): return self.model.getNumberOfComponents()

    def getProcessValue(self, timeIndex, componentIndex=None):
        if timeIndex < 0 or timeIndex >= self.timeDiscretization.getNumberOfTimes():
            raise IndexError("time index %d out of bounds" % timeIndex)      # ArrayIndexOutOfBoundsException in the reference (no wrap-around)
        with self._lock:
            if self._discreteProcess is None:
                self._precalculate()
        if componentIndex is None:
            return [self._discreteProcess[timeIndex][c] for c in range(self.getNumberOfComponents())]
        return self._discreteProcess[timeIndex][componentIndex]

    def getMonteCarloWeights(self, timeIndex):
        with self._lock:
            if self._discreteProcess is None:
                self._precalculate()
        return self._weights

    def clone(self):
        return EulerSchemeFromProcessModel(self.model, self.stochasticDriver, self.scheme, self.forceGeneric)

    def getCloneWithModifiedModel(self, model):
        return EulerSchemeFromProcessModel(model, self.stochasticDriver, self.scheme, self.forceGeneric)

    def getCloneWithModifiedSeed(self, seed):
        return EulerSchemeFromProcessModel(self.model, self.stochasticDriver.getCloneWithModifiedSeed(seed), self.scheme, self.forceGeneric)

    def getCloneWithModifiedData(self, dataModified):         # EulerSchemeFromProcessModel.java:370-391: model, seed | stochasticDriver, scheme
        d = dict(dataModified or {})
        if "seed" in d and "stochasticDriver" in d:
            raise ValueError("Simultaneous specification of stochasticDriver and seed.")
        driver = self.stochasticDriver
        if "seed" in d:
            driver = driver.getCloneWithModifiedSeed(int(d["seed"]))
        elif "stochasticDriver" in d:
            driver = d["stochasticDriver"]
        return EulerSchemeFromProcessModel(d.get("model", self.model), driver, d.get("scheme", self.scheme), self.forceGeneric)

    # ---- evolution -------------------------------------------------------------------------------------------------
    def _precalculate(self):
        self._weights = self.stochasticDriver.getRandomVariableForConstant(1.0 / self.getNumberOfPaths())    # :184
        differentiable = any(hasattr(getattr(o, "randomVariableFactory", None), "createRandomVariableNonDifferentiable")
                             for o in (self.model, self.stochasticDriver))
        # parameters or increments from a differentiable factory: every operation has to be recorded, so the reference's generic
        # recipe runs on (wrapped) device RandomVariables instead of the fused kernel (autodiff.py)
        spec = None if (self.forceGeneric or differentiable) else getattr(self.model, "getFusedSpecification", lambda p: None)(self)
        if spec is not None and spec["kernel"] in ("heston", "hull_white") and self.stochasticDriver.getNumberOfFactors() != 2:
            # the fused kernels of these two models read exactly two increments per step; any other driver runs the reference's
            # generic recipe (addSumProduct over whatever factor loadings the model returns)
            spec = None
        if spec is not None and isinstance(self.stochasticDriver, BrownianMotionCuda):
            self._precalculate_fused(spec)
            self.usedFusedKernel = spec["kernel"]
        else:
            self._precalculate_generic()
            self.usedFusedKernel = None

    def _precalculate_fused(self, spec):
        bm = self.stochasticDriver
        td = self.timeDiscretization
        T, N, F = td.getNumberOfTimeSteps(), self.getNumberOfComponents(), bm.getNumberOfFactors()
        P = bm.getNumberOfLocalPaths()
        dt = np.array([td.getTime(t + 1) - td.getTime(t) for t in range(T)], dtype=np.float64)        # :206
        dW = nv.handles(bm.getIncrementHandles())
        out = np.zeros((T + 1) * N, dtype=np.uint64)
        lib = nv.load()
        k = spec["kernel"]
        if k == "black_scholes":
            nv.check(lib.fmb_euler_black_scholes(self.scheme, T, F, P, nv.dptr(dt), nv.hptr(dW), spec["initialValue"], spec["riskFreeRate"],
                                                 spec["volatility"], nv.hptr(out)))
        elif k == "heston":
            rates = nv.as_f64(spec["riskFreeRates"])
            nv.check(lib.fmb_euler_heston(self.scheme, spec["hestonScheme"], T, P, nv.dptr(dt), nv.hptr(dW), spec["initialValue"], nv.dptr(rates),
                                          spec["volatility"], spec["theta"], spec["kappa"], spec["xi"], spec["rho"], nv.hptr(out)))
        elif k == "hull_white":
            d0, d1, fl = nv.as_f64(spec["drift0"]), nv.as_f64(spec["drift1"]), nv.as_f64(spec["factorLoadings"])
            nv.check(lib.fmb_euler_hull_white(T, P, nv.dptr(dt), nv.hptr(dW), nv.dptr(d0), nv.dptr(d1), nv.dptr(fl), nv.hptr(out)))
        elif k == "lmm":
            y0, pl = nv.as_f64(spec["initialState"]), nv.as_f64(spec["periodLength"])
            fl, var = nv.as_f64(spec["factorLoading"]), nv.as_f64(spec["variance"])
            first = np.ascontiguousarray(spec["firstLive"], dtype=np.int32)
            nv.check(lib.fmb_euler_lmm(self.scheme, spec["measure"], spec["stateSpace"], spec["liborCap"], T, N, F, P, nv.dptr(dt), nv.hptr(dW),
                                       nv.dptr(y0), nv.dptr(pl), nv.dptr(fl), nv.dptr(var), first.ctypes.data_as(nv.c_ip), nv.hptr(out)))
        else:
            raise ValueError("unknown fused kernel " + str(k))
        # RandomVariable wrappers are created on demand (a simulation has (T+1)*N = 1640 process values for C4, a product touches a
        # few dozen); handles never wrapped are released when the process is collected.
        self._discreteProcess = _LazyProcessValues(out, T, N, P, bm, td, spec["initialValues"])

    def _precalculate_generic(self):
        model, driver, td = self.model, self.stochasticDriver, self.timeDiscretization
        T, N = td.getNumberOfTimeSteps(), self.getNumberOfComponents()
        functional = self.scheme in (Scheme.EULER_FUNCTIONAL, Scheme.PREDICTOR_CORRECTOR_FUNCTIONAL)
        pc = self.scheme in (Scheme.PREDICTOR_CORRECTOR, Scheme.PREDICTOR_CORRECTOR_FUNCTIONAL)
        current = list(model.getInitialState(self))
        proc = [[model.applyStateSpaceTransform(self, 0, c, current[c]) for c in range(N)]]
        for ti in range(1, T + 1):
            deltaT = td.getTime(ti) - td.getTime(ti - 1)
            drift = model.getDrift(self, ti - 1, proc[ti - 1], None)
            dW = driver.getIncrement(ti - 1)
            row = []
            for c in range(N):
                if drift[c] is None:
                    row.append(proc[ti - 1][c])
                    continue
                if functional:
                    current[c] = model.applyStateSpaceTransformInverse(self, ti - 1, c, proc[ti - 1][c])
                fl = model.getFactorLoading(self, ti - 1, c, proc[ti - 1])
                if fl is None:
                    row.append(proc[ti - 1][c])
                    continue
                current[c] = current[c].addProduct(drift[c], deltaT)
                current[c] = current[c].addSumProduct(fl, dW)
                row.append(model.applyStateSpaceTransform(self, ti, c, current[c]))
            proc.append(row)
            if pc:
                driftP = model.getDrift(self, ti - 1, proc[ti], None)
                for c in range(N):
                    if driftP[c] is None or drift[c] is None:
                        continue
                    adj = driftP[c].sub(drift[c]).div(2.0).mult(deltaT)
                    current[c] = current[c].add(adj)
                    proc[ti][c] = model.applyStateSpaceTransform(self, ti, c, current[c])
        self._discreteProcess = proc


class RegressionBasisFunctionsGiven:
    """MonteCarloConditionalExpectationRegression.java:45-60 (RegressionBasisFunctions: a supplier of basis functions)."""

    def __init__(self, basisFunctions):
        self.basisFunctions = list(basisFunctions)

    def getBasisFunctions(self):
        return self.basisFunctions


class MonteCarloConditionalExpectationRegression:
    """Least-squares conditional expectation.  XtX and Xty are accumulated in ONE fused pass instead of K(K+1)/2 + K multiply-and-reduce
    passes.  Default path (single GPU, or shards with the library's own communicator): everything stays on the device —
    fmb_regression_conditional_expectation queues moments -> (all-gather of the shards' moments) -> K x K solve -> prediction on the
    compute stream and returns at once, so a backward induction never waits for an exercise date.  The coefficients are downloaded
    only when somebody asks for them (getLinearRegressionParameters / lastParameters).  With host-side exchange of the partials
    (FMB_TINY_COLLECTIVES=shm|torch) the moments come back to the host, are merged across shards and solved there."""

    def __init__(self, basisFunctionsEstimator, basisFunctionsPredictor=None):
        est = [b for b in basisFunctionsEstimator if b is not None]            # :79-95 drops nulls
        pre = est if basisFunctionsPredictor is None else [b for b in basisFunctionsPredictor if b is not None]
        self.basisFunctionsEstimator, self.basisFunctionsPredictor = est, pre
        self._XTX = None
        self._cachedFit = None                               # device-resident first fit: its XtX is reused (the solver is cached per instance, :125-138)
        self._lastFit = None
        self._lastK = 0
        self._lastParameters = None
        self._lastConditionNumber = None

    @staticmethod
    def _as_cuda(rv, shard):
        rv = _unwrap(rv)                                      # basis functions / dependents may be AAD wrappers: regress on their device values
        if isinstance(rv, RandomVariableCuda):
            return rv
        if rv.isDeterministic():
            return RandomVariableCuda(rv.getFiltrationTime(), rv.doubleValue(), shard)
        return RandomVariableCuda(rv.getFiltrationTime(), rv.getRealizations(), shard)

    # ---- results of the last fit (downloaded on demand) ------------------------------------------------------------
    def _fetch_fit(self):
        if self._lastFit is not None and self._lastParameters is None:
            K = self._lastK
            x, cond = np.zeros(K), C.c_double()
            nv.check(nv.load().fmb_regression_fit_get(self._lastFit.h, K, None, None, nv.dptr(x), C.byref(cond)))
            self._lastParameters, self._lastConditionNumber = x, cond.value

    @property
    def lastParameters(self):
        self._fetch_fit()
        return self._lastParameters

    @property
    def lastConditionNumber(self):
        self._fetch_fit()
        return self._lastConditionNumber

    @staticmethod
    def _basis_args(basis):
        hs = np.array([b.dv.h if b.dv is not None else 0 for b in basis], dtype=np.uint64)
        sc = np.array([b.valueIfNonStochastic if b.dv is None else 0.0 for b in basis], dtype=np.float64)
        return hs, sc

    def _device_resident(self, y, basis):
        shard = y.shard
        return (y.dv is not None and len(basis) <= 8 and (shard.world == 1 or shard.native_comm)
                and all(b.dv is None or b.dv.n == y.dv.n for b in basis))

    def _fit_on_device(self, basis, y):
        K = len(basis)
        hs, sc = self._basis_args(basis)
        fit = C.c_uint64()
        cached = self._cachedFit.h if self._cachedFit is not None else 0
        nv.check(nv.load().fmb_regression_fit(K, nv.hptr(hs), nv.dptr(sc), y.dv.h, y.size(), cached, C.byref(fit)))
        self._set_fit(nv.DeviceVector(fit.value, 96), K)

    def _set_fit(self, fit, K):
        self._lastFit, self._lastK = fit, K
        self._lastParameters = self._lastConditionNumber = None
        if self._cachedFit is None:
            self._cachedFit = fit

    def _moments(self, basis, y):
        K = len(basis)
        shard = y.shard
        hs, sc = self._basis_args(basis)
        xh, xl = np.zeros(K * K), np.zeros(K * K)
        yh, yl = np.zeros(K), np.zeros(K)
        nv.check(nv.load().fmb_regression_moments(K, nv.hptr(hs), nv.dptr(sc), y.dv.h, nv.dptr(xh), nv.dptr(xl), nv.dptr(yh), nv.dptr(yl)))
        H, L = shard.sum_dd_many(np.concatenate([xh, yh]), np.concatenate([xl, yl]), force=True)
        n = y.size()
        tot = (H + L) / n
        XTX, XTy = tot[:K * K].reshape(K, K).copy(), tot[K * K:].copy()
        for i in range(K):                                    # deterministic x deterministic: mult() stays a scalar, average = the product
            for j in range(K):
                if basis[i].dv is None and basis[j].dv is None:
                    XTX[i, j] = basis[i].valueIfNonStochastic * basis[j].valueIfNonStochastic
        return XTX, XTy

    def getLinearRegressionParameters(self, dependents):      # :118-150
        shard = dependents.shard if isinstance(dependents, RandomVariableCuda) else LOCAL
        y = self._as_cuda(dependents, shard)
        basis = [self._as_cuda(b, shard) for b in self.basisFunctionsEstimator]
        K = len(basis)
        if self._device_resident(y, basis):
            self._fit_on_device(basis, y)
            return self.lastParameters
        if y.dv is None or K > 8:
            # deterministic dependents / more than 8 basis functions: the generic op path (one kernel per product + reduction)
            if self._XTX is None:
                self._XTX = np.array([[basis[i].mult(basis[j]).getAverage() for j in range(K)] for i in range(K)])
            XTX = self._XTX
            XTy = np.array([y.mult(b).getAverage() for b in basis])
        else:
            XTX, XTy = self._moments(basis, y)
            if self._XTX is None:
                self._XTX = XTX                               # the solver is cached per instance (:125-138)
            XTX = self._XTX
        x = np.zeros(K)
        cond = C.c_double()
        A = nv.as_f64(XTX)
        b = nv.as_f64(XTy)
        nv.check(nv.load().fmb_regression_solve_svd(K, nv.dptr(A), nv.dptr(b), nv.dptr(x), C.byref(cond)))
        self._lastFit = None
        self._lastConditionNumber = cond.value
        self._lastParameters = x
        return x

    def getBasisFunctionsEstimator(self):                     # MonteCarloConditionalExpectationRegression.java:152-158
        return RegressionBasisFunctionsGiven(self.basisFunctionsEstimator)

    def getBasisFunctionsPredictor(self):
        return RegressionBasisFunctionsGiven(self.basisFunctionsPredictor)

    def getConditionalExpectation(self, randomVariable):      # :97-110
        shard = randomVariable.shard if isinstance(randomVariable, RandomVariableCuda) else LOCAL
        y = self._as_cuda(randomVariable, shard)
        est = [self._as_cuda(b, shard) for b in self.basisFunctionsEstimator]
        same = self.basisFunctionsPredictor is self.basisFunctionsEstimator
        basis = est if same else [self._as_cuda(b, shard) for b in self.basisFunctionsPredictor]
        K = len(basis)
        stochastic = next((b for b in basis if b.dv is not None), None)
        plain = type(self).getLinearRegressionParameters is MonteCarloConditionalExpectationRegression.getLinearRegressionParameters
        if (plain and stochastic is not None and K == len(est) and self._device_resident(y, est)
                and all(b.dv is None or b.dv.n == y.dv.n for b in basis)):
            # one call, nothing comes back: moments -> (exchange) -> solve -> prediction are queued on the compute stream
            hs, sc = self._basis_args(est)
            fit, out = C.c_uint64(), C.c_uint64()
            cached = self._cachedFit.h if self._cachedFit is not None else 0
            if same:
                rc = nv.load().fmb_regression_conditional_expectation(K, nv.hptr(hs), nv.dptr(sc), y.dv.h, y.size(), cached, K, None, None, C.byref(fit), C.byref(out))
            else:
                hp, sp = self._basis_args(basis)
                rc = nv.load().fmb_regression_conditional_expectation(K, nv.hptr(hs), nv.dptr(sc), y.dv.h, y.size(), cached, K, nv.hptr(hp), nv.dptr(sp),
                                                                      C.byref(fit), C.byref(out))
            nv.check(rc)
            self._set_fit(nv.DeviceVector(fit.value, 96), K)
            time = max(b.getFiltrationTime() for b in basis)
            return RandomVariableCuda(time, None, shard, _dv=nv.DeviceVector(out.value, stochastic.dv.n), _n=stochastic.nGlobal)
        x = self.getLinearRegressionParameters(randomVariable)
        if K <= 8 and stochastic is not None:
            hs, sc = self._basis_args(basis)
            out = C.c_uint64()
            xs = nv.as_f64(x)
            nv.check(nv.load().fmb_regression_predict(K, nv.hptr(hs), nv.dptr(sc), nv.dptr(xs), C.byref(out)))
            time = max(b.getFiltrationTime() for b in basis)
            return RandomVariableCuda(time, None, shard, _dv=nv.DeviceVector(out.value, stochastic.dv.n), _n=stochastic.nGlobal)
        ce = basis[0].mult(float(x[0]))
        for i in range(1, K):
            ce = ce.addProduct(basis[i], float(x[i]))
        return ce


class MonteCarloConditionalExpectationRegressionLocalizedOnDependents(MonteCarloConditionalExpectationRegression):
    """J/montecarlo/conditionalexpectation/MonteCarloConditionalExpectationRegressionLocalizedOnDependents.java:85-128: the regression
    is restricted to the paths where |dependents| < standardDeviations * stddev(dependents) (a 0/1 weight multiplies basis functions
    and dependents), no caching of XtX.  The moment accumulation is the same fused kernel."""

    def __init__(self, basisFunctionsEstimator, basisFunctionsPredictor=None, standardDeviations=4.0):
        super().__init__(basisFunctionsEstimator, basisFunctionsPredictor)
        self.standardDeviations = standardDeviations

    def getLinearRegressionParameters(self, dependents):
        localizerWeights = dependents.squared().sub(math.pow(dependents.getStandardDeviation() * self.standardDeviations, 2.0)).choose(Scalar(0.0), Scalar(1.0))
        saved = self.basisFunctionsEstimator
        try:
            self.basisFunctionsEstimator = [b.mult(localizerWeights) for b in saved]
            self._XTX = self._cachedFit = None
            return super().getLinearRegressionParameters(dependents.mult(localizerWeights))
        finally:
            self.basisFunctionsEstimator = saved
            self._XTX = self._cachedFit = None


class LinearRegression:
    """J/montecarlo/conditionalexpectation/LinearRegression.java:30-88 (closed forms for one and two basis functions, least squares else)."""

    def __init__(self, basisFunctions):
        self.basisFunctions = list(basisFunctions)

    def getRegressionCoefficients(self, value):
        b = self.basisFunctions
        if len(b) == 0:
            return np.array([])
        if len(b) == 1:
            return np.array([value.mult(b[0]).getAverage() / b[0].squared().getAverage()])
        if len(b) == 2:
            a = b[0].squared().getAverage()
            bb = b[0].mult(b[1]).average().squared().doubleValue()           # :45 (sic: the squared average)
            c, d = bb, b[1].squared().getAverage()
            determinant = a * d - bb * c
            if determinant != 0:
                x, y = value.mult(b[0]).getAverage(), value.mult(b[1]).getAverage()
                return np.array([(d * x - bb * y) / determinant, (a * y - c * x) / determinant])
        K = len(b)
        BTB = np.zeros((K, K))
        for i in range(K):
            for j in range(i + 1):
                BTB[i, j] = BTB[j, i] = b[i].mult(b[j]).getAverage()
        BTX = np.array([b[i].mult(value).getAverage() for i in range(K)])
        x = np.zeros(K)
        A, rhs = nv.as_f64(BTB), nv.as_f64(BTX)
        nv.check(nv.load().fmb_regression_solve_svd(K, nv.dptr(A), nv.dptr(rhs), nv.dptr(x), None))
        return x
