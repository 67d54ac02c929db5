"""autodiff.py (AAD on top of the GPU type), every case twice:

* ``-m "not gpu"``: the book-keeping (operator tree, retention, backward sweep, type priority, the generic Euler recipe recording a
  model) on a numpy test type (tests/numpy_rv.py — test infrastructure, the product has no CPU arithmetic);
* ``-m gpu``: the same cases with ``RandomVariableDifferentiableAADFactory(RandomVariableCudaFactory())`` — values, partial derivatives and
  adjoints are device vectors — plus the cases that need the device (finite differences through the FUSED kernels, regressions).

The cases follow the reference's own tests for the class: T/montecarlo/automaticdifferentiation/RandomVariableDifferentiableTest.java
(:140-178 simple gradient, :219-255 big sum, :350-377 expectation, :425-520 interface vs finite differences),
…/backward/RandomVariableDifferentiableAADTest.java (:22-62 second order, :65-262 operators, :265-318 expectation / variance) and
…/RandomVariableDifferentiableTypePriorityTest.java (result type upon commutation)."""
from collections import namedtuple

import numpy as np
import pytest

from numpy_rv import make

Env = namedtuple("Env", "pkg RV Factory f BM priority device")


@pytest.fixture(scope="module", params=["numpy", pytest.param("device", marks=pytest.mark.gpu)])
def env(request, pkg):
    if request.param == "numpy":
        RV, Factory, BM = make(pkg)
        return Env(pkg, RV, Factory, pkg.RandomVariableDifferentiableAADFactory(Factory()), BM, 1, False)
    pkg.native.init(0)                                       # fails loudly without a GPU (no CPU fallback)
    return Env(pkg, pkg.RandomVariableCuda, pkg.RandomVariableCudaFactory, pkg.RandomVariableDifferentiableAADFactory(pkg.RandomVariableCudaFactory()),
               pkg.BrownianMotionCuda, 2, True)


def values(rv, n=5):
    """Realizations, a deterministic result broadcast (Scalar.getRealizations() is null in the reference, Scalar.java:78-80)."""
    return np.full(n, rv.doubleValue()) if rv.isDeterministic() else np.asarray(rv.getRealizations())


X1 = np.array([3.0, 1.0, 0.0, 2.0, 4.0])
X2 = np.array([-4.0, -2.0, 0.0, 2.0, 4.0])


def test_simple_gradient(env):
    pkg, RV, _, f = env[:4]
    x1, x2 = f.createRandomVariable(0.0, X1), f.createRandomVariable(0.0, X2)
    y = x1.add(x2).mult(x1).add(x1)                          # x1^2 + x1 x2 + x1
    g = y.getGradient()
    assert sorted(g) == [x1.getID(), x2.getID()]            # leaves only by default
    assert np.array_equal(g[x1.getID()].getRealizations(), 2.0 * X1 + X2 + 1.0)
    assert np.array_equal(g[x2.getID()].getRealizations(), X1)
    assert np.array_equal(y.getRealizations(), X1 * X1 + X1 * X2 + X1)
    # a subset of independents
    assert list(y.getGradient({x2.getID()})) == [x2.getID()]
    with pytest.raises(NotImplementedError):
        y.getTangents()
    with pytest.raises(NotImplementedError):
        y.apply(lambda v: v)


def test_big_sum_and_constants(env):
    pkg, RV, _, f = env[:4]
    x = f.createRandomVariable(0.0, X1)
    s = f.createRandomVariable(0.0)
    for i in range(1000):
        s = s.add(x)                                         # :219-255: the derivative of the sum of n copies is n
    assert np.array_equal(values(s.getGradient()[x.getID()]), np.full(5, 1000.0))
    s = x
    for i in range(50):
        s = s.add(x.mult(2.0)).sub(1.5).add(RV(0.0, X2)).add(pkg.Scalar(4.0))
    g = s.getGradient()
    assert sorted(g) == [x.getID()] and np.array_equal(values(g[x.getID()]), np.full(5, 101.0))


def everything(v, RV, pkg):
    """One expression through every differentiable operation (arguments: 4 random variables, positive where needed)."""
    a, b, c, d = v
    k = RV(0.0, np.linspace(0.5, 1.5, a.size()))             # a non-differentiable stochastic constant
    y = a.squared().add(b.sqrt()).sub(c.exp().mult(0.1)).add(d.log())
    y = y.add(a.sin()).add(b.cos()).add(c.invert()).add(d.sub(0.9).abs())
    y = y.add(a.mult(b)).add(a.div(c)).add(a.mult(k)).add(k.div(b)).add(d.vid(2.0)).add(d.bus(3.0)).add(k.sub(c)).add(c.bus(k)).add(b.vid(k))
    y = y.add(a.cap(b)).add(c.floor(d)).add(a.cap(1.0)).add(b.floor(1.0)).add(k.cap(a)).add(k.floor(c))
    y = y.add(a.pow(2.5)).add(a.addProduct(b, c)).add(a.addProduct(b, 0.7)).add(k.addProduct(c, d))
    y = y.add(a.addRatio(b, c)).add(b.subRatio(c, d)).add(k.addRatio(a, k)).add(k.subRatio(k, b))
    y = y.add(a.accrue(b, 0.5)).add(c.discount(d, 0.25)).add(k.accrue(a, 0.5)).add(k.discount(b, 0.5)).add(pkg.Scalar(2.0).discount(c, 0.5))
    y = y.add(a.sub(1.0).choose(b, c).mult(d))                # (the derivative w.r.t. the trigger is a separate test)
    y = y.add(b.average().mult(c)).add(a.mult(d).expectation())
    return y


def test_every_operator_against_finite_differences(env):
    pkg, RV, _, _ = env[:4]
    f = pkg.RandomVariableDifferentiableAADFactory(env.Factory(), {"diracDeltaApproximationMethod": "ZERO"})
    rng = np.random.default_rng(7)
    n = 257
    base = [rng.uniform(0.6, 1.4, n) for _ in range(4)]
    base[0][::3] += 0.45                                      # both sides of the trigger a - 1
    leaves = [f.createRandomVariable(0.0, v) for v in base]
    y = everything(leaves, RV, pkg)
    want = everything([RV(0.0, v) for v in base], RV, pkg)
    # wrapping does not change a value (up to the re-ordered arithmetic of the delegations, e.g. k.discount(b) = (1 + b dt)^-1 · k)
    assert np.allclose(y.getRealizations(), want.getRealizations(), rtol=1e-14, atol=0)
    g = y.getGradient()
    eps = 1e-6
    for j in range(4):
        for averaged in (False, True):
            # pathwise bump (all paths at once); terms under an average respond to the bump of every path: compare E[gradient] there
            up, dn = [RV(0.0, v) for v in base], [RV(0.0, v) for v in base]
            up[j], dn[j] = up[j].add(eps), dn[j].sub(eps)
            fd = everything(up, RV, pkg).sub(everything(dn, RV, pkg)).div(2 * eps).getRealizations()
            ad = g[leaves[j].getID()].getRealizations()
            if averaged:
                assert abs(np.mean(fd) - np.mean(ad)) < 2e-7 * max(1.0, abs(np.mean(fd)))
        # pathwise: remove the two averaged terms from both sides
        def pathwise(v):
            return everything(v, RV, pkg).sub(v[1].average().mult(v[2])).sub(v[0].mult(v[3]).expectation())
        up, dn = [RV(0.0, v) for v in base], [RV(0.0, v) for v in base]
        up[j], dn[j] = up[j].add(eps), dn[j].sub(eps)
        fd = pathwise(up).sub(pathwise(dn)).div(2 * eps).getRealizations()
        ad = pathwise(leaves).getGradient()[leaves[j].getID()].getRealizations()
        assert np.max(np.abs(fd - ad) / np.maximum(1.0, np.abs(fd))) < 5e-8, j


def test_expectation_operator_and_variance(env):
    pkg, RV, _, _ = env[:4]
    f = pkg.RandomVariableDifferentiableAADFactory(env.Factory(), {"isGradientRetainsLeafNodesOnly": False})
    x = f.createRandomVariable(0.0, [0.0, 2.0, 1.0, -2.0, -1.0])
    a = f.createRandomVariable(0.0, [3.0] * 5)
    eps = 1e-8
    for fn in (lambda v: v.expectation(), lambda v: v.variance(), lambda v: v.getVarianceAsRandomVariableAAD() if hasattr(v, "getID") else
               RV(0.0, v.getVariance()), lambda v: v.squared().average().sqrt()):
        y = a.mult(fn(x))
        dydx = y.getGradient()[x.getID()]
        xv, av = x.getValues(), a.getValues()
        fd = av.mult(fn(xv.add(eps))).sub(av.mult(fn(xv.add(-eps)))).div(2 * eps)
        assert abs(dydx.expectation().doubleValue() - fd.expectation().doubleValue()) < 1e-7
    # E[ x (W - E W) ] W has derivative with vanishing average (:350-377)
    w = RV(1.0, np.random.default_rng(3141).standard_normal(100000))
    one = f.createRandomVariable(1.0)
    y = one.mult(w.sub(w.average())).average().mult(w)
    assert abs(y.getAverage()) < 1e-8 and abs(y.getGradient()[one.getID()].getAverage()) < 1e-8


def test_statistics_vertices(env):
    pkg, RV, _, f = env[:4]
    v = np.array([0.3, -1.2, 2.5, 0.7, 1.1, -0.4])
    eps = 1e-7
    for name, plain in (("getVarianceAsRandomVariableAAD", "getVariance"), ("getSampleVarianceAsRandomVariableAAD", "getSampleVariance"),
                        ("getStandardDeviationAsRandomVariableAAD", "getStandardDeviation"), ("getStandardErrorAsRandomVariableAAD", "getStandardError"),
                        ("getMinAsRandomVariableAAD", "getMin"), ("getMaxAsRandomVariableAAD", "getMax")):
        x = f.createRandomVariable(0.0, v)
        s = getattr(x, name)()
        assert s.isDeterministic() and s.doubleValue() == getattr(x, plain)()
        g = s.getGradient()[x.getID()].getRealizations()
        for i in range(v.size):                              # d statistic / d x_i by bumping one path (the reference's formulas, :262-300, are
            up, dn = v.copy(), v.copy()                      # written per path up to its own (2n-1)/n convention: check the exact ones only)
            up[i] += eps
            dn[i] -= eps
            fd = (getattr(RV(0.0, up), plain)() - getattr(RV(0.0, dn), plain)()) / (2 * eps)
            if plain in ("getMin", "getMax"):
                assert abs(g[i] - fd) < 1e-6, (name, i)
        if plain == "getVariance":                           # the reference's expression, evaluated by hand
            n = v.size
            assert np.allclose(g, (v - np.mean(v) * (2.0 * n - 1.0) / n) * 2.0 / n, rtol=0, atol=1e-15)


def test_second_order_by_nesting(env):
    pkg, RV, Factory, _ = env[:4]
    props = {"isGradientRetainsLeafNodesOnly": False}
    inner = pkg.RandomVariableDifferentiableAADFactory(Factory(), props)
    outer = pkg.RandomVariableDifferentiableAADFactory(inner, props)
    a, b = RV(0.0, 5.0), RV(0.0, 1.0)
    x, y = outer.createRandomVariable(5.0), outer.createRandomVariable(2.0)
    result = x.mult(a).pow(2).add(y.mult(b).pow(3))          # (a x)^2 + (b y)^3
    dx = result.getGradient()[x.getID()]
    assert abs(dx.getAverage() - 2 * 25.0 * 5.0) < 1e-12
    ddxx = dx.getGradient()[x.getValues().getID()]           # the inner independent
    assert abs(ddxx.getAverage() - 2 * 25.0) < 1e-13
    dy = result.getGradient()[y.getID()]
    assert abs(dy.getGradient()[y.getValues().getID()].getAverage() - 6 * 2.0) < 1e-13


def test_type_priority_upon_commutation(env):
    pkg, RV, Factory, f = env[:4]
    AAD = pkg.RandomVariableDifferentiableAAD
    x, y = f.createRandomVariable(2.0), Factory().createRandomVariable(3.0)
    s = pkg.Scalar(3.0)
    assert x.getTypePriority() == 3 and x.getValues().getTypePriority() == env.priority and y.getValues() is y
    for other in (y, s):
        for name in ("add", "sub", "bus", "mult", "div", "vid", "cap", "floor"):
            z1, z2 = getattr(x, name)(other), getattr(other, name)(x)
            assert type(z1) is AAD and type(z2) is AAD, name
            mirror = {"sub": "bus", "bus": "sub", "div": "vid", "vid": "div"}.get(name, name)
            assert getattr(x, mirror)(other).getAverage() == z2.getAverage(), name
            assert x.getID() in z2.getGradient()
        for z in (other.accrue(x, 0.5), other.discount(x, 0.5), other.addProduct(x, other), other.addProduct(other, x), other.addProduct(x, 2.0),
                  other.addRatio(x, other), other.addRatio(other, x), other.subRatio(x, other), other.subRatio(other, x)):
            assert type(z) is AAD and x.getID() in z.getGradient()
    assert other.accrue(x, 0.5).getAverage() == 3.0 * (1 + 2.0 * 0.5) and abs(y.discount(x, 0.5).getAverage() - 1.5) < 1e-15
    assert y.subRatio(x, y).getAverage() == 3.0 - 2.0 / 3.0
    # end points are plain numbers / plain random variables
    assert isinstance(x.getAverage(), float) and type(x.isNaN()) is not AAD
    assert x.getCloneIndependent().getID() != x.getID() and x.getCloneIndependent().getValues() is x.getValues()


def test_retention_rules_and_leaf_only_gradients(env):
    pkg, RV, Factory, f = env[:4]
    x, z = f.createRandomVariable(0.0, X1), f.createRandomVariable(0.0, X2)
    k = RV(0.0, X2)
    node = lambda r: r.getOperatorTreeNode()
    assert node(x).arguments is None and node(x).argumentValues is None
    assert node(x.add(z)).argumentValues is None and node(x.sub(3.0)).argumentValues is None and node(x.average()).argumentValues is None
    m = node(x.mult(k))
    assert m.argumentValues[0] is None and m.argumentValues[1] is k and m.arguments[1] is None
    assert node(x.mult(2.0)).argumentValues[0] is None
    assert node(x.div(k)).argumentValues[0] is None and node(k.div(x)).argumentValues[0] is k
    ap = node(x.addProduct(k, z))
    assert ap.argumentValues[0] is None and ap.argumentValues[2] is None and ap.argumentValues[1] is k
    ch = node(k.choose(x, z)) if hasattr(k.choose(x, z), "getOperatorTreeNode") else None
    assert ch is None                                        # a non-differentiable trigger does not delegate (as in the reference): plain values
    c2 = node(x.choose(k, k))
    assert c2.argumentValues[1] is k
    # all vertices vs leaves only
    fAll = pkg.RandomVariableDifferentiableAADFactory(Factory(), {"isGradientRetainsLeafNodesOnly": False})
    a = fAll.createRandomVariable(0.0, X1)
    b = a.squared()
    c = b.exp()
    g = c.getGradient()
    assert sorted(g) == [a.getID(), b.getID(), c.getID()]
    assert np.allclose(g[b.getID()].getRealizations(), np.exp(X1 * X1), rtol=4e-16, atol=0) and g[c.getID()].doubleValue() == 1.0
    assert np.allclose(g[a.getID()].getRealizations(), 2 * X1 * np.exp(X1 * X1), rtol=1e-15, atol=0)
    assert sorted(c.getGradient({b.getID()})) == [b.getID()]
    # ids grow with creation order; properties and their defaults
    assert a.getID() < b.getID() < c.getID()
    d = pkg.RandomVariableDifferentiableAADFactory(Factory())
    assert (d.getDiracDeltaApproximationMethod(), d.getDiracDeltaApproximationWidthPerStdDev(), d.getDiracDeltaApproximationDensityRegressionWidthPerStdDev(),
            d.isGradientRetainsLeafNodesOnly()) == ("DISCRETE_DELTA", 0.05, 0.5, True)
    assert pkg.RandomVariableDifferentiableAADFactory(Factory(), {"barrierDiracWidth": 0.2}).getBarrierDiracWidth() == 0.2
    with pytest.raises(ValueError):
        pkg.RandomVariableDifferentiableAADFactory(Factory(), {"diracDeltaApproximationMethod": "NOPE"})


def test_indicator_derivative_methods(env):
    pkg, RV, Factory, _ = env[:4]
    rng = np.random.default_rng(11)
    xv, yv, zv = rng.standard_normal(100001), rng.uniform(1, 2, 100001), rng.uniform(-1, 0, 100001)
    for method in ("ONE", "ZERO", "DISCRETE_DELTA"):
        f = pkg.RandomVariableDifferentiableAADFactory(Factory(), {"diracDeltaApproximationMethod": method, "diracDeltaApproximationWidthPerStdDev": 0.1})
        x, y, z = (f.createRandomVariable(0.0, v) for v in (xv, yv, zv))
        g = x.choose(y, z).getGradient()
        assert np.array_equal(g[y.getID()].getRealizations(), (xv >= 0) * 1.0) and np.array_equal(g[z.getID()].getRealizations(), (xv < 0) * 1.0)
        got = g[x.getID()]
        if method == "ONE":
            assert np.array_equal(got.getRealizations(), yv - zv)
        elif method == "ZERO":
            assert got.isDeterministic() and got.doubleValue() == 0.0
        else:
            eps = 0.1 * RV(0.0, xv).getStandardDeviation()
            want = (yv - zv) * ((xv + eps / 2 >= 0) * 1.0) * ((xv - eps / 2 < 0) * 1.0) / eps
            assert np.array_equal(got.getRealizations(), want)
            # E[(y - z) δ(x)] ≈ E[y - z] φ(0): the discrete delta is a consistent estimator of the density-weighted jump
            assert abs(np.mean(want) - np.mean(yv - zv) / np.sqrt(2 * np.pi)) < 0.05


# ---- models recorded by the generic Euler recipe -------------------------------------------------------------------------------------
def _black_scholes(env, factory, s0, r, sigma, td, paths, seed=3141):
    pkg = env.pkg
    bm = env.BM(td, 1, paths, seed, env.Factory())           # plain increments; the model parameters carry the differentiability
    model = pkg.BlackScholesModel(s0, r, sigma, factory)
    return model, pkg.MonteCarloAssetModel(model, pkg.EulerSchemeFromProcessModel(model, bm))


def test_black_scholes_greeks_through_the_recorded_simulation(env):
    from scipy.stats import norm
    pkg = env.pkg
    paths = 200_000 if env.device else 20_000
    td = pkg.TimeDiscretizationFromArray(0.0, 10, 0.5)
    s0, r, sigma, T, K = 1.0, 0.05, 0.30, 5.0, 1.05
    option = pkg.EuropeanOption(T, K)
    model, mc = _black_scholes(env, env.f, s0, r, sigma, td, paths)
    value = option.getValueRV(0.0, mc)
    assert type(value) is pkg.RandomVariableDifferentiableAAD and mc.getProcess().usedFusedKernel is None
    g = value.getGradient()
    ids = [model.getInitialValue()[0].getID(), model.getRiskFreeRate().getID(), model.getVolatility().getID()]
    assert sorted(g) == sorted(ids)
    aad = [g[i].getAverage() for i in ids]
    # the same paths, bumped parameters, plain factory: on the device this is the FUSED kernel (the recorded recipe must agree with it)
    plain, mcPlain = _black_scholes(env, env.Factory(), s0, r, sigma, td, paths)
    assert abs(option.getValue(mcPlain) - value.getAverage()) < 1e-12
    assert mcPlain.getProcess().usedFusedKernel == ("black_scholes" if env.device else None)
    eps = 1e-5

    def smooth(mc):                                          # a pay-off without a kink: the bumped runs differentiate it to O(eps^2)
        v = mc.getAssetValue(T, 0)
        return v.log().squared().add(v.sqrt()).div(mc.getNumeraire(T))
    gs = smooth(mc).getGradient()
    for k, bump in enumerate(((eps, 0, 0), (0, eps, 0), (0, 0, eps))):
        up = _black_scholes(env, env.Factory(), s0 + bump[0], r + bump[1], sigma + bump[2], td, paths)[1]
        dn = _black_scholes(env, env.Factory(), s0 - bump[0], r - bump[1], sigma - bump[2], td, paths)[1]
        fd = (smooth(up).getAverage() - smooth(dn).getAverage()) / (2 * eps)
        assert abs(fd - gs[ids[k]].getAverage()) < 1e-8 * max(1.0, abs(fd)), (k, fd, gs[ids[k]].getAverage())
        # the call: the ~N p(K) |dS/dθ| 2 eps paths whose kink lies inside the bump make the difference quotient noisy (not the adjoint)
        fd = (option.getValue(up) - option.getValue(dn)) / (2 * eps)
        assert abs(fd - aad[k]) < 2e-4 * max(1.0, abs(aad[k])), (k, fd, aad[k])
    d1 = (np.log(s0 / K) + (r + 0.5 * sigma * sigma) * T) / (sigma * np.sqrt(T))
    d2 = d1 - sigma * np.sqrt(T)
    analytic = [norm.cdf(d1), K * T * np.exp(-r * T) * norm.cdf(d2), s0 * np.sqrt(T) * norm.pdf(d1)]
    tol = 0.02 if env.device else 0.06                      # Monte-Carlo error of the pathwise estimators (vega is the noisy one)
    for got, want in zip(aad, analytic):
        assert abs(got - want) < tol * max(1.0, abs(want)), (aad, analytic)


def test_digital_option_delta(env):
    """T/montecarlo/automaticdifferentiation/MonteCarloBlackScholesModelDigitalOptionAADRegressionSensitivitiesTest.java:53-85 (its set-up
    and its tolerances: 1e-2 for the discrete delta, 4e-3 for the regression on the distribution)."""
    pkg = env.pkg
    paths = 200_000
    td = pkg.TimeDiscretizationFromArray(0.0, 1, 1.0)
    s0, r, sigma, T, K = 1.0, 0.05, 0.50, 1.0, 1.05
    dPlus = (np.log(s0 / K) + (r + 0.5 * sigma * sigma) * T) / (sigma * np.sqrt(T))
    dMinus = dPlus - sigma * np.sqrt(T)
    analytic = np.exp(-r * T) * np.exp(-0.5 * dMinus * dMinus) / (np.sqrt(2.0 * np.pi * T) * s0 * sigma)
    option = pkg.DigitalOption(T, K)
    # on the device seed 3141 reproduces the reference's own sample (bit-exact uniforms), so its tolerances apply as they are; the numpy
    # stand-in draws other normals: three standard errors of the windowed estimator
    cases = [({"diracDeltaApproximationWidthPerStdDev": 0.05}, 1e-2 if env.device else 4e-2)]
    if env.device:                                           # the regressions solve on the device (LinearRegression -> fmb_regression_solve_svd)
        cases += [({"diracDeltaApproximationWidthPerStdDev": 0.05, "diracDeltaApproximationMethod": "REGRESSION_ON_DISTRIBUITON",
                    "diracDeltaApproximationDensityRegressionWidthPerStdDev": 0.75}, 4e-3),
                  ({"diracDeltaApproximationWidthPerStdDev": 0.05, "diracDeltaApproximationMethod": "REGRESSION_ON_DENSITY",
                    "diracDeltaApproximationDensityRegressionWidthPerStdDev": 0.75}, 1e-2)]
    for props, tol in cases:
        f = pkg.RandomVariableDifferentiableAADFactory(env.Factory(), props)
        model, mc = _black_scholes(env, f, s0, r, sigma, td, paths)
        value = option.getValueRV(0.0, mc)
        delta = value.getGradient()[model.getInitialValue()[0].getID()].getAverage()
        assert abs(delta - analytic) < tol, (props, delta, analytic)
    # zero / infinite width (:178-190): no contribution / the jump itself
    for width, want in ((0.0, 0.0), (float("inf"), None)):
        f = pkg.RandomVariableDifferentiableAADFactory(env.Factory(), {"diracDeltaApproximationWidthPerStdDev": width})
        model, mc = _black_scholes(env, f, s0, r, sigma, td, 1000)
        d = option.getValueRV(0.0, mc).getGradient()[model.getInitialValue()[0].getID()]
        if want is not None:
            assert d.getAverage() == want
        else:
            assert d.getAverage() > 0.5                      # E[dS/dS0 · e^{-rT}] ≈ e^{-rT} E[S_T]/S_0-ish: the un-localised jump


def test_conditional_expectation_operator(env):
    """The adjoint of E(·|F) is E(·|F) with the same estimator (:185-189; ssrn 2995695): checked through a Bermudan-style step
    value = max(exercise, E(continuation | S_1)) against bumping the initial value on the same paths."""
    if not env.device:
        pytest.skip("the regression estimator is device code")
    pkg = env.pkg
    paths = 100_000
    td = pkg.TimeDiscretizationFromArray(0.0, 2, 1.0)
    s0, r, sigma, K = 1.0, 0.05, 0.30, 1.0

    def bermudan(factory, s0, frozenTrigger=None):
        model, mc = _black_scholes(env, factory, s0, r, sigma, td, paths)
        s1, s2 = mc.getAssetValue(1, 0), mc.getAssetValue(2, 0)
        continuation = s2.bus(K).floor(0.0).div(mc.getNumeraire(2.0))           # put pay-off at t = 2
        exercise = s1.bus(K).floor(0.0).div(mc.getNumeraire(1.0))
        s1v = s1.getValues()
        estimator = pkg.MonteCarloConditionalExpectationRegression([s1v.mult(0.0).add(1.0), s1v, s1v.squared()])
        expected = continuation.getConditionalExpectation(estimator)
        trigger = expected.sub(exercise) if frozenTrigger is None else frozenTrigger
        return model, trigger.choose(continuation, exercise), trigger

    # Dirac method ZERO = the pathwise derivative at FIXED exercise decisions.  (Bumped runs that re-decide move paths across the
    # boundary; with a quadratic regression of a kinked pay-off the rule is not optimal there, so that term does not vanish: measured
    # -0.327 / -0.360 for bumps of 1e-4 / 2e-2 against -0.345.)  The comparison therefore freezes the decisions of the base run.
    fZero = pkg.RandomVariableDifferentiableAADFactory(env.Factory(), {"diracDeltaApproximationMethod": "ZERO"})
    model, value, trigger = bermudan(fZero, s0)
    assert type(trigger) is pkg.RandomVariableDifferentiableAAD and trigger.getOperatorTreeNode().arguments[0].operatorType == "CONDITIONAL_EXPECTATION"
    delta = value.getGradient()[model.getInitialValue()[0].getID()].getAverage()
    eps = 1e-5
    frozen = trigger.getValues()
    up = bermudan(env.Factory(), s0 + eps, frozen)[1].getAverage()
    dn = bermudan(env.Factory(), s0 - eps, frozen)[1].getAverage()
    fd = (up - dn) / (2 * eps)
    assert delta < 0 and abs(fd - delta) < 2e-4, (fd, delta)  # (put kinks inside the bump: see the Black-Scholes test)
    # with the default discrete delta the boundary term is part of the adjoint: the estimate moves, by less than its noise band
    model, value, _ = bermudan(env.f, s0)
    withBoundary = value.getGradient()[model.getInitialValue()[0].getID()].getAverage()
    assert withBoundary != delta and abs(withBoundary - delta) < 0.1
    # the operator itself: d/dx E[ E(x·W² | W) ] = E[W²]
    w = pkg.BrownianMotionCuda(td, 1, paths, 77).getBrownianIncrement(0, 0)
    x = env.f.createRandomVariable(2.0)
    est = pkg.MonteCarloConditionalExpectationRegression([w.mult(0.0).add(1.0), w, w.squared()])
    y = x.mult(w.squared()).getConditionalExpectation(est)
    d = y.getGradient()[x.getID()]
    assert abs(d.getAverage() - w.squared().getAverage()) < 1e-12
    assert np.allclose(d.getRealizations(), w.squared().getRealizations(), rtol=0, atol=1e-9)     # W² lies in the span of the basis


def test_lmm_forward_rate_deltas(env):
    """Caplet and swaption deltas w.r.t. the initial forward rates of a small LIBOR market model, recorded through the generic recipe
    (drift prefix sums, numeraire, interpolation) against bumping the curve on the same paths (on the device: the FUSED kernel)."""
    from common import lmm_setup
    pkg = env.pkg
    paths = 20_000 if env.device else 2_000
    s = lmm_setup(pkg, n_libors=6, n_factors=2, period=0.5, dt=0.5)

    def simulation(factory, L0):
        # no separate discount curve: with one, the numeraire adjustment P_forward(T) / P_discount(T) is computed from the host-side
        # curve (not recorded), and a bump of L0 would cancel there but not in the recorded numeraire
        model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, L0, None, factory, s["cov"], None, {"measure": "SPOT"})
        bm = env.BM(s["sim"], s["F"], paths, 3141, env.Factory())
        return model, pkg.LIBORMonteCarloSimulationFromLIBORModel(pkg.EulerSchemeFromProcessModel(model, bm))

    products = [pkg.Caplet(1.0, 0.5, 0.05), pkg.Swaption(1.0, [1.0, 1.5, 2.0], [1.5, 2.0, 2.5], [0.05] * 3)]
    model, sim = simulation(env.f, s["L0"])
    for product in products:
        value = product.getValueRV(0.0, sim)
        assert type(value) is pkg.RandomVariableDifferentiableAAD
        g = value.getGradient()
        state = model.getInitialState(sim.getProcess())
        plainValue = product.getValue(simulation(env.Factory(), s["L0"])[1])
        assert abs(plainValue - value.getAverage()) < 1e-12 * max(1.0, abs(plainValue))
        for j in (0, 2, 3, 5):
            # the independents are the log forward rates (log-normal state space): dV/dL_j = dV/dlog L_j / L_j
            aad = g[state[j].getID()].getAverage() / s["L0"][j] if state[j].getID() in g else 0.0
            eps = 1e-6
            up, dn = s["L0"].copy(), s["L0"].copy()
            up[j] += eps
            dn[j] -= eps

            fd = (product.getValue(simulation(env.Factory(), up)[1]) - product.getValue(simulation(env.Factory(), dn)[1])) / (2 * eps)
            assert abs(fd - aad) < 1e-6 * max(1.0, abs(fd)), (type(product).__name__, j, fd, aad)
