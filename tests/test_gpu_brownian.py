"""-m gpu: MT19937 jump-ahead + AS241 Brownian kernel against the oracle, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from common import rel_err

pytestmark = pytest.mark.gpu


def _words(nv, seed, off, n):
    out = np.empty(n, dtype=np.uint32)
    nv.check(nv.load().fmb_mt_words(seed, off, n, out.ctypes.data_as(nv.c_u32p)))
    return out


@pytest.mark.parametrize("seed,offset,n", [(3141, 0, 5000), (3141, 1, 700), (31415, 623, 1300), (-1, 123457, 2000),
                                           (53252, 240 * 13524, 1000), (3141, 3_000_000_001, 1500)])
def test_mt_words_bit_exact(gpu, orc, seed, offset, n):
    got = _words(gpu.native, seed, offset, n)
    assert np.array_equal(got, orc.mt_words(seed, offset, n))


def test_mt_uniforms_bit_exact(gpu, orc):
    nv = gpu.native
    for seed, off, n in [(3141, 0, 4096), (3141, 999_999, 3000), (-7, 5, 10)]:
        u = np.empty(n)
        nv.check(nv.load().fmb_mt_uniforms(seed, off, n, nv.dptr(u)))
        ref = orc.mt_uniforms(seed, off, n)
        assert np.array_equal(u.view(np.uint64), ref.view(np.uint64))
    u = np.empty(4)
    nv.check(nv.load().fmb_mt_uniforms(3141, 0, 4, nv.dptr(u)))
    assert list(u) == [0.48112854170930563, 0.8554104072506368, 0.7730656542235694, 0.9557774714716378]   # tests/golden/mt_as241.json


def test_icdf_device_matches_oracle(gpu, orc):
    nv = gpu.native
    rng = np.random.default_rng(7)
    p = np.concatenate([rng.random(200000), [0.0, 1.0, 0.5, 0.075, 0.925, 1e-300, 1 - 2.0 ** -52, 2.0 ** -52, 0.074999, 0.925001]])
    out = np.empty_like(p)
    nv.check(nv.load().fmb_icdf(nv.dptr(p), p.size, nv.dptr(out)))
    ref = orc.icdf(p)
    central = np.abs(p - 0.5) <= 0.425
    # central branch has no transcendental: bit-exact with the no-FMA oracle
    assert np.array_equal(out[central].view(np.uint64), ref[central].view(np.uint64))
    assert rel_err(out, ref) < 2e-15                         # tails: device log() vs libm log() may differ by an ulp
    assert out[-10] == 0.0 and out[-9] == 0.0                # quirk: p = 0 or 1 -> 0.0 (NormalDistribution.java:141-143)


@pytest.mark.parametrize("T,F,P,off", [(7, 3, 1000, 0), (100, 1, 5000, 0), (40, 3, 4097, 0), (5, 2, 1, 0), (5, 2, 3, 0), (3, 1, 131, 0),
                                       (1000, 2, 37, 0), (40, 3, 2500, 1234), (13, 4, 777, 10_000_000),
                                       # bulk-store (TMA) flush: even path counts, full and ragged last tiles, narrow tiles (T*F = 2000: 4-path rows)
                                       (40, 3, 20_000, 0), (40, 3, 1002, 6), (1000, 2, 1000, 0), (1000, 2, 38, 2), (3, 2, 2, 0),
                                       # T*F beyond one path per tile (13 000 columns: 1-path tiles; 30 000: column chunks of the paths, no cap)
                                       (6500, 2, 6, 0), (15000, 2, 5, 1), (10000, 3, 4, 0)])
def test_brownian_increments_match_oracle(gpu, orc, T, F, P, off):
    nv = gpu.native
    td = gpu.TimeDiscretizationFromArray(0.0, T, 0.1 if T < 1000 else 0.005)
    sq = np.sqrt(np.diff(td.times))
    out = np.zeros(T * F, dtype=np.uint64)
    nv.check(nv.load().fmb_bm_generate(3141, T, F, P, off, nv.dptr(sq), nv.hptr(out)))
    got = np.stack([nv.DeviceVector(int(h), P).download() for h in out]).reshape(T, F, P)
    ref = orc.brownian(3141, td.times, F, P, path_offset=off)
    assert rel_err(got, ref) < 2e-15                         # tail draws: device log() vs libm log(), a few ulp after the rational
    assert np.mean(got.view(np.uint64) == ref.view(np.uint64)) > 0.8      # everything but some tail draws is bit-identical


def test_brownian_motion_interface(gpu, orc):
    td = gpu.TimeDiscretizationFromArray(0.0, 10, 0.1)
    bm = gpu.BrownianMotionCuda(td, 2, 1000, 3141)
    inc = bm.getBrownianIncrement(3, 1)
    assert inc.getFiltrationTime() == td.getTime(4) and inc.size() == 1000 and not inc.isDeterministic()
    ref = orc.brownian(3141, td.times, 2, 1000)
    assert rel_err(inc.getRealizations(), ref[3, 1]) < 2e-15
    assert bm.getIncrement(3)[1] is inc
    clone = bm.getCloneWithModifiedSeed(31415)
    assert clone.getSeed() == 31415 and clone != bm and bm == gpu.BrownianMotionCuda(td, 2, 1000, 3141)
    assert bm.getRandomVariableForConstant(2.0).doubleValue() == 2.0


def test_brownian_statistics_like_reference_tests(gpu):
    """T/montecarlo/BrownianMotionTest.java:97-136 (mean / variance within 3 sigma) and :183-241 (sum dW^2 ~ t), at 10^6 paths."""
    P, T = 1_000_000, 10
    td = gpu.TimeDiscretizationFromArray(0.0, T, 1.0)
    bm = gpu.BrownianMotionCuda(td, 1, P, 53252)
    s = bm.getBrownianIncrement(0, 0).squared()
    for t in range(1, T):
        s = s.add(bm.getBrownianIncrement(t, 0).squared())
    assert abs(s.getAverage() - T) < 3 * np.sqrt(2.0 * T / P) * 1.5
    w = bm.getBrownianIncrement(4, 0)
    assert abs(w.getAverage()) < 3.0 / np.sqrt(P) and abs(w.getVariance() - 1.0) < 3 * np.sqrt(2.0 / P) * 1.5


def test_brownian_view_and_correlated(gpu, orc):
    """BrownianMotionView / CorrelatedBrownianMotion on device increments; Heston on a view (HestonModelTest.java:98-100)."""
    td = gpu.TimeDiscretizationFromArray(0.0, 20, 0.25)
    bm = gpu.BrownianMotionCuda(td, 3, 2000, 3141)
    ref = orc.brownian(3141, td.times, 3, 2000)
    view = gpu.BrownianMotionView(bm, [2, 0])
    assert view.getNumberOfFactors() == 2 and view.getBrownianIncrement(5, 0) is bm.getBrownianIncrement(5, 2)
    rho = 0.3
    corr = gpu.CorrelatedBrownianMotion(bm, [[1.0, 0.0, 0.0], [rho, np.sqrt(1 - rho * rho), 0.0]])
    w1 = corr.getBrownianIncrement(4, 1).getRealizations()
    expect = (0.0 + ref[4, 0] * rho) + ref[4, 1] * np.sqrt(1 - rho * rho)
    assert rel_err(w1, expect) < 2e-15
    # an Euler scheme on a view runs the generic device loop and agrees with the fused kernel on the same factor
    model = gpu.BlackScholesModel(1.0, 0.05, 0.3, bm.randomVariableFactory)
    a = gpu.EulerSchemeFromProcessModel(model, gpu.BrownianMotionView(bm, [0])).getProcessValue(20, 0).getRealizations()
    b = gpu.EulerSchemeFromProcessModel(model, bm).getProcessValue(20, 0).getRealizations()
    assert np.array_equal(a, b)


@pytest.mark.parametrize("T,F,P,off", [(5, 2, 1000, 0), (40, 3, 2001, 0), (7, 1, 64, 12345), (1000, 2, 10, 3)])
def test_uniform_increments_are_bit_exact_in_draw_order(gpu, orc, T, F, P, off):
    """fmb_uniforms_generate: out[t*F+f][p] = u_{((off+p)*T+t)*F+f} of `new MersenneTwister(seed)` (IndependentIncrementsFromICDF.java:186-194)."""
    nv = gpu.native
    out = np.zeros(T * F, dtype=np.uint64)
    nv.check(nv.load().fmb_uniforms_generate(53252, T, F, P, off, nv.hptr(out)))
    got = np.stack([nv.DeviceVector(int(h), P).download() for h in out]).reshape(T, F, P)
    seq = orc.mt_uniforms(53252, off * T * F, P * T * F).reshape(P, T, F)
    assert np.array_equal(got, np.transpose(seq, (1, 2, 0)))


def test_independent_increments_from_icdf(gpu, orc):
    """IndependentIncrementsFromICDF (J/montecarlo/IndependentIncrementsFromICDF.java:173-206): per (time, factor) inverse distribution
    functions on device-generated uniforms - the built-in normal transform, a function written with RandomVariable operations
    (exponential distribution) and a plain double -> double callable (applied on the host)."""
    import math
    T, F, P = 6, 3, 4000
    td = gpu.TimeDiscretizationFromArray(0.0, T, 0.25)
    lam = 2.5
    exponential = lambda u: u.mult(-1.0).add(1.0).log().mult(-1.0 / lam)
    plain = lambda x: x * x                                  # double -> double
    icdfs = lambda t: (lambda f: [gpu.IndependentIncrementsFromICDF.NORMAL_ICDF, exponential, plain][f])
    inc = gpu.IndependentIncrementsFromICDF(td, F, P, 3141, icdfs)
    u = np.transpose(orc.mt_uniforms(3141, 0, P * T * F).reshape(P, T, F), (1, 2, 0))
    for t in (0, 3, 5):
        z = inc.getIncrement(t, 0)
        assert z.getFiltrationTime() == td.getTime(t + 1)
        assert rel_err(z.getRealizations(), orc.icdf(u[t, 0])) < 2e-15
        assert rel_err(inc.getIncrement(t, 1).getRealizations(), -np.log(1.0 - u[t, 1]) / lam) < 1e-14
        assert np.array_equal(inc.getIncrement(t, 2).getRealizations(), u[t, 2] * u[t, 2])
    clone = inc.getCloneWithModifiedSeed(31415)
    assert not np.array_equal(clone.getIncrement(0, 0).getRealizations(), inc.getIncrement(0, 0).getRealizations())


def test_brownian_motion_from_random_number_generator(gpu, orc):
    """BrownianMotionFromRandomNumberGenerator (J/montecarlo/BrownianMotionFromRandomNumberGenerator.java:137-186): with the Mersenne
    Twister as generator it is the same stream as BrownianMotionFromMersenneRandomNumbers (device-generated); any other generator is a
    host object asked path by path."""
    T, F, P = 8, 2, 3000
    td = gpu.TimeDiscretizationFromArray(0.0, T, 0.5)
    mt = gpu.RandomNumberGeneratorFrom1D(gpu.MersenneTwister(3141), T * F)
    bm = gpu.BrownianMotionFromRandomNumberGenerator(td, F, P, mt)
    ref = gpu.BrownianMotionCuda(td, F, P, 3141)
    for t, f in ((0, 0), (3, 1), (7, 0)):
        assert np.array_equal(bm.getBrownianIncrement(t, f).getRealizations(), ref.getBrownianIncrement(t, f).getRealizations())
        assert bm.getBrownianIncrement(t, f).getFiltrationTime() == td.getTime(t + 1)
    # an Euler scheme on top of it (generic device loop) gives the paths of the fused kernel on the Mersenne driver
    model = gpu.BlackScholesModel(1.0, 0.05, 0.3, ref.randomVariableFactory)
    a = gpu.EulerSchemeFromProcessModel(model, gpu.BrownianMotionView(bm, [0])).getProcessValue(T, 0).getRealizations()
    b = gpu.EulerSchemeFromProcessModel(model, gpu.BrownianMotionCuda(td, F, P, 3141)).getProcessValue(T, 0).getRealizations()
    assert np.array_equal(a, b)

    class Halton2:                                           # any host-side generator: here a van der Corput / Halton pair per step
        def __init__(self, dim): self.dim, self.i = dim, 0
        def getDimension(self): return self.dim
        def getNext(self):
            self.i += 1
            out = []
            for d in range(self.dim):
                base, f, r, i = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53][d], 1.0, 0.0, self.i
                while i > 0:
                    f /= base
                    r += f * (i % base)
                    i //= base
                out.append(r)
            return out
    q = gpu.BrownianMotionFromRandomNumberGenerator(td, F, 500, Halton2(T * F))
    h = Halton2(T * F)
    u = np.array([h.getNext() for _ in range(500)])
    for t, f in ((0, 0), (5, 1)):
        want = orc.icdf(u[:, t * F + f]) * np.sqrt(td.getTimeStep(t))
        assert rel_err(q.getBrownianIncrement(t, f).getRealizations(), want) < 2e-15
