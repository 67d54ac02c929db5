"""-m gpu: RandomVariableCuda arithmetic and reductions against the oracle's RandomVariableFromDoubleArray semantics."""
import numpy as np
import pytest

from common import rel_err, same_bits

pytestmark = pytest.mark.gpu


def _vectors(n, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n)
    y = rng.standard_normal(n) * 3.0 + 0.5
    z = rng.random(n) + 0.25
    sx = [0.0, -0.0, np.nan, np.inf, -np.inf, 1.0, -1.0, 2.0]
    sy = [-0.0, 0.0, 1.0, np.nan, 2.0, np.inf, -0.0, 0.5]
    m = min(n, 8)
    x[:m], y[:m] = sx[:m], sy[:m]
    return x, y, z


@pytest.mark.parametrize("n", [1, 2, 3, 1000, 100_003])
def test_unary_ops_bit_exact_or_1ulp(gpu, orc, n):
    nv = gpu.native
    x, _, z = _vectors(n)
    exact_ops = [(0, 0), (1, 0), (6, 0), (7, 0), (8, 0), (10, 0.3), (11, 0.3), (12, 0.3), (13, -1.7), (14, 3.0), (15, 2.0), (16, 0.1), (17, -0.2),
                 (16, 0.0), (17, 0.0), (18, 2.0), (18, 0.5)]
    for op, a in exact_ops:
        src = np.abs(x) if op in (1,) or (op == 18 and a == 0.5) else x
        got = nv.unary(op, nv.DeviceVector.upload(src), a).download()
        ref = orc.rv_unary(op, src, a) if not (op == 18) else (src * src if a == 2.0 else np.sqrt(src))
        assert same_bits(got, ref), (op, a)
    for op, a, src in [(2, 0, x), (3, 0, z), (4, 0, x), (5, 0, x), (9, 0, x), (18, 1.7, z), (18, -2.5, z), (18, 3.0, x)]:
        got = nv.unary(op, nv.DeviceVector.upload(src), a).download()
        assert rel_err(got, orc.rv_unary(op, src, a), scale=1e-300) < 5e-16, (op, a)


@pytest.mark.parametrize("n", [1, 7, 4096, 100_003])
def test_binary_ternary_ops_bit_exact(gpu, orc, n):
    nv = gpu.native
    x, y, z = _vectors(n)
    dx, dy, dz = (nv.DeviceVector.upload(v) for v in (x, y, z))
    for op in range(6):
        got = nv.binary(op, dx, 0.0, dy, 0.0).download()
        assert same_bits(got, orc.rv_binary(op, x, y)), op
        got = nv.binary(op, None, 0.75, dy, 0.0).download()                      # scalar broadcast (deterministic receiver)
        assert same_bits(got, orc.rv_binary(op, np.full(n, 0.75), y)), op
    for op, a in [(0, 0), (1, 0.37), (2, 0), (3, 0), (4, 0.5), (5, 0.5), (6, 0)]:
        got = nv.ternary(op, dx, 0.0, dy, 0.0, dz if op in (0, 2, 3, 6) else None, 0.0, a).download()
        ref = orc.rv_ternary(op, x, y, z if op in (0, 2, 3, 6) else None, a)
        assert same_bits(got, ref), op


def test_reference_randomvariable_identities(gpu):
    """T/montecarlo/RandomVariableTest.java:53-142 restated on the device type."""
    RV = gpu.RandomVariableCuda
    rv = RV(0.0, [3.0, 1.0, 0.0, 2.0, 4.0, 1.0 / 3.0])
    assert np.array_equal(rv.sqrt().getRealizations(), rv.pow(0.5).getRealizations())          # tolerance 0.0 in the reference
    assert np.array_equal(rv.squared().getRealizations(), rv.pow(2.0).getRealizations())
    assert rv.getStandardDeviation() == np.sqrt(rv.getVariance())
    c = RV(0.0, 2.0)                                                                             # testRandomVariableDeterministc :53-66
    c = c.mult(2.0)
    assert c.doubleValue() == 4.0 and c.getAverage() == 4.0
    c = c.div(8.0)
    assert c.getAverage() == 0.5 and c.getVariance() == 0.0
    s = RV(0.0, [-4.0, -2.0, 0.0, 2.0, 4.0])                                                   # testRandomVariableStochastic :68-99
    s = s.add(4.0)
    assert s.getAverage() == 4.0
    s = s.div(2.0)
    assert s.getAverage() == 2.0
    assert s.getVariance() == 2.0                                                               # (4+1+0+1+4)/5
    assert abs(s.squared().sub(s.getAverage() ** 2).getAverage() - 2.0) == 0.0
    with pytest.raises(NotImplementedError):
        RV(0.0, [1.0, 2.0]).doubleValue()
    with pytest.raises(NotImplementedError):
        rv.apply(lambda v: v)


def test_type_priority_and_deterministic_branches(gpu, orc):
    RV, Scalar, CPU = gpu.RandomVariableCuda, gpu.Scalar, gpu.RandomVariableFromDoubleArray
    x = np.linspace(-1.0, 2.0, 1001)
    g = RV(1.0, x)
    assert g.getTypePriority() == 2 and Scalar(1.0).getTypePriority() == 0 and CPU(1.0).getTypePriority() == 1
    # Scalar receiver delegates with the reference's re-ordered arithmetic (Scalar.java:276-357)
    assert np.array_equal(Scalar(0.3).sub(g).getRealizations(), (x - 0.3) * -1.0)
    assert np.array_equal(Scalar(0.3).div(g).getRealizations(), (1.0 / x) * 0.3)
    assert np.array_equal(Scalar(0.3).addProduct(g, g).getRealizations(), x * x + 0.3)
    assert np.array_equal(Scalar(0.5).discount(g, 0.5).getRealizations(), 1.0 / (x * (0.5 / 0.5) + 1.0 / 0.5))
    assert np.array_equal(Scalar(2.0).accrue(g, 0.5).getRealizations(), x * (0.5 * 2.0) + 2.0)
    # CPU type with a GPU argument: the GPU type takes over, result stays on the device
    r = CPU(1.0).mult(g)
    assert isinstance(r, RV) and np.array_equal(r.getRealizations(), 1.0 * x)
    # filtration time = max of operands; unary keeps the receiver's
    assert g.add(RV(2.5, x)).getFiltrationTime() == 2.5 and g.exp().getFiltrationTime() == 1.0
    assert g.add(Scalar(1.0)).getFiltrationTime() == 1.0
    # choose: NaN trigger -> negative branch, -0.0 -> non-negative branch (:1359)
    t = RV(0.0, [np.nan, -0.0, 0.0, -1.0, 1.0])
    assert list(t.choose(Scalar(1.0), Scalar(2.0)).getRealizations()) == [2.0, 1.0, 1.0, 2.0, 1.0]
    d = RV(0.0, -1.0)
    a, b = RV(0.0, x), RV(0.0, x)
    assert d.choose(a, b) is b                                                                   # deterministic trigger returns the argument itself
    # cap / floor are Math.min / Math.max
    m = RV(0.0, [np.nan, -0.0, 0.0, 1.0]).cap(0.0).getRealizations()
    assert np.isnan(m[0]) and np.signbit(m[1]) and not np.signbit(m[2]) and m[3] == 0.0


@pytest.mark.parametrize("n", [1, 2, 1000, 1_000_003])
def test_reductions_match_kahan_oracle(gpu, orc, n):
    RV = gpu.RandomVariableCuda
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) * 10.0 + 1e3
    w = rng.random(n)
    g, gw = RV(0.0, x), RV(0.0, w)
    tol = 4e-16
    assert abs(g.getAverage() - orc.rv_reduce(0, x)) <= tol * abs(orc.rv_reduce(0, x))
    assert abs(g.getAverage(gw) - orc.rv_reduce(1, x, w)) <= tol * abs(orc.rv_reduce(1, x, w))
    if n > 1:
        assert abs(g.getVariance() - orc.rv_reduce(2, x)) <= 1e-14 * orc.rv_reduce(2, x)
        assert abs(g.getVariance(gw) - orc.rv_reduce(3, x, w)) <= 1e-14 * orc.rv_reduce(3, x, w)
        assert abs(g.getStandardError() - orc.rv_reduce(8, x)) <= 1e-14 * orc.rv_reduce(8, x)
    else:
        assert g.getVariance() == 0.0
    assert g.getMin() == orc.rv_reduce(4, x) and g.getMax() == orc.rv_reduce(5, x)
    assert g.getQuantile(0.3) == orc.rv_reduce(9, x, a=0.3)
    pts = np.array([990.0, 1000.0, 1010.0])
    assert np.array_equal(g.getHistogram(pts), orc.rv_histogram(x, pts))


def test_order_statistics_by_radix_select_match_the_sorted_reference(gpu, orc):
    """getQuantile / getQuantileExpectation / getHistogram (RandomVariableFromDoubleArray.java:445-575 clone and sort): here a radix select,
    one counting pass and one range-sum pass - same numbers as the sort, including ties, signed zeros and NaN (Arrays.sort order)."""
    RV = gpu.RandomVariableCuda
    rng = np.random.default_rng(5)
    cases = [rng.standard_normal(100_003), np.round(rng.standard_normal(50_000), 1),                 # many ties
             np.array([0.0, -0.0, 1.0, -1.0, 0.0, -0.0, 5e-324, -5e-324]), np.array([3.0, np.nan, -np.inf, 2.0, np.inf, np.nan, -7.5]),
             np.full(1000, 2.5), rng.standard_normal(1_000_003) * 1e-300, np.array([42.0])]
    for x in cases:
        g = RV(0.0, x)
        s = np.sort(x)                                       # numpy sorts like Arrays.sort: NaN last (signed zeros compare equal: checked by value)
        n = x.size
        for q in (0.0, 0.001, 0.3, 0.5, 0.77, 0.999, 1.0):
            got, want = g.getQuantile(q), s[min(max(int(np.floor((n + 1) * q - 1 + 0.5)), 0), n - 1)]
            assert (got == want) or (np.isnan(got) and np.isnan(want)), (n, q, got, want)
            if not np.isnan(x).any():                        # (the oracle's std::sort is undefined with NaN in the data; numpy sorts like Arrays.sort)
                assert got == orc.rv_reduce(9, x, a=q)
        if not np.isnan(x).any():
            for q0, q1 in ((0.1, 0.9), (0.0, 1.0), (0.45, 0.55), (0.3, 0.3)):
                i0 = min(max(int(np.floor((n + 1) * q0 - 1 + 0.5)), 0), n - 1)
                i1 = min(max(int(np.floor((n + 1) * q1 - 1 + 0.5)), 0), n - 1)
                want = float(np.sum(s[i0:i1 + 1].astype(np.longdouble)) / (i1 - i0 + 1))
                got = g.getQuantileExpectation(q0, q1)
                assert abs(got - want) <= 1e-13 * max(abs(want), np.max(np.abs(s[i0:i1 + 1]))), (n, q0, q1, got, want)
            pts = np.array([-1.0, 0.0, 0.5, 2.5, 1.0])           # not ascending on purpose: each threshold is counted on its own
            assert np.array_equal(g.getHistogram(np.sort(pts)), orc.rv_histogram(x, np.sort(pts)))
    # signed zeros: -0.0 sorts below +0.0
    z = RV(0.0, np.array([0.0, -0.0, 0.0, -0.0]))
    assert np.signbit(z.getQuantile(0.0)) and not np.signbit(z.getQuantile(1.0))
    # many thresholds (chunked calls) and counts straight from the C ABI
    import ctypes as C
    nv = gpu.native
    x = rng.standard_normal(200_000)
    pts = np.linspace(-4, 4, 1201)
    h = RV(0.0, x).getHistogram(pts)
    assert np.array_equal(h, orc.rv_histogram(x, pts)) and abs(h.sum() - 1.0) < 1e-12
    cnt = np.zeros(3, dtype=np.uint64)
    probe = np.array([0.5, np.nan, -0.5])
    keep = RV(0.0, x)                                       # (the handle lives as long as its owner)
    nv.check(nv.load().fmb_rv_count_le(keep.dv.h, nv.dptr(probe), 3, cnt.ctypes.data_as(nv.c_hp)))
    assert cnt.tolist() == [int(np.sum(x <= 0.5)), 0, int(np.sum(x <= -0.5))]


def test_empty_and_nan_reductions(gpu):
    RV = gpu.RandomVariableCuda
    assert np.isnan(RV(0.0, np.array([])).getAverage())
    assert np.isnan(RV(0.0, [1.0, np.nan, 3.0]).getMin())
    assert RV(0.0, 3.0).getAverage() == 3.0 and RV(0.0, 3.0).getStandardError() == 0.0


def test_pool_reuse_and_handle_errors(gpu):
    import ctypes as C
    nv = gpu.native
    lib = nv.load()
    a = nv.DeviceVector.upload(np.arange(1000.0))
    h = a.h
    used0, cached0, live0 = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.fmb_pool_stats(C.byref(used0), C.byref(cached0), C.byref(live0))
    del a
    used1, cached1, live1 = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.fmb_pool_stats(C.byref(used1), C.byref(cached1), C.byref(live1))
    assert live1.value == live0.value - 1 and cached1.value >= cached0.value + 8000
    out = np.empty(1000)
    assert lib.fmb_rv_download(h, nv.dptr(out), 1000) == nv.FMB_EHANDLE                         # freed handle is rejected, not dereferenced
    with pytest.raises(ValueError):
        nv.binary(0, nv.DeviceVector.upload(np.zeros(3)), 0.0, nv.DeviceVector.upload(np.zeros(4)), 0.0)


def test_free_from_another_thread_while_a_vector_is_in_use(gpu):
    """The documented threading contract: GC / cleaner threads free handles while pool threads compute.  An entry point pins the handles
    it dereferences until its kernels are queued, so a concurrent fmb_rv_free can neither delete the vector under it nor recycle its block
    early: every call either succeeds with the right numbers or fails cleanly with FMB_EHANDLE."""
    import ctypes as C
    import threading
    nv = gpu.native
    lib = nv.load()
    n = 200_000
    src = np.arange(n, dtype=np.float64)
    errors = []

    def worker(h, results):
        out = C.c_uint64()
        for _ in range(400):
            rc = lib.fmb_rv_unary(nv.U_MULT, h, 2.0, C.byref(out))
            if rc == nv.FMB_EHANDLE:
                results.append("freed")
                return
            if rc != 0:
                errors.append(rc)
                return
            got = np.empty(n)
            if lib.fmb_rv_download(out.value, nv.dptr(got), n) != 0 or got[n - 1] != 2.0 * (n - 1) or got[1] != 2.0:
                errors.append("wrong values")
            lib.fmb_rv_free(out.value)
        results.append("done")

    for rep in range(6):
        h = C.c_uint64()
        nv.check(lib.fmb_rv_upload(nv.dptr(src), n, C.byref(h)))
        results = []
        ts = [threading.Thread(target=worker, args=(h.value, results)) for _ in range(3)]
        for t in ts:
            t.start()
        import time
        time.sleep(0.002 * rep)
        lib.fmb_rv_free(h.value)                            # "the cleaner" frees while the workers are busy
        for t in ts:
            t.join()
        assert not errors, errors
        assert len(results) == 3
    live = C.c_uint64()
    nv.check(lib.fmb_pool_stats(None, None, C.byref(live)))


def test_concurrent_callers_like_the_reference_thread_pool(gpu):
    """EulerSchemeFromProcessModel.java:199,:232-269 submits one task per component to a thread pool, and the JVM frees from GC
    threads: the C ABI must be re-entrant.  8 host threads hammer ops, reductions and frees concurrently (ctypes drops the GIL)."""
    import threading
    RV = gpu.RandomVariableCuda
    n = 200_003
    base = np.linspace(0.5, 2.5, n)
    errors = []

    def worker(k):
        try:
            x = RV(float(k), base * (k + 1))
            for it in range(30):
                y = x.mult(2.0).add(x).sub(x.mult(3.0)).add(1.0)          # == 1.0 everywhere (exact: 2x + x - 3x for these magnitudes up to rounding)
                z = x.log().exp().div(x)
                s = x.squared().getAverage()
                if abs(y.getAverage() - 1.0) > 1e-12 or abs(z.getAverage() - 1.0) > 1e-12:
                    errors.append(("value", k, it))
                ref = float(np.mean((base * (k + 1)) ** 2))
                if abs(s - ref) > 1e-12 * ref:
                    errors.append(("reduce", k, it, s, ref))
                del y, z
        except Exception as e:                                              # noqa: BLE001
            errors.append(("exception", k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


def _random_algebra(gpu, lazy, trials=50, n=1537):
    """The same pseudo-random sequence of RandomVariable operations, evaluated with or without deferral."""
    nvm = gpu.native
    rng = np.random.default_rng(7)
    base = [rng.normal(size=n) for _ in range(4)]
    base[1][:6] = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-310]
    base[2] = np.abs(base[2]) + 0.1
    keep = nvm.lazy_min_n()
    nvm.set_lazy(lazy, min_n=0)
    try:
        vs = [gpu.RandomVariableCuda(float(i), b) for i, b in enumerate(base)]
        rs = np.random.default_rng(11)
        outs, times = [], []
        l0 = nvm.launch_count()
        for _ in range(trials):
            x = vs[rs.integers(4)]
            for _ in range(int(rs.integers(1, 24))):
                c = int(rs.integers(0, 20))
                y, z = vs[rs.integers(4)], vs[rs.integers(4)]
                if c == 0: x = x.add(0.3)
                elif c == 1: x = x.sub(y)
                elif c == 2: x = x.bus(y)
                elif c == 3: x = x.mult(y)
                elif c == 4: x = x.div(y)
                elif c == 5: x = x.vid(1.5)
                elif c == 6: x = x.cap(0.7)
                elif c == 7: x = x.floor(y)
                elif c == 8: x = x.squared()
                elif c == 9: x = x.abs().sqrt()
                elif c == 10: x = x.addProduct(y, z)
                elif c == 11: x = x.addProduct(y, 0.25)
                elif c == 12: x = x.addRatio(y, z)
                elif c == 13: x = x.subRatio(z, y)
                elif c == 14: x = x.accrue(y, 0.5)
                elif c == 15: x = x.discount(y, 0.5)
                elif c == 16: x = y.choose(x, z)
                elif c == 17: x = x.pow(2.0).add(x.pow(0.5))          # x consumed twice; pow special cases
                elif c == 18: x = x.mult(x)                             # the same pending value in two operand positions
                else: x = y.sub(0.03).mult(0.5).div(x).mult(z)          # x extends somebody else's chain from the second operand position
            outs.append(x.getRealizations())
            times.append(x.getFiltrationTime())
        launches = nvm.launch_count() - l0
    finally:
        nvm.set_lazy(True, min_n=keep)
    return outs, times, launches


@pytest.mark.gpu
def test_deferred_chains_are_bit_identical_to_eager_evaluation(gpu):
    """Deferred evaluation (one fmb_rv_eval_chain pass per chain) against one kernel per operation: same bits, fewer launches."""
    lazy, tl, nl = _random_algebra(gpu, True)
    eager, te, ne = _random_algebra(gpu, False)
    assert tl == te
    for a, b in zip(lazy, eager):
        assert same_bits(a, b)
    assert nl < 0.5 * ne, (nl, ne)


@pytest.mark.gpu
def test_deferred_operations_report_size_mismatch_at_once(gpu):
    a, b = gpu.RandomVariableCuda(0.0, np.ones(10)), gpu.RandomVariableCuda(0.0, np.ones(11))
    keep = gpu.native.lazy_min_n()
    for lazy in (True, False):
        gpu.native.set_lazy(lazy, min_n=0)
        try:
            with pytest.raises(ValueError):
                a.add(1.0).mult(b)
        finally:
            gpu.native.set_lazy(True, min_n=keep)


def test_sums_of_many_vectors_in_one_launch(gpu):
    """fmb_rv_reduce_many (the batched numeraire adjustment of the LIBOR market model): the double-double sums of up to 64 vectors from one
    launch equal the one-at-a-time reductions (the grids differ, so the merge order differs: a last-bit tolerance), for plain sums and for
    sum((1 / x) * a) = invert().mult(a) summed, also on deferred chains, and one launch is what it takes."""
    nv = gpu.native
    RV = gpu.RandomVariableCuda
    rng = np.random.default_rng(99)
    for n in (1, 255, 4097, 300_001):
        xs = [rng.uniform(0.5, 2.0, n) for _ in range(41)]
        vs = [RV(0.0, x) for x in xs]
        vs[3] = vs[3].mult(2.0).add(1.0)                     # a recorded chain among the operands
        xs[3] = xs[3] * 2.0 + 1.0
        before = nv.launch_count()
        hi, lo = nv.reduce_many(nv.RM_SUM, [v.dv for v in vs])
        assert nv.launch_count() - before <= 2               # (the chain, then ONE launch for all 41 sums)
        for v, h, l, x in zip(vs, hi, lo, xs):
            single = v.getAverage() * n
            assert abs((h + l) - single) <= 2e-16 * abs(single)
            assert abs((h + l) - float(np.sum(x.astype(np.longdouble)))) <= 1e-15 * abs(single)
        hi, lo = nv.reduce_many(nv.RM_SUM_INVERT_MULT, [v.dv for v in vs], 1.25)
        for v, h, l in zip(vs, hi, lo):
            single = v.invert().mult(1.25).getAverage() * n
            assert abs((h + l) - single) <= 2e-16 * abs(single)
    with pytest.raises(ValueError):
        nv.reduce_many(nv.RM_SUM, [RV(0.0, np.ones(3)).dv, RV(0.0, np.ones(4)).dv])
    with pytest.raises(ValueError):
        nv.reduce_many(nv.RM_SUM, [RV(0.0, np.ones(3)).dv] * 65)
