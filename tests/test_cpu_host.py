"""-m "not gpu": the C-ABI library loads and exports every symbol of include/finmath_b200.h, fails loudly without a GPU
(no CPU fallback), and the host-side logic (time grid, Scalar, dispatch of deterministic values, jump-ahead polynomial
arithmetic, sharding arithmetic, device exp/log compiled for the host) is correct."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "finmath_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(fmb_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 40
    lib = C.CDLL(pkg.native.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(pkg.native.PROTOTYPES.keys())          # the Python binding covers the whole header, nothing else


def test_no_cpu_fallback_without_a_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    nv = pkg.native
    with pytest.raises(nv.NoDeviceError):
        nv.init(0)
    with pytest.raises(nv.NoDeviceError):
        pkg.RandomVariableCuda(0.0, [1.0, 2.0, 3.0])
    td = pkg.TimeDiscretizationFromArray(0.0, 4, 0.25)
    with pytest.raises(nv.NoDeviceError):
        pkg.BrownianMotionCuda(td, 1, 100, 3141).getBrownianIncrement(0, 0)
    # deterministic values never touch the device (host scalars, like the reference's Scalar branch)
    assert pkg.RandomVariableCuda(0.0, 2.0).mult(3.0).add(pkg.Scalar(1.0)).doubleValue() == 7.0


def test_product_path_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under finmath-lib_b200/ may reference it."""
    pkg_dir = os.path.join(ROOT, "finmath-lib_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".java", ".c")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "orc_" not in text, os.path.join(dirpath, f)


def test_jump_ahead_polynomial_arithmetic_on_host(pkg, orc):
    """x^J mod phi applied as a GF(2) combination of shifted raw words == the sequentially advanced state."""
    lib = pkg.native.load()
    st = np.empty(624, dtype=np.uint32)
    w, g = C.c_int(), C.c_int()
    lib.fmb_test_host_jump.argtypes = [C.c_int64, C.c_uint64, pkg.native.c_u32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    assert lib.fmb_test_host_jump(3141, 0, st.ctypes.data_as(pkg.native.c_u32p), C.byref(w), C.byref(g)) == 0
    assert w.value == 135 and g.value >= 64                      # weight of the MT19937 characteristic polynomial
    for seed, J in [(3141, 1), (3141, 624), (-1, 12345), (53252, 240 * 13514), (3141, 2 * 40 * 3 * 500_000)]:
        assert lib.fmb_test_host_jump(seed, J, st.ctypes.data_as(pkg.native.c_u32p), None, None) == 0
        raw = list(int(v) for v in st)
        for k in range(624, 634):                               # ten outputs generated from the jumped state
            y = (raw[k - 624] & 0x80000000) | (raw[k - 623] & 0x7fffffff)
            raw.append(raw[k - 227] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0))
        out = []
        for y in raw[624:]:
            y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680; y ^= (y << 15) & 0xefc60000; y ^= y >> 18
            out.append(y & 0xffffffff)
        assert out == [int(v) for v in orc.mt_words(seed, J, 10)], (seed, J)


def test_time_discretization_mirror_equals_oracle(pkg, orc):
    for args in [(0.0, 100, 0.05), (0.0, 1000, 0.005), (0.0, 40, 0.5), (0.0, 200, 0.1)]:
        td = pkg.TimeDiscretizationFromArray(*args)
        assert np.array_equal(td.times, orc.time_discretization(*args))
        for t in (0.0, 0.3, 0.35, 5.0, 100.0, -1.0):
            assert td.getTimeIndex(t) == orc.time_index(td.times, t)
    td = pkg.TimeDiscretizationFromArray([0.5, 0.0, 0.5, 0.25])
    assert td.times.tolist() == [0.0, 0.25, 0.5] and td.getNumberOfTimeSteps() == 2 and td.getTimeStep(1) == 0.25


def test_scalar_and_deterministic_dispatch(pkg):
    S, G, CPU = pkg.Scalar, pkg.RandomVariableCuda, pkg.RandomVariableFromDoubleArray
    a = S(2.0)
    assert a.add(S(3.0)).doubleValue() == 5.0 and a.sub(S(3.0)).doubleValue() == -1.0 and a.div(S(4.0)).doubleValue() == 0.5
    assert a.discount(S(0.05), 0.5).doubleValue() == 1.0 / (0.05 * (0.5 / 2.0) + 1.0 / 2.0)       # Scalar.java:320-328
    assert a.accrue(S(0.05), 0.5).doubleValue() == 0.05 * (0.5 * 2.0) + 2.0
    assert S(0.0).discount(S(0.05), 0.5).doubleValue() == 0.0
    assert a.addProduct(S(3.0), 4.0).doubleValue() == 14.0 and a.choose(S(1.0), S(-1.0)).doubleValue() == 1.0
    assert a.getFiltrationTime() == float("-inf") and a.getTypePriority() == 0
    d = G(1.5, 2.0)
    r = d.add(S(1.0))                                                            # deterministic GPU value: host arithmetic, keeps its time
    assert isinstance(r, G) and r.isDeterministic() and r.doubleValue() == 3.0 and r.getFiltrationTime() == 1.5
    assert d.addProduct(G(2.5, 3.0), G(0.5, 4.0)).getFiltrationTime() == 2.5 and d.addProduct(S(3.0), 4.0).doubleValue() == 14.0
    assert d.accrue(S(0.1), 0.5).doubleValue() == 2.0 * (1.0 + 0.1 * 0.5) and d.discount(S(0.1), 0.5).doubleValue() == 2.0 / (1.0 + 0.1 * 0.5)
    assert CPU(1.0).mult(d).doubleValue() == 2.0 and isinstance(CPU(1.0).mult(d), G)         # priority 2 wins over the CPU type
    assert d.getAverage() == 2.0 and d.getVariance() == 0.0 and d.getStandardError() == 0.0 and d.getQuantile(0.3) == 2.0
    assert d.pow(0.5).doubleValue() == 2.0 ** 0.5 and d.cap(1.0).doubleValue() == 1.0 and d.floor(3.0).doubleValue() == 3.0
    assert np.isnan(G(0.0, float("nan")).cap(1.0).doubleValue())
    assert d.getHistogram([1.0, 3.0]).tolist() == [1.0, 0.0, 1.0]                        # deterministic branch :505-517
    assert d.invert().doubleValue() == 0.5 and G(0.0, 0.0).invert().doubleValue() == float("inf")


def test_models_host_tables(pkg):
    from common import lmm_setup
    s = lmm_setup(pkg)
    fm = s["factor_matrix"]
    assert fm.shape == (40, 3)
    assert np.allclose(np.sum(fm * fm, axis=1), 1.0, atol=1e-12)             # rows re-normalised by the factor reduction
    assert np.all(fm[0] * np.array([1, 1, 1]) != 0) and fm[0, 0] > 0          # sign convention: first entry positive
    fl, var = s["cov"].getFactorLoadingTable()
    assert fl.shape == (40, 40, 3) and np.all(fl[5, :6] == 0.0) and np.all(fl[5, 6:] != 0.0)   # fixed rates have no volatility
    assert var[3, 10] == s["sigma"][3, 10] * s["sigma"][3, 10] * 1.0
    model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], None, s["cov"], None, {"measure": "SPOT"})
    assert [model._first(t) for t in (0.0, 0.5, 0.75, 19.5)] == [1, 2, 2, 40]


def test_dd_merge_and_local_shard(pkg):
    from finmath_lib_b200.sharding import dd_merge, ShardContext
    h, l = dd_merge(1.0, 1e-20, 1e-16, 3e-33)
    assert h == 1.0000000000000002 or h == 1.0
    assert abs((h + l) - (1.0 + 1e-16)) < 1e-30
    sc = ShardContext(1, 4)
    assert sc.local_range(10) == (2, 5) and [ShardContext(r, 4).local_range(10) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert pkg.LOCAL.sum_dd(1.0, 2.0) == (1.0, 2.0) and pkg.LOCAL.global_count(7) == 7


def test_device_exp_log_accuracy_on_host(tmp_path):
    """fmb_math.cuh (the exp / log of every kernel) compiled for the host and measured against mpmath: < 1 ulp."""
    import mpmath as mp
    so = tmp_path / "libshim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-o", str(so),
                           os.path.join(ROOT, "tests", "host_math_shim.cpp")])
    L = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)

    def run(fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        getattr(L, fn)(x.ctypes.data_as(dp), y.ctypes.data_as(dp), C.c_long(x.size))
        return y

    mp.mp.dps = 40
    rng = np.random.default_rng(1)

    def max_ulp(y, x, f):
        worst = 0.0
        for yi, xi in zip(y, x):
            t = f(mp.mpf(float(xi)))
            worst = max(worst, float(abs(mp.mpf(float(yi)) - t) / mp.mpf(float(np.spacing(abs(float(t)))))))
        return worst

    xe = np.concatenate([rng.uniform(-12, 12, 4000), rng.uniform(-700, 700, 1000), rng.uniform(-1e-3, 1e-3, 500)])
    assert max_ulp(run("shim_exp", xe), xe, mp.exp) < 1.0
    assert np.array_equal(run("shim_exp", xe), run("shim_exp2", xe))
    xl = np.concatenate([np.exp(rng.uniform(-12, 12, 4000)), rng.uniform(0.5, 2.0, 2000), np.exp(rng.uniform(-700, 700, 1000))])
    assert max_ulp(run("shim_log", xl), xl, mp.log) < 0.8
    assert np.array_equal(run("shim_log", xl), run("shim_log2", xl))
    sp = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1.0])
    got = run("shim_log", sp)
    assert got[0] == -np.inf and np.isnan(got[1]) and got[2] == np.inf and np.isnan(got[3]) and abs(got[4] + 744.4400719213812) < 1e-12 and got[5] == 0.0
    se = np.array([0.0, 709.79, -745.2, -746.0, -720.0, np.inf, -np.inf, np.nan])
    ge = run("shim_exp", se)
    assert ge[0] == 1.0 and ge[1] == np.inf and ge[3] == 0.0 and abs(ge[4] / 2.03223080e-313 - 1) < 1e-6 and ge[5] == np.inf and ge[6] == 0.0 and np.isnan(ge[7])


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["FMB_ROOT"])
import numpy as np
import torch.distributed as dist
import __graft_entry__ as graft
pkg = graft.load_package()
shard = pkg.from_environment(backend="gloo")
assert shard.world == 2
n = 1001
lo, hi = shard.local_range(n)
x = np.linspace(-1.0, 3.0, n) ** 3
# what RandomVariableCuda.getAverage does across shards: per-rank double-double partial -> all_gather -> merge in rank order
local_hi = float(np.sum(x[lo:hi])); local_lo = 0.0
H, L = shard.sum_dd(local_hi, local_lo)
tot = shard._all_gather([local_hi])[:, 0]
assert H + L == tot[0] + tot[1]
assert shard.global_count(hi - lo) == n
Hm, Lm = shard.sum_dd_many(np.array([1.0 + shard.rank, 2.0]), np.array([0.0, 1e-20]))
assert Hm.tolist() == [3.0, 4.0]
assert shard.min(float(shard.rank)) == 0.0 and shard.max(float(shard.rank)) == 1.0
# a shard that owns no element does not take part (its local NaN means "no data", not NaN)
assert shard.min(5.0 if shard.rank == 0 else float('nan'), shard.rank == 0) == 5.0 and shard.max(float('nan'), False) != shard.max(float('nan'), False)
assert np.isnan(shard.max(float('nan') if shard.rank == 1 else 1.0))
g = shard.gather(x[lo:hi])
assert np.array_equal(g, x)
# every rank holds identical bits
allH = shard._all_gather([H, L])
assert allH[0].tolist() == allH[1].tolist()
bm = pkg.BrownianMotionCuda(pkg.TimeDiscretizationFromArray(0.0, 4, 0.5), 3, 1001, 3141, shard=shard)
assert bm.getPathRange() == (lo, hi) and bm.getNumberOfLocalPaths() == hi - lo
dist.barrier()
dist.destroy_process_group()
print("rank", shard.rank, "ok")
'''


@pytest.mark.parametrize("tiny", ["shm", "torch"])
def test_sharding_collectives_world_size_2_gloo(tmp_path, tiny):
    """N > 1 host logic on CPU: two processes, gloo backend, 127.0.0.1 rendezvous; the reductions' partials travel through the
    shared-memory mailbox (co-located ranks) or through torch.distributed."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.replace("assert shard.world == 2", "assert shard.world == 2 and (shard._mailbox is not None) == (os.environ['FMB_TINY_COLLECTIVES'] == 'shm')"))
    env = dict(os.environ, FMB_ROOT=ROOT, FMB_TINY_COLLECTIVES=tiny)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29517" if tiny == "shm" else "29518", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2, out.stdout               # both ranks finished every assertion (their prints may interleave)


def test_hull_white_coefficient_tables_equal_oracle(pkg, orc):
    """Host-side Hull-White closed forms (Scalar arithmetic in the reference's order) == oracle, bit for bit."""
    class P:                                                       # minimal process stand-in: only the time grid is consulted
        def __init__(self, td): self.td = td
        def getTime(self, i): return self.td.getTime(i)
        def getTimeDiscretization(self): return self.td
        def getScheme(self): return 0
    for (nsteps, dt, vt, vol, mr) in [(40, 0.5, [0.0], [0.005], [0.1]),
                                      (200, 0.1, list(np.arange(0, 21.0)), list(0.005 + 0.0005 * np.floor(np.arange(0, 21.0)) / 20), [0.1] * 21),
                                      (30, 0.25, [0.0, 1.0, 2.5, 4.0], [0.01, 0.012, 0.008, 0.02], [0.05, 0.1, 0.2, 0.15])]:
        td = pkg.TimeDiscretizationFromArray(0.0, nsteps, dt)
        vm = pkg.ShortRateVolatilityModelAsGiven(pkg.TimeDiscretizationFromArray(vt), vol, mr)
        model = pkg.HullWhiteModel(None, td, vm)
        spec = model.getFusedSpecification(P(td))
        _, coef = orc.hull_white_process(3141, td.times, 4, vm.getTimeDiscretization().times, vol, mr, 0)
        mine = np.column_stack([spec["drift0"], spec["drift1"], spec["factorLoadings"]])
        assert np.array_equal(mine, coef)


_KEEP_ALIVE = []


class _FakeLib:
    """Stands in for libfinmath_b200.so under native.LazyVector: records the C-ABI calls a chain turns into (no device needed).  The chain
    logic lives in the C accelerator, which calls the entry points by ADDRESS: the fakes are handed over as ctypes callbacks."""

    def __init__(self):
        self.calls, self.next_handle = [], 100
        H, D, I = C.c_uint64, C.c_double, C.c_int
        self.callbacks = [
            C.CFUNCTYPE(I, I, H, D, C.POINTER(H))(self.fmb_rv_unary),
            C.CFUNCTYPE(I, I, H, D, H, D, C.POINTER(H))(self.fmb_rv_binary),
            C.CFUNCTYPE(I, I, H, D, H, D, H, D, D, C.POINTER(H))(self.fmb_rv_ternary),
            C.CFUNCTYPE(I, H)(self.fmb_rv_free),
            C.CFUNCTYPE(I, I, C.POINTER(C.c_ubyte), I, C.POINTER(H), I, C.POINTER(D), I, C.POINTER(H))(self.fmb_rv_eval_chain)]

    def addresses(self):
        return [C.cast(cb, C.c_void_p).value for cb in self.callbacks]

    def _out(self, ref):
        self.next_handle += 1
        ref[0] = self.next_handle
        return 0

    def fmb_rv_unary(self, op, x, a, out):
        self.calls.append(("unary", op, x, a))
        return self._out(out)

    def fmb_rv_binary(self, op, x, sx, y, sy, out):
        self.calls.append(("binary", op, x, sx, y, sy))
        return self._out(out)

    def fmb_rv_ternary(self, op, x, sx, y, sy, z, sz, a, out):
        self.calls.append(("ternary", op, x, sx, y, sy, z, sz, a))
        return self._out(out)

    def fmb_rv_eval_chain(self, n, code, start, leaves, nl, scalars, ns, out):
        self.calls.append(("chain", n, bytes(code[i] for i in range(8 * n)), start, [int(leaves[i]) for i in range(nl)], [float(scalars[i]) for i in range(ns)]))
        return self._out(out)

    def fmb_rv_free(self, h):
        return 0


def test_deferred_arithmetic_builds_the_documented_chain_encoding(pkg, monkeypatch):
    """native.LazyVector (host logic of fmb_rv_eval_chain): chains grow while a result has one consumer, a second consumer or a non
    element-wise consumer evaluates them, one-operation chains use the specialised kernel, scalars that must not be merged are not."""
    nv = pkg.native
    fake = _FakeLib()
    _KEEP_ALIVE.append(fake)                                 # (objects of this test release their fake handles through these callbacks later)
    monkeypatch.setattr(nv, "_lib", fake)
    monkeypatch.setattr(nv, "_lazy", True)
    monkeypatch.setattr(nv, "_lazy_min_n", 0)
    nv._F.bind(*fake.addresses(), nv.check, pkg.RandomVariableCuda, nv.DeviceVector, nv.LazyVector)
    nv._F.set_lazy_min_n(0)
    x, y, z = nv.DeviceVector(11, 8), nv.DeviceVector(12, 8), nv.DeviceVector(13, 8)
    try:
        # (x - 0.03) * 0.5 / y, then  z + that * y  (the pending chain sits in the second operand position of the ternary)
        a = nv.binary(nv.B_DIV, nv.unary(nv.U_MULT, nv.unary(nv.U_SUB, x, 0.03), 0.5), 0.0, y, 0.0)
        b = nv.ternary(nv.T_ADD_PRODUCT, z, 0.0, a, 0.0, y, 0.0)
        assert fake.calls == [] and b.pending() and b.n == 8
        h = b.h
        assert h == 101 and not b.pending() and len(fake.calls) == 1
        kind, n, code, start, leaves, scalars = fake.calls[0]
        assert (kind, n, start, leaves, scalars) == ("chain", 4, 0, [11, 12, 13], [0.03, 0.5, 0.0])
        assert code == bytes((0, nv.U_SUB, 0, 128, 0, 0, 0, 0)) + bytes((0, nv.U_MULT, 0, 129, 0, 0, 0, 0)) \
            + bytes((1, nv.B_DIV, 0, 1, 0, 0, 0, 0)) + bytes((2, nv.T_ADD_PRODUCT, 1, 2, 1, 130, 0, 0))
        # one operation: the specialised kernel, with the scalar broadcast in its operand position
        fake.calls.clear()
        assert nv.binary(nv.B_SUB, None, 2.0, x, 0.0).h == 102
        assert fake.calls == [("binary", nv.B_SUB, 0, 2.0, 11, 0.0)]
        # a second consumer evaluates the shared prefix once and uses it as a leaf
        fake.calls.clear()
        p = nv.unary(nv.U_EXP, nv.unary(nv.U_SQUARED, x))
        q1, q2 = nv.unary(nv.U_ADD, p, 1.0), nv.unary(nv.U_ADD, p, 2.0)
        assert [c[0] for c in fake.calls] == ["chain"] and fake.calls[0][1] == 2          # p itself (squared, exp), when q2 was built
        q1.h, q2.h
        assert [c[0] for c in fake.calls] == ["chain", "chain", "unary"]                  # q1 = x..+1 in one pass, q2 = p + 2
        assert fake.calls[1][1] == 3 and fake.calls[2][2] == p.h
        # +0.0 / -0.0 and NaN scalars keep their own slots
        fake.calls.clear()
        nv.unary(nv.U_MULT, nv.unary(nv.U_ADD, nv.unary(nv.U_ADD, x, 0.0), -0.0), float("nan")).h
        sc = fake.calls[0][5]
        assert len(sc) == 3 and np.signbit(sc[1]) and not np.signbit(sc[0]) and np.isnan(sc[2])
        # operands of different lengths are rejected when the operation is issued
        with pytest.raises(ValueError):
            nv.binary(nv.B_ADD, x, 0.0, nv.DeviceVector(14, 9), 0.0)
        # a full chain is evaluated and continued from its result
        fake.calls.clear()
        c = x
        for _ in range(nv.CHAIN_MAX_INSTR + 3):
            c = nv.unary(nv.U_ADD, c, 1.0)
        c.h
        assert [cc[1] for cc in fake.calls if cc[0] == "chain"] == [nv.CHAIN_MAX_INSTR, 3]
        # the same through the RandomVariable fast paths: (X * 2 + Y).exp() is one chain, evaluated when a handle is asked for
        fake.calls.clear()
        RV = pkg.RandomVariableCuda
        X, Y = RV(0.5, None, _dv=x, _n=8), RV(1.5, None, _dv=y, _n=8)
        r = X.mult(2.0).add(Y).exp()
        assert fake.calls == [] and type(r.dv) is nv.LazyVector and r.dv.chain_length() == 3 and r.getFiltrationTime() == 1.5
        r.dv.h
        assert [c[0] for c in fake.calls] == ["chain"] and fake.calls[0][4] == [11, 12]
        # short vectors are launched at once
        nv._F.set_lazy_min_n(9)
        fake.calls.clear()
        e = X.mult(2.0)
        assert type(e.dv) is nv.DeviceVector and [c[0] for c in fake.calls] == ["unary"]
        e.dv.h = 0
    finally:
        for v in (x, y, z):
            v.h = 0                                            # fake handles: nothing to free
        monkeypatch.undo()
        nv._F.set_lazy_min_n(2 ** 64 - 1)
        nv._lib = None if not hasattr(nv._lib, "fmb_init") else nv._lib
        if nv._lib is not None:
            nv._bind_fast()                                    # the real library again for the tests that follow


def test_reference_arm_inputs_are_independent_of_and_equal_to_the_products_tables(pkg, orc):
    """bench.py --impl reference builds its market data with numpy + the oracle only (reference_inputs); the product's host classes build
    the same tables for the device arm.  Two independent constructions of the volatility table and of the 3-factor reduction: they must
    agree (factor matrix to 1e-12: both call LAPACK's symmetric eigen-solver, the sign / scaling / second-pass conventions are restated
    twice)."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    from common import lmm_setup
    r = bench.reference_inputs(orc)
    s = lmm_setup(pkg)
    assert np.array_equal(r["sim"], s["sim"].times) and np.array_equal(r["tenor"], s["tenor"].times)
    assert np.array_equal(r["sigma"], s["sigma"])
    assert np.max(np.abs(r["factor_matrix"] - s["factor_matrix"])) < 1e-12
    assert np.array_equal(r["L0"], s["L0"])


def test_bermudan_host_logic_over_a_handle_only_stub(tmp_path):
    """The host mirror of a whole C5 valuation (fused LMM simulation, 20 exercise dates, numeraire batch, regression calls) driven against
    profiles/tools/null_abi.c - a stand-in for the C ABI that only hands out handles - in a separate process: every call the mirror makes
    exists in the header with the argument types the binding declares, and the number of native launches per valuation stays at the
    figure the design documents (255 on the GPU; the stub does not alias frozen rates, a few more here)."""
    lib = tmp_path / "libnull_abi.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "profiles", "tools", "null_abi.c"),
                           "-o", str(lib)])
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "tools", "host_profile.py"), "--lib", str(lib), "--reps", "2", "--paths", "5000"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"(\d+) native launches per valuation", out.stdout)
    assert m, out.stdout
    assert 200 <= int(m.group(1)) <= 270, out.stdout


def test_volatility_table_equals_the_elementwise_definition_bit_for_bit(pkg):
    """getVolatilityTable evaluates exp once per distinct time to maturity; the table must equal getVolatility(t, j) entry by entry (the
    fused kernel is fed from the table, the generic loop from the method), and the forward-implied discount factors kept per model must
    be the running product the discounting adjustment of a swaption expects."""
    import math
    td = pkg.TimeDiscretizationFromArray(0.0, 12, 0.25)
    tenor = pkg.TimeDiscretizationFromArray(0.0, 7, 0.5)
    vol = pkg.LIBORVolatilityModelFourParameterExponentialForm(td, tenor, 0.21, 0.013, 0.37, 0.11, True)
    table = vol.getVolatilityTable()
    assert table.shape == (12, 7)
    for t in range(12):
        for j in range(7):
            v = vol.getVolatility(t, j)
            v = v.doubleValue() if hasattr(v, "doubleValue") else float(v)
            assert table[t, j] == v, (t, j, table[t, j], v)
    corr = pkg.LIBORCorrelationModelExponentialDecay(td, tenor, 2, 0.1, True)
    cov = pkg.LIBORCovarianceModelFromVolatilityAndCorrelation(td, tenor, vol, corr)
    L0 = [0.02 + 0.003 * i for i in range(7)]
    model = pkg.LIBORMarketModelFromCovarianceModel.of(tenor, None, L0, None, None, cov, None, {"measure": "SPOT"})
    df = model.getDiscountFactorsFromForwardCurve()
    assert df is model.getDiscountFactorsFromForwardCurve() and len(df) == 8 and df[0] == 1.0
    expect = 1.0
    for i in range(7):
        expect = expect / (1.0 + L0[i] * 0.5)
        assert df[i + 1] == expect
    assert math.isclose(df[-1], math.prod(1.0 / (1.0 + l * 0.5) for l in L0), rel_tol=1e-14)


def test_randomvariable_mirror_offers_every_method_of_the_reference_interface(pkg):
    """Every method name of net.finmath.stochastic.RandomVariable (fixture tests/golden/randomvariable_interface.json, written by
    tests/golden/make_interface_list.py from the reference checkout) exists on the device type, on Scalar and on the AAD type, so code
    written against the interface does not meet an AttributeError half way through a valuation."""
    import json
    names = json.load(open(os.path.join(ROOT, "tests", "golden", "randomvariable_interface.json")))["methods"]
    assert len(names) >= 50
    for cls in (pkg.RandomVariableCuda, pkg.Scalar, pkg.RandomVariableDifferentiableAAD):
        assert [n for n in names if not hasattr(cls, n)] == [], cls
    s = pkg.Scalar(2.0)
    assert s.apply(lambda x: x * x).doubleValue() == 4.0 and s.apply(lambda x, y: x + y, pkg.Scalar(1.0)) is None     # Scalar.java:183-197
    assert s.getOperator() is None and s.getRealizationsStream() is None                                          # Scalar.java:88-95
    with pytest.raises(NotImplementedError):
        s.getHistogram([0.0, 1.0])
    assert s.appy(lambda r: r.add(1.0)).doubleValue() == 3.0                                                      # RandomVariable.java:316


def test_reference_arm_of_the_bench_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU (it times the CPU restatement on the host cores): one JSON line with the keys the driver
    reads, the same metric / unit as the product arm, zero copy bytes, and only the oracle's library loaded."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "LMM forward-rate path-steps/sec" and line["unit"] == "path-steps/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0 and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
