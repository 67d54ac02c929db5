"""-m "not gpu": pins the oracle (the checker of the GPU tests) against the committed golden fixtures and against the
reference tests' own exact identities / closed-form tolerances.  Fixtures: tests/golden/mt_as241.json (make_golden.py)."""
import json
import os

import numpy as np
import pytest

from common import rel_err

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mt_as241.json")))


def test_mt19937_public_known_answer(orc):
    assert [int(v) for v in orc.mt_words_key(GOLDEN["mt_kat_key"], 8)] == GOLDEN["mt_kat_words"]


@pytest.mark.parametrize("seed", sorted(GOLDEN["seeds"].keys(), key=int))
def test_mt19937_seeding_and_doubles(orc, seed):
    g = GOLDEN["seeds"][seed]
    s = int(seed)
    assert [int(v) for v in orc.mt_words(s, 0, 8)] == g["words_0_8"]
    assert [int(v) for v in orc.mt_words(s, 1990, 10)] == g["words_1990_2000"]          # across three state refreshes
    assert list(orc.mt_uniforms(s, 0, 4)) == g["uniforms_0_4"]                            # 26+26 bit doubles, bit-exact


def test_mt_skip_equals_sequential(orc):
    full = orc.mt_words(3141, 0, 5000)
    for off in (1, 623, 624, 625, 1247, 4000):
        assert np.array_equal(orc.mt_words(3141, off, 100), full[off:off + 100])
    raw = orc.mt_raw_sequence(3141, 2000)
    # tempering of the raw recurrence reproduces the output stream (output i = temper(raw[624 + i]))
    y = raw[624:1624].astype(np.uint64)
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680; y ^= (y << 15) & 0xefc60000; y ^= y >> 18
    assert np.array_equal((y & 0xffffffff).astype(np.uint32), full[:1000])


def test_as241_against_independent_values(orc):
    p = np.array([e["p"] for e in GOLDEN["icdf"]])
    z = orc.icdf(p)
    for e, v in zip(GOLDEN["icdf"], z):
        assert abs(v - e["z"]) <= 2e-15 * max(1.0, abs(e["z"])), e
    assert orc.icdf([0.0])[0] == 0.0 and orc.icdf([1.0])[0] == 0.0       # quirk NormalDistribution.java:141-143
    assert orc.icdf([0.5])[0] == 0.0
    # symmetry and monotonicity on a grid
    g = np.linspace(1e-6, 1 - 1e-6, 20001)
    zg = orc.icdf(g)
    assert np.all(np.diff(zg) > 0) and rel_err(zg, -zg[::-1], scale=1e-3) < 1e-9


def test_time_discretization_tick_rounding(orc):
    for e in GOLDEN["tick_rounding"]:
        t = orc.time_discretization_from_array([0.0, e["t"]])
        assert t[-1] == e["rounded"], e
    td = orc.time_discretization(0.0, 10, 0.1)
    assert td.size == 11 and td[1] == 0.09999999999999999
    assert orc.time_index(td, 0.3) == 3 and orc.time_index(td, 0.35) == -5 and orc.time_index(td, 5.0) == -12
    fine = orc.time_discretization(0.0, 1000, 0.005)
    steps = np.unique(np.round(np.diff(fine) * 8760))
    assert set(steps) <= {43.0, 44.0}                                      # 0.005y = 43.8 ticks -> a 44/44/.../43 pattern
    assert orc.time_discretization_from_array([0.5, 0.0, 0.5, 0.25]).tolist() == [0.0, 0.25, 0.5]   # sorted, de-duplicated


def test_randomvariable_reference_identities(orc):
    """T/montecarlo/RandomVariableTest.java:53-142 on the oracle's element-wise semantics."""
    x = np.array([3.0, 1.0, 0.0, 2.0, 4.0, 1.0 / 3.0])
    assert np.array_equal(orc.rv_unary(1, x), orc.rv_unary(18, x, 0.5))     # sqrt == pow(0.5), tolerance 0.0
    assert np.array_equal(orc.rv_unary(0, x), orc.rv_unary(18, x, 2.0))     # squared == pow(2.0)
    s = np.array([-4.0, -2.0, 0.0, 2.0, 4.0])
    s = orc.rv_unary(14, orc.rv_unary(10, s, 4.0), 2.0)
    assert orc.rv_reduce(0, s) == 2.0 and orc.rv_reduce(2, s) == 2.0
    assert orc.rv_reduce(7, x) == np.sqrt(orc.rv_reduce(2, x))
    # Math.min / Math.max semantics and choose
    m = orc.rv_unary(16, np.array([np.nan, -0.0, 0.0, 1.0]), 0.0)
    assert np.isnan(m[0]) and np.signbit(m[1]) and not np.signbit(m[2]) and m[3] == 0.0
    c = orc.rv_ternary(6, np.array([np.nan, -0.0, 0.0, -1.0]), np.ones(4), np.full(4, 2.0))
    assert c.tolist() == [2.0, 1.0, 1.0, 2.0]
    # Kahan mean is the correctly rounded mean here where the naive sum is not
    v = np.array([1.0] + [1e-16] * 1000000)
    assert orc.rv_reduce(0, v) == (1.0 + 1e-10) / 1000001
    # quantile index round((n+1)q - 1) clamped (:454-459)
    q = np.arange(10.0)
    assert [orc.rv_reduce(9, q, a=a) for a in (0.0, 0.1, 0.5, 0.95, 1.0)] == [0.0, 0.0, 5.0, 9.0, 9.0]
    assert np.allclose(orc.rv_histogram(q, np.array([2.0, 5.5])), [0.3, 0.3, 0.4])


def test_black_scholes_price_like_reference_test(orc):
    """T/montecarlo/assetderivativevaluation/MonteCarloBlackScholesModelTest.java:80 — MC value within 0.005 of the analytic value."""
    g = GOLDEN["black_scholes_call"]
    td = orc.time_discretization(0.0, 100, 0.05)
    price, proc, vals = orc.bs_european(3141, td, 100000, g["S0"], g["r"], g["sigma"], 2, g["T"], g["K"])
    assert abs(price - g["value"]) < 0.005
    assert proc.shape == (101, 1, 100000) and np.all(proc[0] == 1.0)
    # Euler vs functional Euler differ only by rounding (log(exp(y)) vs y)
    price_e, proc_e, _ = orc.bs_european(3141, td, 2000, g["S0"], g["r"], g["sigma"], 0, g["T"], g["K"])
    price_f, proc_f, _ = orc.bs_european(3141, td, 2000, g["S0"], g["r"], g["sigma"], 2, g["T"], g["K"])
    assert rel_err(proc_e, proc_f) < 1e-13 and not np.array_equal(proc_e, proc_f)


def test_heston_xi_zero_equals_black_scholes(orc):
    """T/montecarlo/assetderivativevaluation/HestonModelTest.java:143-145 (1e-10 on the same Brownian driver)."""
    td = orc.time_discretization(0.0, 50, 0.1)
    ph, _, _ = orc.heston_european(3141, td, 20000, 1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.0, 0.1, 1, 2, 5.0, 1.25)
    # same driver: BS with 2 factors is not exposed, but factor 0 of a 2-factor driver differs from a 1-factor driver's draws;
    # compare through the closed form of the variance path instead: xi = 0 keeps V = sigma^2 exactly
    _, proc, _ = orc.heston_european(3141, td, 2000, 1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.0, 0.1, 1, 2, 5.0, 1.25)
    assert rel_err(proc[:, 1], np.full_like(proc[:, 1], 0.09)) < 1e-15
    assert 0.25 < ph < 0.40


def test_lmm_oracle_against_reference_test_tolerances(orc, pkg):
    """LIBORMarketModelValuationTest.java: bond (:199, 5e-4 ... 1e-2 range of tolerances) and swaption vs analytic approximation."""
    from common import lmm_setup, lmm_oracle
    s = lmm_setup(pkg)
    ref = lmm_oracle(orc, s, 20000, scheme=1)
    # zero bond P(5y;0) = E[1/N(5)] against the curve value (testBond)
    n = ref.numeraire(5.0)
    assert abs(np.mean(1.0 / n) - s["df"][10]) < 5e-4
    # par swaption 5y into 5y: positive, and at-the-money value consistent with the Black approximation to 1e-2 (testSwaption tolerance)
    fixing = [5.0 + 0.5 * i for i in range(10)]
    payment = [5.5 + 0.5 * i for i in range(10)]
    price, vals, se = ref.swaption(5.0, fixing, payment, [0.05] * 10)
    annuity = sum(0.5 * s["df"][11 + i] for i in range(10))
    swaprate = (s["df"][10] - s["df"][20]) / annuity
    from scipy.stats import norm
    # the implied Black volatility of the MC price must be a sensible swap-rate volatility for sigma_j in [0.3, 0.5], rho < 1
    def black(vol):
        d1 = (np.log(swaprate / 0.05) + 0.5 * vol * vol * 5) / (vol * np.sqrt(5))
        return annuity * (swaprate * norm.cdf(d1) - 0.05 * norm.cdf(d1 - vol * np.sqrt(5)))
    assert black(0.25) < price < black(0.50)
    assert se < 2e-3
    # spot-measure frozen rates: L_j stops moving at its fixing
    proc = ref.process()
    assert np.array_equal(proc[7, 3], proc[3, 3]) and not np.array_equal(proc[3, 3], proc[2, 3])
    assert rel_err(proc[0], np.full_like(proc[0], 0.05)) < 3e-16            # exp(log(0.05)): the reference also goes through log/exp at t = 0


def test_fused_cpu_port_equals_reference_shaped_oracle(orc, pkg):
    """The best-effort CPU baseline (fused, path-parallel) is the same arithmetic as the reference-shaped RV-op oracle: bit-identical."""
    from common import lmm_setup, lmm_oracle
    s = lmm_setup(pkg, n_libors=12, n_factors=3)
    for scheme in (0, 1, 2, 3):
        ref = lmm_oracle(orc, s, 500, scheme=scheme, with_discount_curve=False).process()
        _, fused = orc.time_lmm_fused(3141, s["sim"].times, s["tenor"].times, 3, 500, s["L0"], s["sigma"], s["factor_matrix"], scheme, 3, want_process=True)
        assert np.array_equal(ref, fused), scheme


def test_regression_solver_against_numpy(orc):
    rng = np.random.default_rng(5)
    X = rng.standard_normal((1000, 5))
    A, b = X.T @ X / 1000, X.T @ rng.standard_normal(1000) / 1000
    x, cond = orc.solve_pinv(A, b)
    assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-11) and abs(cond - np.linalg.cond(A)) < 1e-8 * cond
    A2 = np.array([[1.0, 1.0], [1.0, 1.0]])
    x2, _ = orc.solve_pinv(A2, np.array([2.0, 2.0]))
    assert np.allclose(x2, [1.0, 1.0])                                      # minimum-norm solution of a singular system


def test_bermudan_replay_with_given_coefficients_reproduces_the_valuation(orc, pkg):
    """The window check of the 8 M-path Bermudan (tests/test_gpu_fullsize.py) hands the device's regression coefficients to the oracle:
    replaying the oracle's own coefficients (and weight) must give back its own per-path values and exercise times exactly, also on a
    window started at a path offset with the coefficients of the full run."""
    from common import lmm_setup, lmm_oracle, bermudan_spec
    s = lmm_setup(pkg)
    b = bermudan_spec(s)
    args = (b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
    paths = 3000
    full = lmm_oracle(orc, s, paths, with_discount_curve=False)
    r = full.bermudan(*args)
    vals, ext = full.bermudan_given(*args, r["regression"], 1.0 / paths)
    assert np.array_equal(vals, r["values"]) and np.array_equal(ext, r["exercise_time"])
    window = lmm_oracle(orc, s, 500, with_discount_curve=False, path_offset=1700)
    vals_w, ext_w = window.bermudan_given(*args, r["regression"], 1.0 / paths)
    assert np.array_equal(vals_w, r["values"][1700:2200]) and np.array_equal(ext_w, r["exercise_time"][1700:2200])
    basis = window.bermudan_basis(b["fixing"][3], b["fixing"], b["payment"])
    assert basis.shape == (6, 500) and np.all(basis[0] == 1.0) and np.all(basis[2] == basis[1] * basis[1])


def test_two_math_libraries_separate_heston_paths_at_the_variance_kink(orc, pkg):
    """The same restatement with libm's exp / log and with the device kernels' exp / log (both on the CPU, both < 1 ulp): under full
    truncation with the Feller condition violated (xi = 0.5) a small fraction of paths touches V ~ 0, where sqrt amplifies a last-bit
    difference - beyond 1e-12 for ~1e-4 of the paths, while the price agrees to 1e-13.  This is why the GPU parity test checks the 1e-12
    path gate against the oracle in the device-math mode and bounds the deviation from the libm oracle by this libm-vs-libm figure.  With
    xi = 0 (no kink) the two agree within 1e-12 on every path."""
    td = pkg.TimeDiscretizationFromArray(0.0, 100, 0.05)
    paths = 20_000
    for xi in (0.5, 0.0):
        args = (31415, td.times, paths, 1.0, 0.05, 0.3, 0.05, 0.09, 0.1, xi, 0.1, 1, 2, 5.0, 1.10)
        p_libm, proc_libm, _ = orc.heston_european(*args)
        orc.set_math(1)
        try:
            p_dev, proc_dev, _ = orc.heston_european(*args)
        finally:
            orc.set_math(0)
        e0 = np.abs(proc_libm[:, 0] - proc_dev[:, 0]) / np.maximum(np.abs(proc_libm[:, 0]), 1.0)
        e1 = np.abs(proc_libm[:, 1] - proc_dev[:, 1]) / np.maximum(np.abs(proc_libm[:, 1]), 0.09)
        bad = np.mean(((e0 > 1e-12) | (e1 > 1e-12)).any(axis=0))
        assert abs(p_libm - p_dev) <= 1e-12 * abs(p_libm)
        if xi > 0:
            assert 0 < bad < 2e-3 and max(e0.max(), e1.max()) < 1e-9
        else:
            assert bad == 0.0


def _ref_fixture(name):
    import json
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


def _unhex(v):
    import struct
    if isinstance(v, list):
        return np.array([_unhex(x) for x in v])
    return struct.unpack(">d", bytes.fromhex(v.rjust(16, "0")))[0]


def test_reference_held_fixtures(orc, pkg):
    """Pins the oracle against numbers computed by finmath-lib ITSELF (tests/golden/ref_*.json, written by tests/golden/GenerateGolden.java
    on a machine with a JDK).  They cannot be produced in this image (no JVM), so until somebody commits them this test is skipped and the
    oracle stays 'parity unpinned' against the reference (pinned against independent sources only: mt_as241.json)."""
    mt, bm, bs, lmm = (_ref_fixture(n) for n in ("ref_mt.json", "ref_brownian.json", "ref_bs.json", "ref_lmm.json"))
    if mt is None and bm is None and bs is None and lmm is None:
        pytest.skip("no reference-held fixtures (tests/golden/ref_*.json): run tests/golden/GenerateGolden.java with a JDK - parity unpinned against the reference")
    from common import rel_err, lmm_setup, lmm_oracle, bermudan_spec
    if mt is not None:
        for seed, u in mt["uniforms"].items():
            assert np.array_equal(orc.mt_uniforms(int(seed), 0, 64), _unhex(u)), seed                     # bit-exact
        z = _unhex(mt["icdf_of_seed_3141"])
        got = orc.icdf(orc.mt_uniforms(3141, 0, 64))
        assert rel_err(got, z) < 2e-15
        central = np.abs(orc.mt_uniforms(3141, 0, 64) - 0.5) <= 0.425
        assert np.array_equal(got[central], z[central])                                                   # no libm call in the central branch
    if bm is not None:
        td = pkg.TimeDiscretizationFromArray(0.0, 4, 0.5)
        ref = np.array([_unhex(r) for r in bm["increments"]]).reshape(4, 3, 16)
        assert rel_err(orc.brownian(3141, td.times, 3, 16), ref) < 2e-15
    if bs is not None:
        td = pkg.TimeDiscretizationFromArray(0.0, 100, 0.05)
        price, proc, _ = orc.bs_european(3141, td.times, 1000, 1.0, 0.05, 0.30, 2, 5.0, 1.05)
        assert rel_err(proc[100, 0][:16], _unhex(bs["asset_at_maturity_paths_0_16"]), scale=1.0) < 1e-12
        assert abs(price - _unhex(bs["call_price"])) <= 1e-10 * abs(price)
    if lmm is not None:
        s = lmm_setup(pkg)
        fl0 = np.array([_unhex(r) for r in lmm["factor_loadings_t0"]])
        mine = s["sigma"][0][:, None] * s["factor_matrix"]
        assert rel_err(mine, fl0, scale=0.1) < 1e-10, "factor reduction (PCA) differs from the reference's"
        ref = lmm_oracle(orc, s, 2000, scheme=2)
        proc = ref.process()
        for key, window in lmm["libor_windows_paths_0_16"].items():
            t, j = (int(v) for v in key.split(","))
            assert rel_err(proc[t, j][:16], _unhex(window), scale=0.05) < 1e-12, key
        assert rel_err(ref.numeraire(5.0)[:16], _unhex(lmm["numeraire_5y_paths_0_16"])) < 1e-12
        fixing, payment = [5.0 + 0.5 * i for i in range(10)], [5.5 + 0.5 * i for i in range(10)]
        price, _, _ = ref.swaption(5.0, fixing, payment, [0.05] * 10)
        assert abs(price - _unhex(lmm["swaption_5y_into_5y_at_5pct"])) <= 1e-10 * abs(price)
        b = bermudan_spec(s)
        r = ref.bermudan(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
        assert abs(r["price"] - _unhex(lmm["bermudan_20_exercise_dates"])) <= 1e-10 * abs(r["price"])
