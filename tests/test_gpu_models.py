"""-m gpu: fused Euler kernels, generic op path, products and the regression against the oracle (the reference-shaped
CPU restatement), through the reference-facing host classes.  Tolerances are the north star's: path realizations 1e-12
relative (per-component scale), prices and regression coefficients 1e-10 relative."""
import numpy as np
import pytest

from common import rel_err, lmm_setup, lmm_device, lmm_oracle, device_process_array, bermudan_spec

pytestmark = pytest.mark.gpu

PATH_TOL = 1e-12
PRICE_TOL = 1e-10


# ---- Black-Scholes (C1) ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scheme", [0, 1, 2, 3])
@pytest.mark.parametrize("paths", [1000, 100_000])
def test_black_scholes_paths_and_price(gpu, orc, scheme, paths):
    td = gpu.TimeDiscretizationFromArray(0.0, 100, 0.05)
    bm = gpu.BrownianMotionCuda(td, 1, paths, 3141)
    model = gpu.BlackScholesModel(1.0, 0.05, 0.30, bm.randomVariableFactory)
    mc = gpu.MonteCarloAssetModel(model, gpu.EulerSchemeFromProcessModel(model, bm, scheme))
    price = gpu.EuropeanOption(5.0, 1.05).getValue(mc)
    ref_price, ref_proc, ref_vals = orc.bs_european(3141, td.times, paths, 1.0, 0.05, 0.30, scheme, 5.0, 1.05)
    assert mc.getProcess().usedFusedKernel == "black_scholes"
    for t in (0, 1, 37, 100):
        rv = mc.getAssetValue(t, 0)
        got = rv.getRealizations() if not rv.isDeterministic() else np.full(paths, rv.doubleValue())
        assert rel_err(got, ref_proc[t, 0], scale=1.0) < PATH_TOL
    assert abs(price - ref_price) <= PRICE_TOL * abs(ref_price)
    if paths == 100_000 and scheme == 2:
        # T/montecarlo/assetderivativevaluation/MonteCarloBlackScholesModelTest.java:80 — within 0.005 of the analytic value
        from scipy.stats import norm
        d1 = (np.log(1 / 1.05) + (0.05 + 0.045) * 5) / (0.3 * np.sqrt(5))
        analytic = norm.cdf(d1) - 1.05 * np.exp(-0.25) * norm.cdf(d1 - 0.3 * np.sqrt(5))
        assert abs(price - analytic) < 0.005


def test_black_scholes_generic_path_equals_fused(gpu, orc):
    td = gpu.TimeDiscretizationFromArray(0.0, 20, 0.25)
    bm = gpu.BrownianMotionCuda(td, 1, 5000, 3141)
    model = gpu.BlackScholesModel(1.0, 0.05, 0.30, bm.randomVariableFactory)
    fused = gpu.EulerSchemeFromProcessModel(model, bm)
    generic = gpu.EulerSchemeFromProcessModel(model, bm, forceGeneric=True)
    a, b = fused.getProcessValue(20, 0).getRealizations(), generic.getProcessValue(20, 0).getRealizations()
    assert generic.usedFusedKernel is None and fused.usedFusedKernel == "black_scholes"
    assert np.array_equal(a, b)                              # same device arithmetic, same order: bit-identical


# ---- Heston (C3 shape, small) -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("xi,hscheme,scheme", [(0.5, 1, 2), (0.5, 0, 2), (0.0, 1, 2), (0.5, 1, 0), (0.5, 1, 1), (0.5, 1, 3)])
def test_heston_paths_and_price(gpu, orc, xi, hscheme, scheme):
    paths, T = 20_000, 100
    td = gpu.TimeDiscretizationFromArray(0.0, T, 0.05)
    bm = gpu.BrownianMotionCuda(td, 2, paths, 31415)
    sigma = 0.30
    model = gpu.HestonModel(1.0, 0.05, sigma, 0.05, sigma * sigma, 0.1, xi, 0.1, hscheme, bm.randomVariableFactory)
    mc = gpu.MonteCarloAssetModel(model, gpu.EulerSchemeFromProcessModel(model, bm, scheme))
    price = gpu.EuropeanOption(5.0, 1.10).getValue(mc)
    ref_price, ref_proc, _ = orc.heston_european(31415, td.times, paths, 1.0, 0.05, sigma, 0.05, sigma * sigma, 0.1, xi, 0.1, hscheme, scheme, 5.0, 1.10)
    assert mc.getProcess().usedFusedKernel == "heston"
    got = device_process_array(mc, T, 2)

    def errors(ref):
        # per-component scale (S0 = 1, theta = 0.09): the variance crosses zero under full truncation, SURVEY.md 8d parity gates
        return (np.abs(got[:, 0] - ref[:, 0]) / np.maximum(np.abs(ref[:, 0]), 1.0), np.abs(got[:, 1] - ref[:, 1]) / np.maximum(np.abs(ref[:, 1]), 0.09))

    # (1) The arithmetic: against the oracle evaluated with the SAME exp / log as the kernels (diagnostic mode of the oracle: the device's
    #     functions compiled for the host) every stored value of every path is within 1e-12 - in fact bit-identical almost everywhere.
    orc.set_math(1)
    try:
        same_price, same_proc, _ = orc.heston_european(31415, td.times, paths, 1.0, 0.05, sigma, 0.05, sigma * sigma, 0.1, xi, 0.1, hscheme, scheme, 5.0, 1.10)
    finally:
        orc.set_math(0)
    e0, e1 = errors(same_proc)
    assert e0.max() < PATH_TOL and e1.max() < PATH_TOL, (e0.max(), e1.max())
    assert abs(price - same_price) <= PRICE_TOL * abs(same_price)
    # (2) Against the oracle with libm's exp / log.  sqrt(V+) has infinite slope at 0 (Feller violated for xi = 0.5), so a path that touches
    #     V ~ 0 amplifies a 1-ulp difference between two exp / log implementations.  The oracle shows exactly this between ITS two modes
    #     (libm vs the device functions, both on the CPU): ~1.5e-4 of the paths beyond 1e-12, worst ~1e-11.  The device deviates from the
    #     libm oracle by no more than the libm oracle deviates from its own second math library.
    e0, e1 = errors(ref_proc)
    l0 = np.abs(same_proc[:, 0] - ref_proc[:, 0]) / np.maximum(np.abs(ref_proc[:, 0]), 1.0)
    l1 = np.abs(same_proc[:, 1] - ref_proc[:, 1]) / np.maximum(np.abs(ref_proc[:, 1]), 0.09)
    bad_paths = np.mean(((e0 > PATH_TOL) | (e1 > PATH_TOL)).any(axis=0))
    bad_libm = np.mean(((l0 > PATH_TOL) | (l1 > PATH_TOL)).any(axis=0))
    assert max(e0.max(), e1.max()) <= max(2.0 * max(l0.max(), l1.max()), PATH_TOL) and max(e0.max(), e1.max()) < 1e-9
    assert bad_paths <= (max(2.0 * bad_libm, 1e-3) if xi > 0 else 1e-9), (bad_paths, bad_libm)   # xi = 0: deterministic variance, no kink, every path within 1e-12
    assert abs(price - ref_price) <= PRICE_TOL * abs(ref_price)


def test_heston_xi_zero_equals_black_scholes(gpu):
    """T/montecarlo/assetderivativevaluation/HestonModelTest.java:143-145: xi = 0 reproduces Black-Scholes on the same driver (1e-10)."""
    paths = 50_000
    td = gpu.TimeDiscretizationFromArray(0.0, 50, 0.1)
    bm = gpu.BrownianMotionCuda(td, 2, paths, 3141)
    f = bm.randomVariableFactory
    heston = gpu.MonteCarloAssetModel(gpu.HestonModel(1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.0, 0.1, 1, f), bm)
    bs = gpu.MonteCarloAssetModel(gpu.BlackScholesModel(1.0, 0.05, 0.3, f), bm)
    opt = gpu.EuropeanOption(5.0, 1.25)
    assert abs(opt.getValue(heston) - opt.getValue(bs)) < 1e-10


def test_heston_generic_path_equals_fused(gpu):
    td = gpu.TimeDiscretizationFromArray(0.0, 30, 0.1)
    bm = gpu.BrownianMotionCuda(td, 2, 4000, 31415)
    model = gpu.HestonModel(1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.5, 0.1, 1, bm.randomVariableFactory)
    a = gpu.EulerSchemeFromProcessModel(model, bm)
    b = gpu.EulerSchemeFromProcessModel(model, bm, forceGeneric=True)
    for c in (0, 1):
        assert np.array_equal(a.getProcessValue(30, c).getRealizations(), b.getProcessValue(30, c).getRealizations())


# ---- LIBOR market model (C4 / C5 shapes, small) ----------------------------------------------------------------------------
@pytest.mark.parametrize("scheme,measure,state", [(2, "SPOT", "LOGNORMAL"), (1, "SPOT", "LOGNORMAL"), (0, "SPOT", "LOGNORMAL"), (3, "SPOT", "LOGNORMAL"),
                                                  (2, "TERMINAL", "LOGNORMAL"), (1, "TERMINAL", "LOGNORMAL"), (2, "SPOT", "NORMAL"), (0, "TERMINAL", "NORMAL")])
def test_lmm_process_matches_oracle(gpu, orc, scheme, measure, state):
    paths = 3000
    s = lmm_setup(gpu, a=0.2 if state == "LOGNORMAL" else 0.002, d=0.3 if state == "LOGNORMAL" else 0.003)
    dev = lmm_device(gpu, s, paths, scheme=scheme, measure=measure, state_space=state)
    ref = lmm_oracle(orc, s, paths, scheme=scheme, measure=0 if measure == "SPOT" else 1, state_space=1 if state == "LOGNORMAL" else 0)
    got = device_process_array(dev, s["T"], s["N"])
    assert dev.getProcess().usedFusedKernel == "lmm"
    assert rel_err(got, ref.process(), scale=0.05) < PATH_TOL
    # frozen rates alias the previous time index (EulerSchemeFromProcessModel.java:285)
    p = dev.getProcess()
    # rate 3 fixes at T_3 = t_3: last written at time index 3, every later index is the same object
    assert p.getProcessValue(7, 3) is p.getProcessValue(3, 3) and p.getProcessValue(3, 3) is not p.getProcessValue(2, 3)


@pytest.mark.parametrize("n_factors,schemes", [(1, [2, 1]), (4, [2, 3]), (5, [2, 0]), (6, [0, 1, 2, 3]), (7, [2]), (8, [2, 1]), (12, [2, 3])])
def test_lmm_factor_counts_match_oracle(gpu, orc, n_factors, schemes):
    """Every compile-time factor count of the fused kernel (F = 1..8: factor vectors in registers) and the run-time-F kernel above that
    (factor vectors in shared memory).  T/montecarlo/interestrate/LIBORMarketModelValuationTest.java:75 uses 6 factors,
    T/montecarlo/interestrate/HullWhiteModelTest.java:93 uses 1 for its LMM; all four schemes at F = 6, both measures."""
    s = lmm_setup(gpu, n_libors=20, n_factors=n_factors)
    for scheme in schemes:
        for paths, measure in ((1500, "SPOT"), (333, "TERMINAL")):
            dev = lmm_device(gpu, s, paths, scheme=scheme, measure=measure)
            ref = lmm_oracle(orc, s, paths, scheme=scheme, measure=0 if measure == "SPOT" else 1)
            got = device_process_array(dev, s["T"], s["N"])
            assert dev.getProcess().usedFusedKernel == "lmm"
            assert dev.getProcess().getNumberOfFactors() == n_factors
            assert rel_err(got, ref.process(), scale=0.05) < PATH_TOL, (n_factors, scheme, measure)
    if n_factors == 6:                                       # the op-by-op device path gives the same bits as the fused F = 6 kernel
        a = lmm_device(gpu, s, 700, scheme=2)
        b = lmm_device(gpu, s, 700, scheme=2, force_generic=True)
        assert np.array_equal(device_process_array(a, s["T"], s["N"]), device_process_array(b, s["T"], s["N"]))


def test_lmm_generic_path_equals_fused(gpu):
    s = lmm_setup(gpu, n_libors=10)
    a = lmm_device(gpu, s, 2000, scheme=1)
    b = lmm_device(gpu, s, 2000, scheme=1, force_generic=True)
    ga, gb = device_process_array(a, s["T"], s["N"]), device_process_array(b, s["T"], s["N"])
    assert b.getProcess().usedFusedKernel is None
    assert np.array_equal(ga, gb)


def test_lmm_ragged_and_capped(gpu, orc):
    for paths, cap in [(1, 1e5), (129, 1e5), (1000, 0.08)]:
        s = lmm_setup(gpu, n_libors=12, n_factors=2)
        dev = lmm_device(gpu, s, paths, libor_cap=cap)
        ref = lmm_oracle(orc, s, paths, libor_cap=cap)
        assert rel_err(device_process_array(dev, s["T"], s["N"]), ref.process(), scale=0.05) < PATH_TOL


@pytest.mark.parametrize("case", ["zero_forward", "huge_vol", "huge_vol_terminal", "cap_zero", "cap_negative"])
@pytest.mark.parametrize("scheme", [2, 3])
def test_lmm_special_values_take_the_cold_paths(gpu, orc, case, scheme):
    """Arguments outside the fast paths of the fused kernel's log / reciprocal / exp / min (zero and subnormal rates, exp underflow, a zero or
    negative cap with its NaN propagation): the cold fix-up branches must reproduce the reference's special-value semantics
    (Math.log(0) = -inf, Math.exp(-inf) = 0, Math.min NaN-propagating; LIBORMarketModelFromCovarianceModel.java:1085, :1146-1160)."""
    big = case.startswith("huge_vol")
    s = lmm_setup(gpu, n_libors=8, n_factors=2, a=60.0 if big else 0.2, d=60.0 if big else 0.3)
    if case == "zero_forward":
        s["L0"][2] = 0.0
        s["L0"][5] = 0.0
        for i in range(s["N"]):
            s["df"][i + 1] = s["df"][i] / (1.0 + s["L0"][i] * s["tenor"].getTimeStep(i))
    cap = {"cap_zero": 0.0, "cap_negative": -1.0}.get(case, 1e5)
    measure = "TERMINAL" if case == "huge_vol_terminal" else "SPOT"
    for paths in (1, 257):
        dev = lmm_device(gpu, s, paths, scheme=scheme, measure=measure, libor_cap=cap)
        ref = lmm_oracle(orc, s, paths, scheme=scheme, measure=0 if measure == "SPOT" else 1, libor_cap=cap)
        got, want = device_process_array(dev, s["T"], s["N"]), ref.process()
        assert dev.getProcess().usedFusedKernel == "lmm"
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
        assert rel_err(got, want, scale=0.05) < PATH_TOL
        if case == "zero_forward":
            assert np.all(got[1:3, 2] == 0.0)
        if case == "cap_negative":
            assert np.isnan(got).any()


@pytest.mark.parametrize("n_libors,period", [(130, 0.25), (260, 0.125)])
def test_lmm_many_rates_shrinks_the_block(gpu, orc, n_libors, period):
    """More forward rates than a 128-thread block can keep in shared memory (8*N bytes per thread): the fused kernel runs with 128-thread
    blocks at one block per SM (N = 130) and with 64-thread blocks (N = 260); the warps still pull 32-path tiles."""
    s = lmm_setup(gpu, n_libors=n_libors, n_factors=3, period=period, dt=1.0, horizon=8.0)
    for paths in (97, 700):
        dev = lmm_device(gpu, s, paths, scheme=2)
        ref = lmm_oracle(orc, s, paths, scheme=2)
        got = device_process_array(dev, s["T"], s["N"])
        assert dev.getProcess().usedFusedKernel == "lmm"
        assert rel_err(got, ref.process(), scale=0.05) < PATH_TOL


def test_lmm_numeraire_forward_rate_swaption_caplet(gpu, orc):
    paths = 20_000
    s = lmm_setup(gpu)
    dev = lmm_device(gpu, s, paths, scheme=1)
    ref = lmm_oracle(orc, s, paths, scheme=1)
    assert rel_err(dev.getNumeraire(5.0).getRealizations(), ref.numeraire(5.0)) < PATH_TOL
    assert rel_err(dev.getForwardRate(5.0, 5.0, 10.0).getRealizations(), ref.forward_rate(5.0, 5.0, 10.0), scale=0.05) < PATH_TOL
    fixing = [5.0 + 0.5 * i for i in range(10)]
    payment = [5.5 + 0.5 * i for i in range(10)]
    swaption = gpu.Swaption(5.0, fixing, payment, [0.05] * 10)
    price = swaption.getValue(dev)
    ref_price, ref_vals, ref_se = ref.swaption(5.0, fixing, payment, [0.05] * 10)
    assert abs(price - ref_price) <= PRICE_TOL * abs(ref_price)
    vals = swaption.getValue(0.0, dev)
    assert abs(vals.getStandardError() - ref_se) <= 1e-9 * ref_se
    caplet = gpu.Caplet(5.0, 0.5, 0.05)
    ref_c, _ = ref.caplet(5.0, 0.5, 0.05)
    assert abs(caplet.getValue(dev) - ref_c) <= PRICE_TOL * abs(ref_c)


@pytest.mark.parametrize("measure,method", [("SPOT", "LOG_LINEAR_UNCORRECTED"), ("SPOT", "LINEAR"), ("TERMINAL", "LOG_LINEAR_UNCORRECTED")])
def test_lmm_forward_rate_and_numeraire_between_tenor_points(gpu, orc, measure, method):
    """Interpolation on fractional tenor points: LIBORMarketModelFromCovarianceModel.getForwardRate :1244-1281 with
    getOnePlusInterpolatedLIBORDt :1324-1395 (LINEAR and the default LOG_LINEAR_UNCORRECTED), the unadjusted numeraire between tenor
    points :969-1006 and the log-linear interpolation of the discount-curve adjustment :886-905; simulation times off the grid are rounded
    to the nearest point (:1248-1253)."""
    paths = 5000
    # (under the terminal measure the reference reads the LIBORs at the requested time itself, :986, so that time must be a simulation time:
    #  a quarter-year simulation grid under the half-year tenor)
    s = lmm_setup(gpu, dt=0.5 if measure == "SPOT" else 0.25)
    tenor = s["tenor"]
    s["df"] = np.array([np.exp(-0.035 * tenor.getTime(i)) for i in range(s["N"] + 1)])
    factory = gpu.RandomVariableCudaFactory()
    props = {"measure": measure, "stateSpace": "LOGNORMAL", "interpolationMethod": method}
    model = gpu.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], factory, s["cov"], None, props)
    dev = gpu.LIBORMonteCarloSimulationFromLIBORModel(gpu.EulerSchemeFromProcessModel(model, gpu.BrownianMotionCuda(s["sim"], s["F"], paths, 3141, factory)))
    ref = lmm_oracle(orc, s, paths, measure=0 if measure == "SPOT" else 1)
    ref.set_interpolation(0 if method == "LINEAR" else 1)
    cases = [(2.0, 2.0, 2.3),            # period end between tenor points
             (2.0, 2.2, 2.5),            # period start between tenor points, end = the next tenor point
             (1.5, 2.2, 4.0),            # start between tenor points, several whole periods behind it
             (3.0, 3.1, 5.3),            # both ends off the grid
             (2.7, 3.25, 3.75),          # simulation time off the grid too (rounded to the nearest point), half-period shifted LIBOR
             (4.26, 6.0, 6.4)]
    for t, a, b in cases:
        got = dev.getForwardRate(t, a, b).getRealizations()
        want = ref.forward_rate(t, a, b)
        assert rel_err(got, want, scale=0.05) < PATH_TOL, (t, a, b)
    for t in ((2.3, 0.7, 7.25, 2.0) if measure == "SPOT" else (2.25, 0.75, 7.25, 2.0)):
        assert rel_err(dev.getNumeraire(t).getRealizations(), ref.numeraire(t)) < PATH_TOL, t
    if measure == "TERMINAL":                                # a time that is not a simulation time: the reference indexes out of bounds there
        with pytest.raises(IndexError):
            dev.getNumeraire(2.3)


def test_hull_white_numeraire_between_simulation_times(gpu, orc):
    """HullWhiteModel.getNumeraire :317-333: log-linear interpolation between the neighbouring simulation times (a caplet paying between
    two simulation dates needs it)."""
    paths = 20_000
    td = gpu.TimeDiscretizationFromArray(0.0, 20, 0.5)
    tenor = gpu.TimeDiscretizationFromArray(0.0, 40, 0.25)
    vt = np.arange(0, 11.0)
    vol, mr = 0.005 + 0.0005 * np.floor(vt) / 20, np.full(vt.size, 0.1)
    ct = tenor.times
    df = np.exp(-(0.03 + 0.01 * ct / 40.0) * ct)
    vm = gpu.ShortRateVolatilityModelAsGiven(gpu.TimeDiscretizationFromArray(vt), vol, mr)
    bm = gpu.BrownianMotionCuda(td, 2, paths, 3141)
    model = gpu.HullWhiteModel(bm.randomVariableFactory, tenor, vm, None, df, df)
    sim = gpu.LIBORMonteCarloSimulationFromLIBORModel(model, gpu.EulerSchemeFromProcessModel(model, bm, 0))
    price = gpu.Caplet(5.0, 0.25, 0.03).getValue(sim)                       # pays at 5.25, between the simulation times 5.0 and 5.5
    ref_price, _, ref_num, ref_fr = orc.hull_white_caplet(3141, td.times, paths, vt, vol, mr, ct, df, df, 0, 5.0, 0.25, 0.03)
    assert rel_err(sim.getNumeraire(5.25).getRealizations(), ref_num) < PATH_TOL
    assert rel_err(sim.getForwardRate(5.0, 5.0, 5.25).getRealizations(), ref_fr, scale=0.03) < 1e-11
    assert abs(price - ref_price) <= PRICE_TOL * abs(ref_price)


def test_multi_period_forward_rate_in_one_kernel_equals_the_accrue_loop(gpu):
    """getForwardRate over several tenor periods (:1288-1302) runs as ONE kernel (fmb_rv_accrue_chain) instead of one accrue() pass per
    period: same operations, same order, so the bits equal the op-by-op loop - also beyond 32 periods (chunked)."""
    for n_libors, period, t_index, start, end in ((40, 0.5, 10, 5.0, 15.0), (80, 0.25, 1, 0.25, 20.0)):
        s = lmm_setup(gpu, n_libors=n_libors, period=period, dt=period)
        dev = lmm_device(gpu, s, 3001)
        model, process = dev.getModel(), dev.getProcess()
        fused = dev.getForwardRate(start, start, end)
        k0, k1 = model.getLiborPeriodIndex(start), model.getLiborPeriodIndex(end)
        acc = None
        for k in range(k0, k1):
            libor, sub = model.getLIBOR(process, t_index, k), model.getLiborPeriod(k + 1) - model.getLiborPeriod(k)
            acc = libor.mult(sub).add(1.0) if acc is None else acc.accrue(libor, sub)
        loop = acc.sub(1.0).div(end - start)
        assert np.array_equal(fused.getRealizations(), loop.getRealizations())
        assert fused.getFiltrationTime() == loop.getFiltrationTime()


def test_swaption_discounting_adjustment_with_a_separate_discount_curve(gpu, orc):
    """Swaption.java:160-171: with a discount curve that is NOT the one implied by the forward curve every period's value is scaled by
    forwardBondOnForwardCurve / forwardBondOnDiscountCurve.  The oracle gets the adjustments as an explicit array (computed here from
    the two curves), the device-side product derives them from the model's curves."""
    paths = 20_000
    s = lmm_setup(gpu)
    tenor = s["tenor"]
    s["df"] = np.array([np.exp(-0.03 * tenor.getTime(i) - 0.0005 * tenor.getTime(i) ** 2) for i in range(s["N"] + 1)])     # OIS-like curve below the 5 % forwards
    dev = lmm_device(gpu, s, paths, scheme=2)
    ref = lmm_oracle(orc, s, paths, scheme=2)
    fixing = [5.0 + 0.5 * i for i in range(10)]
    payment = [5.5 + 0.5 * i for i in range(10)]
    implied = [1.0]
    for i in range(s["N"]):
        implied.append(implied[-1] / (1.0 + s["L0"][i] * tenor.getTimeStep(i)))
    adj = []
    for f, p in zip(fixing, payment):
        i0, i1 = tenor.getTimeIndex(max(f, 5.0)), tenor.getTimeIndex(p)
        adj.append((implied[i0] / implied[i1]) / (s["df"][i0] / s["df"][i1]))
    assert max(abs(a - 1.0) for a in adj) > 1e-3                      # the adjustment matters in this set-up
    price = gpu.Swaption(5.0, fixing, payment, [0.05] * 10).getValue(dev)
    ref_price, _, _ = ref.swaption(5.0, fixing, payment, [0.05] * 10, discounting_adjustments=adj)
    assert abs(price - ref_price) <= PRICE_TOL * abs(ref_price)
    plain, _, _ = ref.swaption(5.0, fixing, payment, [0.05] * 10)
    assert abs(price - plain) > 1e-6 * abs(plain)


def test_fused_kernels_are_not_used_with_a_mismatching_driver(gpu):
    """The fused Heston / Hull-White kernels read exactly two increments per step and the fused LMM kernel takes the factor count from
    the driver: a driver that does not match the model's tables must run the generic loop (the reference's addSumProduct over whatever
    loadings the model returns), never a fused kernel with wrong strides."""
    td = gpu.TimeDiscretizationFromArray(0.0, 30, 0.1)
    bm3 = gpu.BrownianMotionCuda(td, 3, 3000, 31415)
    model = gpu.HestonModel(1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.5, 0.1, 1, bm3.randomVariableFactory)
    a = gpu.EulerSchemeFromProcessModel(model, bm3)
    b = gpu.EulerSchemeFromProcessModel(model, gpu.BrownianMotionView(bm3, [0, 1]))
    for c in (0, 1):
        assert np.array_equal(a.getProcessValue(30, c).getRealizations(), b.getProcessValue(30, c).getRealizations())
    assert a.usedFusedKernel is None and b.usedFusedKernel is None
    # LMM tables built for 3 factors, driver with 2: no fused kernel
    s = lmm_setup(gpu, n_libors=8, n_factors=3)
    factory = gpu.RandomVariableCudaFactory()
    lmm = gpu.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], factory, s["cov"], None, {})
    proc = gpu.EulerSchemeFromProcessModel(lmm, gpu.BrownianMotionCuda(s["sim"], 2, 500, 3141, factory))
    assert lmm.getFusedSpecification(proc) is None
    # covariance model on a finer grid than the process: rows are mapped by time (AbstractLIBORCovarianceModel.java:70-76), fused == generic
    s2 = lmm_setup(gpu, n_libors=8, n_factors=2, dt=0.25)
    coarse = gpu.TimeDiscretizationFromArray(0.0, 8, 0.5)
    lmm2 = gpu.LIBORMarketModelFromCovarianceModel.of(s2["tenor"], None, s2["L0"], s2["df"], factory, s2["cov"], None, {})
    bm = gpu.BrownianMotionCuda(coarse, 2, 700, 3141, factory)
    f, g = gpu.EulerSchemeFromProcessModel(lmm2, bm), gpu.EulerSchemeFromProcessModel(lmm2, bm, forceGeneric=True)
    assert np.array_equal(f.getProcessValue(5, 7).getRealizations(), g.getProcessValue(5, 7).getRealizations())
    assert f.usedFusedKernel == "lmm" and g.usedFusedKernel is None


@pytest.mark.parametrize("scheme", [2, 1])
def test_lmm_bermudan_swaption_matches_oracle(gpu, orc, scheme):
    paths = 20_000
    s = lmm_setup(gpu)
    b = bermudan_spec(s)
    dev = lmm_device(gpu, s, paths, scheme=scheme)
    ref = lmm_oracle(orc, s, paths, scheme=scheme)
    product = gpu.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
    res = product.getValues(0.0, dev)
    price = res["value"].getAverage()
    r = ref.bermudan(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
    assert abs(price - r["price"]) <= PRICE_TOL * abs(r["price"])
    assert abs(res["error"] - r["std_error"]) <= 1e-8 * r["std_error"]
    # exercise decisions: sign flips of a near-zero trigger are legitimate at the 1e-13 level -> count them
    ex = res["exerciseTime"].getRealizations()
    mism = int(np.sum(ex != r["exercise_time"]))
    assert mism <= max(2, paths // 5000), mism
    # regression coefficients: deviation bounded by cond(XtX) * a few ulp (SURVEY.md §7 "hard parts"); 1e-10 where conditioning allows
    for e, est in enumerate(product.lastRegressions):
        x, xr, cond = est.lastParameters, r["regression"][e], r["cond"][e]
        dev_rel = np.max(np.abs(x - xr)) / np.max(np.abs(xr))
        assert dev_rel <= max(1e-10, 50 * cond * 2.0 ** -53), (e, dev_rel, cond)
        # what matters: the fitted conditional expectation
    # fitted values agree
    assert rel_err(res["value"].getRealizations(), r["values"], scale=1e-2) < 1e-9 or mism > 0


def test_regression_moments_and_solver(gpu, orc):
    RV = gpu.RandomVariableCuda
    rng = np.random.default_rng(3)
    n = 200_000
    b1, b2 = rng.random(n) + 0.5, rng.standard_normal(n)
    y = 2.0 + 3.0 * b1 - 0.5 * b2 + 0.01 * rng.standard_normal(n)
    basis = [gpu.RandomVariableFromDoubleArray(1.0), RV(0.0, b1), RV(0.0, b2), RV(0.0, b1).pow(2.0)]
    est = gpu.MonteCarloConditionalExpectationRegression(basis)
    x = est.getLinearRegressionParameters(RV(0.0, y))
    X = np.stack([np.ones(n), b1, b2, b1 * b1], axis=1)
    xtx = np.array([[orc.rv_reduce(0, X[:, i] * X[:, j]) for j in range(4)] for i in range(4)])
    xty = np.array([orc.rv_reduce(0, y * X[:, i]) for i in range(4)])
    xr, cond = orc.solve_pinv(xtx, xty)
    assert np.max(np.abs(x - xr)) <= max(1e-10, 50 * cond * 2.0 ** -53) * np.max(np.abs(xr))
    ce = est.getConditionalExpectation(RV(0.0, y)).getRealizations()
    assert rel_err(ce, X @ x) < 1e-13
    # rank-deficient system: the SVD cut-off gives the minimum-norm solution instead of blowing up
    est2 = gpu.MonteCarloConditionalExpectationRegression([RV(0.0, b1), RV(0.0, b1)])
    x2 = est2.getLinearRegressionParameters(RV(0.0, 4.0 * b1))
    assert np.allclose(x2, [2.0, 2.0], atol=1e-9)


def test_device_resident_regression_equals_the_host_resident_one(gpu):
    """fmb_regression_conditional_expectation (moments, solve and prediction queued on the device, nothing returns to the host) against the
    host-resident sequence fmb_regression_moments -> fmb_regression_solve_svd -> fmb_regression_predict: same moments, same Jacobi code on
    both sides, so the coefficients and fitted values agree to the last bit; the solver (XtX) is cached per estimator (:125-138)."""
    import ctypes as C
    nv = gpu.native
    RV = gpu.RandomVariableCuda
    rng = np.random.default_rng(11)
    for n, K in ((1, 2), (257, 3), (300_001, 6), (50_000, 8)):
        cols = [rng.random(n) + 0.25 * k for k in range(K - 1)]
        basis = [gpu.RandomVariableFromDoubleArray(1.0)] + [RV(0.0, c) for c in cols]
        y = RV(1.0, sum((k + 1.0) * c for k, c in enumerate(cols)) + 0.1 * rng.standard_normal(n))
        est = gpu.MonteCarloConditionalExpectationRegression(basis)
        ce = est.getConditionalExpectation(y)
        assert est._lastFit is not None                       # took the device-resident path
        x_dev, cond_dev = est.lastParameters.copy(), est.lastConditionNumber
        b = [est._as_cuda(v, y.shard) for v in basis]
        XTX, XTy = est._moments(b, y)
        x = np.zeros(K)
        cond = C.c_double()
        nv.check(nv.load().fmb_regression_solve_svd(K, nv.dptr(nv.as_f64(XTX)), nv.dptr(nv.as_f64(XTy)), nv.dptr(x), C.byref(cond)))
        assert np.array_equal(x, x_dev) and cond.value == cond_dev, (n, K, x, x_dev)
        hs, sc = est._basis_args(b)
        out = C.c_uint64()
        nv.check(nv.load().fmb_regression_predict(K, nv.hptr(hs), nv.dptr(sc), nv.dptr(x), C.byref(out)))
        assert np.array_equal(nv.DeviceVector(out.value, n).download(), ce.getRealizations())
        # second dependent on the same estimator: XtX comes from the first fit
        y2 = RV(1.0, rng.standard_normal(n))
        est.getConditionalExpectation(y2)
        XtX1, XtX2 = np.zeros(K * K), np.zeros(K * K)
        nv.check(nv.load().fmb_regression_fit_get(est._cachedFit.h, K, nv.dptr(XtX1), None, None, None))
        nv.check(nv.load().fmb_regression_fit_get(est._lastFit.h, K, nv.dptr(XtX2), None, None, None))
        assert np.array_equal(XtX1, XtX2) and est._lastFit is not est._cachedFit


# ---- Hull-White (C2 shape, small) -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scheme", [0, 2, 1])
def test_hull_white_process_matches_oracle(gpu, orc, scheme):
    paths = 20_000
    td = gpu.TimeDiscretizationFromArray(0.0, 200, 0.1)
    vt = np.arange(0, 21.0)
    vol, mr = 0.005 + 0.0005 * np.floor(vt) / 20, np.full(vt.size, 0.1)
    vm = gpu.ShortRateVolatilityModelAsGiven(gpu.TimeDiscretizationFromArray(vt), vol, mr)
    bm = gpu.BrownianMotionCuda(td, 2, paths, 3141)
    model = gpu.HullWhiteModel(bm.randomVariableFactory, gpu.TimeDiscretizationFromArray(0.0, 40, 0.5), vm)
    process = gpu.EulerSchemeFromProcessModel(model, bm, scheme)
    ref, _ = orc.hull_white_process(3141, td.times, paths, vt, vol, mr, scheme)
    got = np.empty_like(ref)
    for t in range(201):
        for c in range(2):
            rv = process.getProcessValue(t, c)
            got[t, c] = rv.doubleValue() if rv.isDeterministic() else rv.getRealizations()
    assert process.usedFusedKernel == ("hull_white" if scheme != 1 else None)
    # short-rate state ~ 1e-2, log-numeraire ~ 1e-1: absolute scales for the relative test
    assert rel_err(got[:, 0], ref[:, 0], scale=1e-2) < PATH_TOL and rel_err(got[:, 1], ref[:, 1], scale=1e-1) < PATH_TOL


def test_hull_white_caplet_numeraire_forward_rate(gpu, orc):
    """C2: caplet on the Hull-White model through Caplet + LIBORMonteCarloSimulationFromLIBORModel (unchanged product code)."""
    paths = 50_000
    td = gpu.TimeDiscretizationFromArray(0.0, 40, 0.5)
    tenor = gpu.TimeDiscretizationFromArray(0.0, 40, 0.5)
    vt = np.arange(0, 21.0)
    vol, mr = 0.005 + 0.0005 * np.floor(vt) / 20, np.full(vt.size, 0.1)
    ct = tenor.times
    zero = 0.03 + 0.01 * (np.clip(ct, 0.5, 40.0) - 0.5) / 39.5
    df = np.exp(-zero * ct)
    vm = gpu.ShortRateVolatilityModelAsGiven(gpu.TimeDiscretizationFromArray(vt), vol, mr)
    bm = gpu.BrownianMotionCuda(td, 2, paths, 3141)
    model = gpu.HullWhiteModel(bm.randomVariableFactory, tenor, vm, None, df, df)
    sim = gpu.LIBORMonteCarloSimulationFromLIBORModel(model, gpu.EulerSchemeFromProcessModel(model, bm, 0))
    price = gpu.Caplet(5.0, 0.5, 0.03).getValue(sim)
    ref_price, _, ref_num, ref_fr = orc.hull_white_caplet(3141, td.times, paths, vt, vol, mr, ct, df, df, 0, 5.0, 0.5, 0.03)
    assert rel_err(sim.getNumeraire(5.5).getRealizations(), ref_num) < PATH_TOL
    assert rel_err(sim.getForwardRate(5.0, 5.0, 5.5).getRealizations(), ref_fr, scale=0.03) < 1e-11
    assert abs(price - ref_price) <= PRICE_TOL * abs(ref_price)
    # zero bond reproduces the curve exactly through the control variate (HullWhiteModel.java:340-341)
    assert abs(sim.getNumeraire(5.5).invert().getAverage() - df[11]) < 1e-13


# ---- FAST floating-point mode -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scheme,cap", [(2, 1e5), (3, 1e5), (0, 1e5), (1, 1e5), (2, 0.08), (3, 0.08)])
def test_fast_mode_lmm_within_path_tolerance(gpu, orc, scheme, cap):
    """fmb_set_fp_mode(1): FMA contraction + carried log-state.  Still within the 1e-12 path tolerance of the north star."""
    paths = 3000
    s = lmm_setup(gpu)
    ref = lmm_oracle(orc, s, paths, scheme=scheme, libor_cap=cap).process()
    gpu.native.set_fp_mode(1)
    try:
        dev = lmm_device(gpu, s, paths, scheme=scheme, libor_cap=cap)
        got = device_process_array(dev, s["T"], s["N"])
    finally:
        gpu.native.set_fp_mode(0)
    strict = device_process_array(lmm_device(gpu, s, paths, scheme=scheme, libor_cap=cap), s["T"], s["N"])
    assert rel_err(got, ref, scale=0.05) < PATH_TOL
    assert not np.array_equal(got, strict)                      # it really is a different arithmetic
    assert rel_err(strict, ref, scale=0.05) < PATH_TOL


def test_fast_mode_black_scholes(gpu, orc):
    td = gpu.TimeDiscretizationFromArray(0.0, 100, 0.05)
    ref_price, ref_proc, _ = orc.bs_european(3141, td.times, 20000, 1.0, 0.05, 0.30, 2, 5.0, 1.05)
    gpu.native.set_fp_mode(1)
    try:
        mc = gpu.MonteCarloBlackScholesModel(td, 20000, 1.0, 0.05, 0.30)
        price = gpu.EuropeanOption(5.0, 1.05).getValue(mc)
        got = mc.getAssetValue(5.0, 0).getRealizations()
    finally:
        gpu.native.set_fp_mode(0)
    assert rel_err(got, ref_proc[100, 0]) < PATH_TOL and abs(price - ref_price) <= PRICE_TOL * abs(ref_price)


# ---- localized regression, LinearRegression, asset BermudanOption ------------------------------------------
@pytest.mark.parametrize("binning,nbasis", [(False, 5), (True, 20)])
def test_asset_bermudan_option_matches_oracle(gpu, orc, binning, nbasis):
    paths = 20_000
    td = gpu.TimeDiscretizationFromArray(0.0, 20, 0.25)
    mc = gpu.MonteCarloBlackScholesModel(td, paths, 1.0, 0.05, 0.30)
    dates, notionals, strikes = [1.0, 2.0, 3.0, 4.0, 5.0], [1.0] * 5, [1.05] * 5
    product = gpu.BermudanOption(dates, notionals, strikes, numberOfBasisFunctions=nbasis, useBinning=binning)
    price = product.getValue(mc)
    r = orc.bs_bermudan_option(3141, td.times, paths, 1.0, 0.05, 0.30, 2, dates, notionals, strikes, n_basis=nbasis, binning=binning)
    # the monomial basis 1, S, ..., S^4 is ill-conditioned (cond ~ 1e8) and pow() differs by an ulp between libm and the device:
    # a handful of paths with |continuation - exercise| ~ 1e-9 may flip; each flip moves the average by < 1e-9 / paths
    ex = product.lastValuationExerciseTime.getRealizations()
    flips = int(np.sum(ex != r["exercise_time"]))
    assert flips <= 3, flips
    assert abs(price - r["price"]) <= (PRICE_TOL if flips == 0 else 1e-8) * abs(r["price"])


def test_localized_regression_and_linear_regression(gpu, orc):
    RV = gpu.RandomVariableCuda
    rng = np.random.default_rng(11)
    n = 100_000
    b1 = rng.standard_normal(n)
    y = 1.0 + 2.0 * b1 + 0.3 * rng.standard_normal(n)
    y[:50] += 40.0                                                              # outliers the localization must cut away
    basis_np = np.stack([np.ones(n), b1])
    est = gpu.MonteCarloConditionalExpectationRegressionLocalizedOnDependents([RV(0.0, basis_np[0]), RV(0.0, b1)], standardDeviations=3.0)
    x = est.getLinearRegressionParameters(RV(0.0, y))
    xr, cer = orc.regression_localized(basis_np, y, 3.0)
    assert np.max(np.abs(x - xr)) <= 1e-10 * np.max(np.abs(xr))
    assert rel_err(est.getConditionalExpectation(RV(0.0, y)).getRealizations(), cer) < 1e-10
    plain = gpu.MonteCarloConditionalExpectationRegression([RV(0.0, basis_np[0]), RV(0.0, b1)]).getLinearRegressionParameters(RV(0.0, y))
    assert abs(x[0] - 1.0) < abs(plain[0] - 1.0)                                # the outliers bias the plain regression, not the localized one
    lr = gpu.LinearRegression([RV(0.0, b1)]).getRegressionCoefficients(RV(0.0, y))
    assert abs(lr[0] - np.mean(y * b1) / np.mean(b1 * b1)) < 1e-12
    lr3 = gpu.LinearRegression([RV(0.0, np.ones(n)), RV(0.0, b1), RV(0.0, b1 * b1)]).getRegressionCoefficients(RV(0.0, y))
    X = np.stack([np.ones(n), b1, b1 * b1], axis=1)
    assert np.allclose(lr3, np.linalg.lstsq(X, y, rcond=None)[0], rtol=1e-9, atol=1e-10)


def test_products_price_identically_with_deferred_arithmetic(gpu):
    """Bermudan swaption, swaption and caplet through the unchanged product algebra, once with every element-wise operation deferred
    into fused chains (fmb_rv_eval_chain) and once with one kernel per operation: identical bits, fewer launches."""
    nvm = gpu.native
    s = lmm_setup(gpu)
    b = bermudan_spec(s)
    keep = nvm.lazy_min_n()
    prices, launches = [], []
    for lazy in (True, False):
        nvm.set_lazy(lazy, min_n=0)
        try:
            sim = lmm_device(gpu, s, 2500)
            sim.getProcess().getProcessValue(s["T"], s["N"] - 1)
            l0 = nvm.launch_count()
            berm = gpu.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"]).getValue(sim)
            swpt = gpu.Swaption(1.0, [1.0, 1.5, 2.0, 2.5], [1.5, 2.0, 2.5, 3.0], [0.05] * 4).getValue(sim)
            capl = gpu.Caplet(2.0, 0.5, 0.05).getValue(sim)
            launches.append(nvm.launch_count() - l0)
            prices.append((berm, swpt, capl))
        finally:
            nvm.set_lazy(True, min_n=keep)
    assert prices[0] == prices[1]
    assert launches[0] < 0.6 * launches[1], launches
