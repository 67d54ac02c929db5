"""-m gpu: BASELINE.json's full sizes.  The oracle cannot run 4 M paths in seconds, so the full-size device results are checked
through size-independent properties: (1) any window of paths of the big simulation equals the oracle started at that path offset
(the single MT19937 stream makes every path individually addressable), (2) martingale / curve-reproduction identities."""
import numpy as np
import pytest

from common import rel_err, lmm_setup, lmm_device, lmm_oracle

pytestmark = pytest.mark.gpu


def test_c4_lmm_4m_paths_windows_match_oracle(gpu, orc):
    P = 4_000_000
    s = lmm_setup(gpu)
    sim = lmm_device(gpu, s, P, scheme=2)
    proc = sim.getProcess()
    picks = [(1, 39), (7, 8), (20, 21), (20, 39), (39, 39), (40, 39)]
    full = {tj: proc.getProcessValue(*tj).getRealizations() for tj in picks}
    assert proc.usedFusedKernel == "lmm"
    for off in (0, 1_999_473, P - 1000):
        ref = lmm_oracle(orc, s, 1000, scheme=2, path_offset=off).process()
        for (t, j), v in full.items():
            assert rel_err(v[off:off + 1000], ref[t, j], scale=0.05) < 1e-12, (off, t, j)
    # martingale property under the spot measure: E[1/N(T)] reproduces the initial discount factors within the Monte-Carlo error
    for T_, i in ((5.0, 10), (10.0, 20)):
        n = sim.getNumeraire(T_)
        zb = n.invert()
        assert abs(zb.getAverage() - s["df"][i]) < 5 * zb.getStandardError() + 1e-12


def test_c2_hull_white_1m_paths_windows_match_oracle(gpu, orc):
    P = 1_000_000
    td = gpu.TimeDiscretizationFromArray(0.0, 200, 0.1)
    vt = np.arange(0, 21.0)
    vol, mr = 0.005 + 0.0005 * np.floor(vt) / 20, np.full(vt.size, 0.1)
    vm = gpu.ShortRateVolatilityModelAsGiven(gpu.TimeDiscretizationFromArray(vt), vol, mr)
    bm = gpu.BrownianMotionCuda(td, 2, P, 3141)
    process = gpu.EulerSchemeFromProcessModel(gpu.HullWhiteModel(bm.randomVariableFactory, gpu.TimeDiscretizationFromArray(0.0, 40, 0.5), vm), bm, 0)
    x0, x1 = process.getProcessValue(200, 0).getRealizations(), process.getProcessValue(200, 1).getRealizations()
    for off in (0, P - 500):
        ref, _ = orc.hull_white_process(3141, td.times, 500, vt, vol, mr, 0, path_offset=off)
        assert rel_err(x0[off:off + 500], ref[200, 0], scale=1e-2) < 1e-12 and rel_err(x1[off:off + 500], ref[200, 1], scale=1e-1) < 1e-12


def test_c3_heston_4m_paths_1000_steps_window_and_xi_zero_identity(gpu, orc):
    """4 M paths x 1000 steps x 2 components (64 GB of process values + 64 GB of increments on one GPU)."""
    P, T = 4_000_000, 1000
    td = gpu.TimeDiscretizationFromArray(0.0, T, 0.005)
    bm = gpu.BrownianMotionCuda(td, 2, P, 31415)
    f = bm.randomVariableFactory
    model = gpu.HestonModel(1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.0, 0.1, 1, f)          # xi = 0: no kink, every path must agree
    mc = gpu.MonteCarloAssetModel(model, bm)
    sT = mc.getAssetValue(5.0, 0).getRealizations()
    off = 1_000_003
    _, ref, _ = orc.heston_european(31415, td.times, 200, 1.0, 0.05, 0.3, 0.05, 0.09, 0.1, 0.0, 0.1, 1, 2, 5.0, 1.1, path_offset=off)
    assert rel_err(sT[off:off + 200], ref[T, 0]) < 1e-12
    # Heston(xi = 0) == Black-Scholes on the same driver to 1e-10 (HestonModelTest.java:143-145), at full size
    opt = gpu.EuropeanOption(5.0, 1.1)
    heston_value = opt.getValue(mc)
    del mc, model, sT                                          # give the 64 GB of Heston process values back before the next 32 GB
    bs = gpu.MonteCarloAssetModel(gpu.BlackScholesModel(1.0, 0.05, 0.3, f), bm)
    assert abs(heston_value - opt.getValue(bs)) < 1e-10


def test_c5_bermudan_8m_paths_price_is_consistent_with_1m(gpu):
    """8 M paths on one GPU (58 GB): the price agrees with the 1 M-path price within 4 combined standard errors, and lies above
    the European swaption on the first exercise date (a Bermudan is worth at least its most valuable European)."""
    from common import bermudan_spec
    s = lmm_setup(gpu)
    b = bermudan_spec(s)
    product = gpu.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
    res = {}
    for P in (1_000_000, 8_000_000):
        sim = lmm_device(gpu, s, P, scheme=2)
        r = product.getValues(0.0, sim)
        res[P] = (r["value"].getAverage(), r["error"])
        if P == 8_000_000:
            european = gpu.Swaption(b["fixing"][0], b["fixing"], b["payment"], b["swaprates"]).getValue(0.0, sim)
            assert res[P][0] > european.getAverage() - 3 * european.getStandardError()
        del sim, r
    (v1, e1), (v8, e8) = res[1_000_000], res[8_000_000]
    assert abs(v1 - v8) < 4 * np.hypot(e1, e8)
    assert e8 < e1 / 2.5                                       # standard error shrinks like 1/sqrt(8)


def test_c5_bermudan_8m_paths_windows_match_oracle(gpu, orc):
    """The north-star configuration at full size on one GPU: 8 M paths, 20 exercise dates, 6 basis functions.  The regression coefficients
    are the only quantities that depend on all paths; everything else is path-local.  So windows of paths are checked against the oracle
    started at the same path offset: (1) the regression INPUTS (the six basis functions at several exercise dates), (2) with the device's
    coefficients handed to the oracle's backward induction, the per-path Bermudan values and the exercise times.  (No discount-curve
    adjustment here: its getAverage over all paths would be a second all-path quantity.)"""
    from common import bermudan_spec
    P = 8_000_000
    s = lmm_setup(gpu)
    b = bermudan_spec(s)
    product = gpu.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
    sim = lmm_device(gpu, s, P, scheme=2, with_discount_curve=False)
    res = product.getValues(0.0, sim)
    values, exercise = res["value"].getRealizations(), res["exerciseTime"].getRealizations()
    coefficients = np.array([est.lastParameters for est in product.lastRegressions])
    conds = np.array([est.lastConditionNumber for est in product.lastRegressions])
    assert coefficients.shape == (20, 6) and np.all(np.isfinite(coefficients)) and np.all(conds > 1)
    n = 1500
    offsets = (0, 3_999_777, P - n)
    dates = (b["fixing"][0], b["fixing"][7], b["fixing"][19])
    device_basis = {}
    for date in dates:                                         # keep only the windows of the 8 M-element vectors on the host
        got = product.getBasisFunctions(date, sim)
        for k in range(1, 6):
            full = got[k].getRealizations()
            for off in offsets:
                device_basis[(date, k, off)] = full[off:off + n].copy()
        del got, full
    for off in offsets:
        ref = lmm_oracle(orc, s, n, scheme=2, with_discount_curve=False, path_offset=off)
        for date in dates:
            want = ref.bermudan_basis(date, b["fixing"], b["payment"])
            for k in range(1, 6):
                assert rel_err(device_basis[(date, k, off)], want[k]) < 1e-12, (off, date, k)
        vals, ext = ref.bermudan_given(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"], coefficients, 1.0 / P)
        same = ext == exercise[off:off + n]
        # an exercise decision may flip where the fitted trigger is zero to rounding (|trigger| ~ 1e-13 of its scale): count, do not hide
        assert np.sum(~same) <= 2, (off, int(np.sum(~same)))
        assert rel_err(values[off:off + n][same], vals[same], scale=1e-3) < 1e-11, off
    # and the price is the mean of exactly these per-path values
    assert abs(res["value"].getAverage() - float(np.mean(values))) < 1e-13
