"""Shared set-up helpers for the parity tests: the same synthetic market data on the oracle side and on the device side.

LMM set-up after T/montecarlo/interestrate/LIBORMarketModelValuationTest.java:91-162 (40 semi-annual forwards at 5 %,
vol (a,b,c,d) = (0.2, 0, 0.25, 0.3), correlation decay 0.1, seed 3141, spot measure, log-normal).
"""
import numpy as np


def rel_err(a, b, scale=1.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b) / den
    d = np.where((a == b) | (np.isnan(a) & np.isnan(b)), 0.0, d)
    return float(np.max(d)) if d.size else 0.0


def same_bits(a, b):
    """Bit-identical, except that any NaN matches any NaN (payloads are not portable between CPU and GPU)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return bool(np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64)))


def lmm_setup(pkg, n_libors=40, n_factors=3, period=0.5, dt=0.5, horizon=None, a=0.2, b=0.0, c=0.25, d=0.3, decay=0.1, forward=0.05):
    """Returns dict with host-side tables shared by both sides."""
    horizon = n_libors * period if horizon is None else horizon
    tenor = pkg.TimeDiscretizationFromArray(0.0, n_libors, period)
    sim = pkg.TimeDiscretizationFromArray(0.0, int(round(horizon / dt)), dt)
    vol = pkg.LIBORVolatilityModelFourParameterExponentialForm(sim, tenor, a, b, c, d, False)
    corr = pkg.LIBORCorrelationModelExponentialDecay(sim, tenor, n_factors, decay)
    cov = pkg.LIBORCovarianceModelFromVolatilityAndCorrelation(sim, tenor, vol, corr)
    T, N = sim.getNumberOfTimeSteps(), tenor.getNumberOfTimeSteps()
    sigma = np.array([[vol.getVolatility(t, j) for j in range(N)] for t in range(T)])
    L0 = np.full(N, forward)
    df = np.ones(N + 1)
    for i in range(N):                                       # DiscountCurveFromForwardCurve.getDiscountFactor :130-142
        df[i + 1] = df[i] / (1.0 + L0[i] * tenor.getTimeStep(i))
    return dict(tenor=tenor, sim=sim, cov=cov, corr=corr, sigma=sigma, factor_matrix=corr.factorMatrix.copy(), L0=L0, df=df, F=n_factors, N=N, T=T)


def lmm_device(pkg, s, paths, seed=3141, scheme=None, measure="SPOT", state_space="LOGNORMAL", with_discount_curve=True, shard=None,
               force_generic=False, libor_cap=1e5):
    factory = pkg.RandomVariableCudaFactory(shard)
    props = {"measure": measure, "stateSpace": state_space, "liborCap": libor_cap}
    model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"] if with_discount_curve else None, factory, s["cov"], None, props)
    bm = pkg.BrownianMotionCuda(s["sim"], s["F"], paths, seed, factory)
    process = pkg.EulerSchemeFromProcessModel(model, bm, scheme, forceGeneric=force_generic)
    return pkg.LIBORMonteCarloSimulationFromLIBORModel(process)


def lmm_oracle(orc, s, paths, seed=3141, scheme=2, measure=0, state_space=1, with_discount_curve=True, path_offset=0, libor_cap=1e5):
    return orc.LMM(seed, s["sim"].times, s["tenor"].times, s["F"], paths, s["L0"], s["sigma"], s["factor_matrix"],
                   discount_factors=s["df"] if with_discount_curve else None, measure=measure, state_space=state_space, libor_cap=libor_cap,
                   scheme=scheme, path_offset=path_offset)


def device_process_array(sim_model, T, N):
    """[T+1][N][P] realizations (deterministic entries broadcast)."""
    proc = sim_model.getProcess()
    P = proc.getNumberOfPaths()
    out = np.empty((T + 1, N, P))
    for t in range(T + 1):
        for j in range(N):
            rv = proc.getProcessValue(t, j)
            out[t, j] = rv.doubleValue() if rv.isDeterministic() else rv.getRealizations()
    return out


def bermudan_spec(s, first_period=10, n_periods=20, strike=0.05):
    """C5: exercise at 5y, 20 semi-annual periods, all period starts exercisable."""
    tenor = s["tenor"]
    fixing = [tenor.getTime(first_period + i) for i in range(n_periods)]
    payment = [tenor.getTime(first_period + i + 1) for i in range(n_periods)]
    lengths = [p - f for f, p in zip(fixing, payment)]
    return dict(is_exercise=[True] * n_periods, fixing=fixing, lengths=lengths, payment=payment, notionals=[1.0] * n_periods,
                swaprates=[strike] * n_periods)
