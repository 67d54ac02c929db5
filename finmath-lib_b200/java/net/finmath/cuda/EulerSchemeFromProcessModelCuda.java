package net.finmath.cuda;

import java.util.Map;

import net.finmath.montecarlo.IndependentIncrements;
import net.finmath.montecarlo.assetderivativevaluation.models.BlackScholesModel;
import net.finmath.montecarlo.interestrate.models.LIBORMarketModelFromCovarianceModel;
import net.finmath.montecarlo.model.ProcessModel;
import net.finmath.montecarlo.process.EulerSchemeFromProcessModel;
import net.finmath.montecarlo.process.EulerSchemeFromProcessModel.Scheme;
import net.finmath.montecarlo.process.MonteCarloProcess;
import net.finmath.montecarlo.process.MonteCarloProcessFromProcessModel;
import net.finmath.stochastic.RandomVariable;
import net.finmath.stochastic.Scalar;
import net.finmath.time.TimeDiscretization;

/**
 * Drop-in for EulerSchemeFromProcessModel (J/montecarlo/process/EulerSchemeFromProcessModel.java:60-403).  For the models of the hot path
 * the whole time loop is ONE fused kernel (fmb_euler_*); every other ProcessModel is delegated to the reference's own Euler scheme,
 * which then runs op by op on RandomVariableCuda (still on the device).  Executable twin: finmath-lib_b200/montecarlo.py.
 * NOT COMPILED IN THIS REPOSITORY'S CI (no JDK in the image).
 */
public class EulerSchemeFromProcessModelCuda extends MonteCarloProcessFromProcessModel {
	private final BrownianMotionCuda driver;
	private final Scheme scheme;
	private transient RandomVariable[][] discreteProcess;
	private transient MonteCarloProcess fallback;

	public EulerSchemeFromProcessModelCuda(final ProcessModel model, final BrownianMotionCuda driver, final Scheme scheme) {
		super(driver.getTimeDiscretization(), model);
		this.driver = driver; this.scheme = scheme;
	}

	@Override public synchronized RandomVariable getProcessValue(final int timeIndex, final int componentIndex) {
		if (discreteProcess == null && fallback == null) precalculate();
		return fallback != null ? fallback.getProcessValue(timeIndex, componentIndex) : discreteProcess[timeIndex][componentIndex];
	}
	@Override public RandomVariable getMonteCarloWeights(final int timeIndex) { return driver.getRandomVariableForConstant(1.0 / getNumberOfPaths()); }
	@Override public int getNumberOfPaths() { return driver.getNumberOfPaths(); }
	@Override public int getNumberOfFactors() { return driver.getNumberOfFactors(); }
	@Override public IndependentIncrements getStochasticDriver() { return driver; }

	private void precalculate() {
		final TimeDiscretization td = getTimeDiscretization();
		final int T = td.getNumberOfTimeSteps(), N = getNumberOfComponents(), F = driver.getNumberOfFactors();
		final double[] dt = new double[T];
		for (int t = 0; t < T; t++) dt[t] = td.getTime(t + 1) - td.getTime(t);
		final ProcessModel model = getModel();
		long[] out;
		double[] initialValues;
		if (model instanceof LIBORMarketModelFromCovarianceModel) {
			final LIBORMarketModelFromCovarianceModel lmm = (LIBORMarketModelFromCovarianceModel) model;
			// deterministic tables, evaluated through the model's own callbacks on Scalars (the reference's arithmetic, once, on the host)
			final double[] y0 = new double[N], pl = new double[N], fl = new double[T * N * F], var = new double[T * N];
			final int[] first = new int[T];
			final RandomVariable[] initialState = lmm.getInitialState(this);
			for (int j = 0; j < N; j++) { y0[j] = initialState[j].doubleValue(); pl[j] = lmm.getLiborPeriodDiscretization().getTimeStep(j); }
			for (int t = 0; t < T; t++) {
				int f = lmm.getLiborPeriodIndex(td.getTime(t)) + 1;
				if (f < 0) f = -f - 1 + 1;
				first[t] = f;
				for (int j = 0; j < N; j++) {
					final RandomVariable[] l = lmm.getFactorLoading(this, t, j, null);
					for (int k = 0; k < F; k++) fl[(t * N + j) * F + k] = l[k].doubleValue();
					var[t * N + j] = lmm.getCovarianceModel().getCovariance(td.getTime(t), j, j, null).doubleValue();
				}
			}
			out = FinmathB200.eulerLmm(scheme.ordinal() /* enum order == C-ABI codes (:68-73) */,
					lmm.getMeasure().ordinal(), lmm.getStateSpace().ordinal(), lmm.getLiborCap(), T, N, F, driver.getNumberOfPaths(), dt, driver.getIncrementHandles(),
					y0, pl, fl, var, first);
			initialValues = new double[N];
			for (int j = 0; j < N; j++) initialValues[j] = lmm.applyStateSpaceTransform(this, 0, j, new Scalar(y0[j])).doubleValue();
		}
		else if (model instanceof BlackScholesModel) {
			final BlackScholesModel bs = (BlackScholesModel) model;
			out = FinmathB200.eulerBlackScholes(scheme.ordinal(), T, F, driver.getNumberOfPaths(), dt, driver.getIncrementHandles(),
					bs.getInitialValue().doubleValue(), bs.getRiskFreeRate().doubleValue(), bs.getVolatility().doubleValue());
			initialValues = new double[] { Math.exp(Math.log(bs.getInitialValue().doubleValue())) };
		}
		else {
			fallback = new EulerSchemeFromProcessModel(model, driver, scheme);      // op-by-op on RandomVariableCuda
			return;
		}
		discreteProcess = new RandomVariable[T + 1][N];
		for (int t = 0; t <= T; t++) {
			for (int c = 0; c < N; c++) {
				final long h = out[t * N + c];
				if (h == 0) discreteProcess[t][c] = driver.getRandomVariableForConstant(initialValues[c]);
				else if (t > 0 && h == out[(t - 1) * N + c]) { FinmathB200.free(h); discreteProcess[t][c] = discreteProcess[t - 1][c]; }   // frozen component: alias (:285)
				else discreteProcess[t][c] = new RandomVariableCuda(td.getTime(t), h, driver.getNumberOfPaths());
			}
		}
	}

	@Override public Object getCloneWithModifiedSeed(final int seed) { return new EulerSchemeFromProcessModelCuda(getModel(), (BrownianMotionCuda) driver.getCloneWithModifiedSeed(seed), scheme); }
	@Override public MonteCarloProcess getCloneWithModifiedModel(final ProcessModel model) { return new EulerSchemeFromProcessModelCuda(model, driver, scheme); }
	@Override public MonteCarloProcess getCloneWithModifiedData(final Map<String, Object> dataModified) { throw new UnsupportedOperationException("clone with modified data: construct a new process"); }
	@Override public EulerSchemeFromProcessModelCuda clone() { return new EulerSchemeFromProcessModelCuda(getModel(), driver, scheme); }
}
