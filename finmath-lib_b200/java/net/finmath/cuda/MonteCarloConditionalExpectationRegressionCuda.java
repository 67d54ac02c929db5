package net.finmath.cuda;

import net.finmath.stochastic.ConditionalExpectationEstimator;
import net.finmath.stochastic.RandomVariable;

/**
 * Drop-in for MonteCarloConditionalExpectationRegression (J/montecarlo/conditionalexpectation/MonteCarloConditionalExpectationRegression.java:33-180).
 * getConditionalExpectation is ONE native call that returns at once (fmb_regression_conditional_expectation): XtX and Xty in one fused pass,
 * the shards' moments exchanged on the device when the library holds a communicator, the K x K pseudo-inverse solve (commons-math3 cut-off)
 * and the prediction b_0 x_0 + sum b_i x_i all queued on the compute stream - a Bermudan backward induction never waits for an exercise
 * date.  The coefficients are downloaded only by getLinearRegressionParameters.  Executable twin: montecarlo.py.
 * Use by overriding BermudanSwaption.getConditionalExpectationEstimator (public, BermudanSwaption.java:182).
 * NOT COMPILED IN THIS REPOSITORY'S CI (no JDK in the image).
 */
public class MonteCarloConditionalExpectationRegressionCuda implements ConditionalExpectationEstimator {
	private final RandomVariable[] basisFunctions;
	private long cachedFit;      // device-resident first fit: its XtX is reused, like the reference's cached solver (:125-138)
	private long lastFit;

	public MonteCarloConditionalExpectationRegressionCuda(final RandomVariable[] basisFunctions) {
		this.basisFunctions = java.util.Arrays.stream(basisFunctions).filter(b -> b != null).toArray(RandomVariable[]::new);   // :79-95
	}

	private long[] handles;
	private double[] scalars;
	private double time;
	private void collectBasis() {
		final int K = basisFunctions.length;
		handles = new long[K];
		scalars = new double[K];
		time = Double.NEGATIVE_INFINITY;
		for (int i = 0; i < K; i++) {
			final RandomVariable b = basisFunctions[i];
			time = Math.max(time, b.getFiltrationTime());
			if (b.isDeterministic()) scalars[i] = b.doubleValue();
			else handles[i] = (b instanceof RandomVariableCuda) ? ((RandomVariableCuda) b).handle() : FinmathB200.upload(b.getRealizations());
		}
	}
	private void remember(final long fit) {
		if (lastFit != 0 && lastFit != cachedFit) FinmathB200.free(lastFit);
		lastFit = fit;
		if (cachedFit == 0) cachedFit = fit;
	}

	public double[] getLinearRegressionParameters(final RandomVariable dependents) {
		collectBasis();
		final RandomVariableCuda y = (RandomVariableCuda) new RandomVariableCuda(0.0, 0.0).add(dependents);      // lands on the GPU type
		remember(FinmathB200.regressionFit(handles, scalars, y.handle(), dependents.size(), cachedFit));
		final int K = basisFunctions.length;
		final double[] all = FinmathB200.regressionFitGet(lastFit, K);        // XtX[K*K], Xty[K], x[K], cond
		return java.util.Arrays.copyOfRange(all, K * K + K, K * K + 2 * K);
	}

	@Override
	public RandomVariable getConditionalExpectation(final RandomVariable randomVariable) {
		collectBasis();
		final RandomVariableCuda y = (RandomVariableCuda) new RandomVariableCuda(0.0, 0.0).add(randomVariable);
		final long[] fitAndResult = FinmathB200.regressionConditionalExpectation(handles, scalars, y.handle(), randomVariable.size(), cachedFit, null, null);
		remember(fitAndResult[0]);
		return new RandomVariableCuda(time, fitAndResult[1], randomVariable.size());
	}
}
