package net.finmath.cuda;

import net.finmath.stochastic.ConditionalExpectationEstimator;
import net.finmath.stochastic.RandomVariable;

/**
 * Drop-in for MonteCarloConditionalExpectationRegression (J/montecarlo/conditionalexpectation/MonteCarloConditionalExpectationRegression.java:33-180):
 * XtX and Xty in ONE fused pass (fmb_regression_moments), K x K pseudo-inverse solve on the host with the commons-math3 cut-off
 * (fmb_regression_solve_svd), prediction b_0 x_0 + sum b_i x_i in one kernel (fmb_regression_predict).  Executable twin: montecarlo.py.
 * Use by overriding BermudanSwaption.getConditionalExpectationEstimator (public, BermudanSwaption.java:182).
 * NOT COMPILED IN THIS REPOSITORY'S CI (no JDK in the image).
 */
public class MonteCarloConditionalExpectationRegressionCuda implements ConditionalExpectationEstimator {
	private final RandomVariable[] basisFunctions;
	private double[] XTX;        // cached like the reference's solver (:125-138)

	public MonteCarloConditionalExpectationRegressionCuda(final RandomVariable[] basisFunctions) {
		this.basisFunctions = java.util.Arrays.stream(basisFunctions).filter(b -> b != null).toArray(RandomVariable[]::new);   // :79-95
	}

	public double[] getLinearRegressionParameters(final RandomVariable dependents) {
		final int K = basisFunctions.length;
		final long[] handles = new long[K];
		final double[] scalars = new double[K];
		for (int i = 0; i < K; i++) {
			final RandomVariable b = basisFunctions[i];
			if (b.isDeterministic()) scalars[i] = b.doubleValue();
			else handles[i] = (b instanceof RandomVariableCuda) ? ((RandomVariableCuda) b).handle() : FinmathB200.upload(b.getRealizations());
		}
		final RandomVariableCuda y = (RandomVariableCuda) new RandomVariableCuda(0.0, 0.0).add(dependents);      // lands on the GPU type
		final double[] moments = FinmathB200.regressionMoments(handles, scalars, y.handle());
		final double n = dependents.size();
		final double[] xtx = new double[K * K], xty = new double[K];
		for (int i = 0; i < K * K; i++) xtx[i] = moments[i] / n;
		for (int i = 0; i < K; i++) xty[i] = moments[K * K + i] / n;
		for (int i = 0; i < K; i++) for (int j = 0; j < K; j++)
			if (handles[i] == 0 && handles[j] == 0) xtx[i * K + j] = scalars[i] * scalars[j];
		if (XTX == null) XTX = xtx;
		return FinmathB200.solveSvd(K, XTX, xty);
	}

	@Override
	public RandomVariable getConditionalExpectation(final RandomVariable randomVariable) {
		final double[] x = getLinearRegressionParameters(randomVariable);
		final int K = basisFunctions.length;
		final long[] handles = new long[K];
		final double[] scalars = new double[K];
		double time = Double.NEGATIVE_INFINITY;
		for (int i = 0; i < K; i++) {
			final RandomVariable b = basisFunctions[i];
			time = Math.max(time, b.getFiltrationTime());
			if (b.isDeterministic()) scalars[i] = b.doubleValue();
			else handles[i] = ((RandomVariableCuda) b).handle();
		}
		return new RandomVariableCuda(time, FinmathB200.regressionPredict(handles, scalars, x), randomVariable.size());
	}
}
