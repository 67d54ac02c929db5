package net.finmath.cuda;

import net.finmath.montecarlo.AbstractRandomVariableFactory;
import net.finmath.stochastic.RandomVariable;

/**
 * The seam: hand this factory to BrownianMotionCuda and to the model constructors
 * (LIBORMarketModelFromCovarianceModel.of(..., randomVariableFactory, ...), new BlackScholesModel(..., factory), ...),
 * exactly where the reference's tests inject their factories (T/montecarlo/interestrate/LIBORMarketModelValuationTest.java:64-73).
 * NOT COMPILED IN THIS REPOSITORY'S CI (no JDK in the image).
 */
public class RandomVariableCudaFactory extends AbstractRandomVariableFactory {
	private static final long serialVersionUID = 1L;

	@Override
	public RandomVariable createRandomVariable(final double time, final double value) {
		return new RandomVariableCuda(time, value);          // deterministic values stay host scalars
	}

	@Override
	public RandomVariable createRandomVariable(final double time, final double[] values) {
		return new RandomVariableCuda(time, FinmathB200.upload(values), values.length);
	}
}
