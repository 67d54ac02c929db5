package net.finmath.cuda;

import net.finmath.montecarlo.BrownianMotion;
import net.finmath.montecarlo.RandomVariableFactory;
import net.finmath.stochastic.RandomVariable;
import net.finmath.time.TimeDiscretization;

/**
 * Drop-in for BrownianMotionFromMersenneRandomNumbers (J/montecarlo/BrownianMotionFromMersenneRandomNumbers.java:41-258): same
 * constructor arguments, same lazy generation, same draw order and seeding — the increments are produced on the device by
 * fmb_bm_generate (counter-addressable MT19937 + AS241) and never visit the host.  finmath-lib_b200/montecarlo.py is the executable twin.
 * NOT COMPILED IN THIS REPOSITORY'S CI (no JDK in the image).
 */
public class BrownianMotionCuda implements BrownianMotion {
	private static final long serialVersionUID = 1L;
	private final TimeDiscretization timeDiscretization;
	private final int numberOfFactors, numberOfPaths, seed;
	private final RandomVariableFactory randomVariableFactory = new RandomVariableCudaFactory();
	private transient RandomVariable[][] brownianIncrements;
	private transient long[] handles;
	private final Object lock = new Object();

	public BrownianMotionCuda(final TimeDiscretization timeDiscretization, final int numberOfFactors, final int numberOfPaths, final int seed) {
		this.timeDiscretization = timeDiscretization; this.numberOfFactors = numberOfFactors; this.numberOfPaths = numberOfPaths; this.seed = seed;
	}

	private void generate() {
		final int T = timeDiscretization.getNumberOfTimeSteps();
		final double[] sqrtDt = new double[T];
		for (int t = 0; t < T; t++) sqrtDt[t] = Math.sqrt(timeDiscretization.getTimeStep(t));
		handles = FinmathB200.brownianGenerate(seed, T, numberOfFactors, numberOfPaths, 0L, sqrtDt);
		brownianIncrements = new RandomVariable[T][numberOfFactors];
		for (int t = 0; t < T; t++)
			for (int f = 0; f < numberOfFactors; f++)
				brownianIncrements[t][f] = new RandomVariableCuda(timeDiscretization.getTime(t + 1), handles[t * numberOfFactors + f], numberOfPaths);
	}

	@Override public RandomVariable getBrownianIncrement(final int timeIndex, final int factor) {
		synchronized (lock) { if (brownianIncrements == null) generate(); }
		return brownianIncrements[timeIndex][factor];
	}
	@Override public RandomVariable getIncrement(final int timeIndex, final int factor) { return getBrownianIncrement(timeIndex, factor); }
	long[] getIncrementHandles() { synchronized (lock) { if (brownianIncrements == null) generate(); } return handles; }
	@Override public TimeDiscretization getTimeDiscretization() { return timeDiscretization; }
	@Override public int getNumberOfFactors() { return numberOfFactors; }
	@Override public int getNumberOfPaths() { return numberOfPaths; }
	@Override public RandomVariable getRandomVariableForConstant(final double value) { return randomVariableFactory.createRandomVariable(value); }
	@Override public BrownianMotion getCloneWithModifiedSeed(final int seed) { return new BrownianMotionCuda(timeDiscretization, numberOfFactors, numberOfPaths, seed); }
	@Override public BrownianMotion getCloneWithModifiedTimeDiscretization(final TimeDiscretization td) { return new BrownianMotionCuda(td, numberOfFactors, numberOfPaths, seed); }
}
