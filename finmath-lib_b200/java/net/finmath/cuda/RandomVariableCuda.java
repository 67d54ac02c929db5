package net.finmath.cuda;

import java.lang.ref.Cleaner;
import java.util.function.DoubleBinaryOperator;
import java.util.function.DoubleUnaryOperator;
import java.util.function.IntToDoubleFunction;
import java.util.stream.DoubleStream;

import net.finmath.functions.DoubleTernaryOperator;
import net.finmath.stochastic.ConditionalExpectationEstimator;
import net.finmath.stochastic.RandomVariable;

import static net.finmath.cuda.FinmathB200.*;

/**
 * Device-resident RandomVariable: type priority 2 (wins against Scalar = 0 and RandomVariableFromDoubleArray = 1, stays below the
 * AAD wrapper = 3, RandomVariable.java:45-51).  Deterministic values are host doubles, stochastic values a handle to a
 * double[paths] on the GPU; every operation is one kernel behind the C ABI (include/finmath_b200.h).  Branch structure and
 * rounding order follow RandomVariableFromDoubleArray line by line; finmath-lib_b200/stochastic.py is the executable twin of
 * this file (same structure, tested against the oracle on a B200).
 * NOT COMPILED IN THIS REPOSITORY'S CI (no JDK in the image).
 */
public class RandomVariableCuda implements RandomVariable {
	private static final long serialVersionUID = 1L;
	private static final Cleaner CLEANER = Cleaner.create();

	private final double time;
	private final double valueIfNonStochastic;
	private final long handle;      // 0 = deterministic
	private final long size;

	public RandomVariableCuda(final double time, final double value) {
		this.time = time; this.valueIfNonStochastic = value; this.handle = 0; this.size = 1;
	}
	RandomVariableCuda(final double time, final long handle, final long size) {
		this.time = time; this.valueIfNonStochastic = Double.NaN; this.handle = handle; this.size = size;
		final long h = handle;
		CLEANER.register(this, () -> FinmathB200.free(h));        // device memory returns to the pool when the wrapper is collected
	}

	private RandomVariableCuda vec(final double t, final long h) { return new RandomVariableCuda(t, h, FinmathB200.size(h)); }
	private static long handleOf(final RandomVariable rv) {
		if (rv instanceof RandomVariableCuda) return ((RandomVariableCuda) rv).handle;
		return rv.isDeterministic() ? 0 : FinmathB200.upload(rv.getRealizations());
	}
	private static double scalarOf(final RandomVariable rv) { return rv.isDeterministic() ? rv.doubleValue() : Double.NaN; }
	private RandomVariable map(final int op, final double a, final DoubleUnaryOperator host) {
		if (handle == 0) return new RandomVariableCuda(time, host.applyAsDouble(valueIfNonStochastic));
		return vec(time, unary(op, handle, a));
	}

	@Override public boolean equals(final RandomVariable rv) {
		if (time != rv.getFiltrationTime()) return false;
		if (isDeterministic() && rv.isDeterministic()) return valueIfNonStochastic == rv.doubleValue();
		if (isDeterministic() != rv.isDeterministic()) return false;
		return java.util.Arrays.equals(getRealizations(), rv.getRealizations());
	}
	@Override public double getFiltrationTime() { return time; }
	@Override public int getTypePriority() { return 2; }
	@Override public double get(final int i) { return handle == 0 ? valueIfNonStochastic : FinmathB200.get(handle, i); }
	@Override public int size() { return (int) size; }
	@Override public boolean isDeterministic() { return handle == 0; }
	@Override public double[] getRealizations() { return handle == 0 ? new double[] { valueIfNonStochastic } : download(handle); }
	@Override public Double doubleValue() {
		if (handle == 0) return valueIfNonStochastic;
		if (size == 1) return getAverage();
		throw new UnsupportedOperationException("The random variable is non-deterministic");
	}
	@Override public IntToDoubleFunction getOperator() { throw new UnsupportedOperationException("device-resident values"); }
	@Override public DoubleStream getRealizationsStream() { throw new UnsupportedOperationException("device-resident values"); }
	@Override public RandomVariable cache() { return this; }
	@Override public RandomVariable apply(final DoubleUnaryOperator o) { throw new UnsupportedOperationException("lambdas cannot run on the device"); }
	@Override public RandomVariable apply(final DoubleBinaryOperator o, final RandomVariable a) { throw new UnsupportedOperationException("lambdas cannot run on the device"); }
	@Override public RandomVariable apply(final DoubleTernaryOperator o, final RandomVariable a, final RandomVariable b) { throw new UnsupportedOperationException("lambdas cannot run on the device"); }

	// ---- reductions (double-double sums on the device, RandomVariableFromDoubleArray.java:262-428)
	/** hi + lo of the double-double sum (all shards when the library holds a communicator); min / max in [0]. */
	private static double reduced(final int op, final long x, final long w, final double a) { final double[] r = reduce(op, x, w, a); return r[0] + r[1]; }
	@Override public double getMin() { return handle == 0 ? valueIfNonStochastic : reduce(R_MIN, handle, 0, 0)[0]; }
	@Override public double getMax() { return handle == 0 ? valueIfNonStochastic : reduce(R_MAX, handle, 0, 0)[0]; }
	@Override public double getAverage() {
		if (handle == 0) return valueIfNonStochastic;
		if (size == 0) return Double.NaN;
		return reduced(R_SUM, handle, 0, 0) / size;
	}
	@Override public double getAverage(final RandomVariable p) {
		if (handle == 0) return valueIfNonStochastic * p.getAverage();
		if (size == 0) return Double.NaN;
		if (p.isDeterministic()) return mult(p.doubleValue()).getAverage();
		return reduced(R_SUM_PRODUCT, handle, handleOf(p), 0) / size;
	}
	@Override public double getVariance() {
		if (handle == 0 || size == 1) return 0.0;
		if (size == 0) return Double.NaN;
		return reduced(R_CENTERED_M2, handle, 0, getAverage()) / size;
	}
	@Override public double getVariance(final RandomVariable p) {          // not divided by n (:379)
		if (handle == 0) return 0.0;
		if (size == 0) return Double.NaN;
		final double average = getAverage(p);
		if (p.isDeterministic()) return ((RandomVariableCuda) sub(average).squared().mult(p.doubleValue())).sum();
		return reduced(R_CENTERED_M2_W, handle, handleOf(p), average);
	}
	private double sum() { return reduced(R_SUM, handle, 0, 0); }
	@Override public double getSampleVariance() { return (handle == 0 || size == 1) ? 0.0 : getVariance() * size / (size - 1); }
	@Override public double getStandardDeviation() { return handle == 0 ? 0.0 : Math.sqrt(getVariance()); }
	@Override public double getStandardDeviation(final RandomVariable p) { return handle == 0 ? 0.0 : Math.sqrt(getVariance(p)); }
	@Override public double getStandardError() { return handle == 0 ? 0.0 : getStandardDeviation() / Math.sqrt(size); }
	@Override public double getStandardError(final RandomVariable p) { return handle == 0 ? 0.0 : getStandardDeviation(p) / Math.sqrt(size); }
	// ---- order statistics without a sort (radix select / counting pass / range sum on the device; :445-575)
	private long quantileIndex(final double q) { return Math.min(Math.max(Math.round((size + 1) * q - 1), 0), size - 1); }      // :454-459
	@Override public double getQuantile(final double q) {
		if (handle == 0) return valueIfNonStochastic;
		if (size == 0) return Double.NaN;
		return select(handle, quantileIndex(q));
	}
	@Override public double getQuantile(final double q, final RandomVariable p) { throw new RuntimeException("Method not implemented."); }   // :471
	@Override public double getQuantileExpectation(final double a, final double b) {
		if (handle == 0) return valueIfNonStochastic;
		if (size == 0) return Double.NaN;
		if (a > b) return getQuantileExpectation(b, a);
		final long i0 = quantileIndex(a), i1 = quantileIndex(b);
		final double v0 = select(handle, i0), v1 = select(handle, i1);
		if (v0 == v1) return v0;
		final double[] r = rangeSum(handle, v0, v1);        // {sum hi, sum lo, #(x <= v0), #(x < v1)}
		return (r[0] + r[1] + ((long) r[2] - i0) * v0 + (i1 - (long) r[3] + 1) * v1) / (i1 - i0 + 1);
	}
	@Override public double[] getHistogram(final double[] pts) {
		final double[] h = new double[pts.length + 1];
		if (handle == 0) {                                      // :505-517
			for (int k = 0; k < pts.length; k++) if (valueIfNonStochastic > pts[k]) { h[k] = 1.0; break; }
			h[pts.length] = 1.0;
			return h;
		}
		final long[] counts = countLessOrEqual(handle, pts);    // (at most 511 thresholds per call: chunk longer arrays)
		long prev = 0;
		for (int k = 0; k < pts.length; k++) { final long c = Math.max(counts[k], prev); h[k] = c - prev; prev = c; }   // :528-550
		h[pts.length] = size - prev;
		if (size > 0) for (int k = 0; k < h.length; k++) h[k] /= size;
		return h;
	}
	@Override public double[][] getHistogram(final int n, final double sd) {                                                     // :553-575
		final double[] pts = new double[n], anchors = new double[n + 1];
		final double center = getAverage(), radius = sd * getStandardDeviation(), stepSize = (n - 1) / 2.0;
		for (int i = 0; i < n; i++) {
			final double alpha = (-(double) (n - 1) / 2.0 + i) / stepSize;
			pts[i] = center + alpha * radius;
			anchors[i] = center + alpha * radius - radius / (2 * stepSize);
		}
		anchors[n] = center + 1 * radius + radius / (2 * stepSize);
		return new double[][] { anchors, getHistogram(pts) };
	}

	// ---- unary and rv op double (:742-1020)
	@Override public RandomVariable cap(final double c) { return map(U_CAP, c, x -> Math.min(x, c)); }
	@Override public RandomVariable floor(final double f) { return map(U_FLOOR, f, x -> Math.max(x, f)); }
	@Override public RandomVariable add(final double v) { return map(U_ADD, v, x -> x + v); }
	@Override public RandomVariable sub(final double v) { return map(U_SUB, v, x -> x - v); }
	@Override public RandomVariable bus(final double v) { return map(U_BUS, v, x -> v - x); }
	@Override public RandomVariable mult(final double v) { return map(U_MULT, v, x -> x * v); }
	@Override public RandomVariable div(final double v) { return map(U_DIV, v, x -> x / v); }
	@Override public RandomVariable vid(final double v) { return map(U_VID, v, x -> v / x); }
	@Override public RandomVariable pow(final double e) { return map(U_POW, e, x -> Math.pow(x, e)); }
	@Override public RandomVariable average() { return new RandomVariableCuda(Double.NEGATIVE_INFINITY, getAverage()); }
	@Override public RandomVariable getConditionalExpectation(final ConditionalExpectationEstimator e) { return e.getConditionalExpectation(this); }
	@Override public RandomVariable squared() { return map(U_SQUARED, 0, x -> x * x); }
	@Override public RandomVariable sqrt() { return map(U_SQRT, 0, Math::sqrt); }
	@Override public RandomVariable exp() { return map(U_EXP, 0, Math::exp); }
	@Override public RandomVariable expm1() { return map(U_EXPM1, 0, Math::expm1); }
	@Override public RandomVariable log() { return map(U_LOG, 0, Math::log); }
	@Override public RandomVariable sin() { return map(U_SIN, 0, Math::sin); }
	@Override public RandomVariable cos() { return map(U_COS, 0, Math::cos); }
	@Override public RandomVariable invert() { return map(U_INVERT, 0, x -> 1.0 / x); }
	@Override public RandomVariable abs() { return map(U_ABS, 0, Math::abs); }
	@Override public RandomVariable isNaN() { return map(U_ISNAN, 0, x -> Double.isNaN(x) ? 1.0 : 0.0); }

	// ---- binary (:1027-1276): the higher type priority handles the operation; time = max; deterministic shortcuts as in the reference
	private RandomVariable bin(final int op, final RandomVariable rv, final DoubleBinaryOperator host, final int shortcut) {
		final double t = Math.max(time, rv.getFiltrationTime());
		if (handle == 0 && rv.isDeterministic()) return new RandomVariableCuda(t, host.applyAsDouble(valueIfNonStochastic, rv.doubleValue()));
		if (rv.isDeterministic() && shortcut >= 0 && handle != 0) return vec(time, unary(shortcut, handle, rv.doubleValue()));
		return vec(t, binary(op, handle, valueIfNonStochastic, handleOf(rv), scalarOf(rv)));
	}
	@Override public RandomVariable add(final RandomVariable rv) { return rv.getTypePriority() > 2 ? rv.add(this) : bin(B_ADD, rv, (a, b) -> a + b, U_ADD); }
	@Override public RandomVariable sub(final RandomVariable rv) { return rv.getTypePriority() > 2 ? rv.bus(this) : bin(B_SUB, rv, (a, b) -> a - b, U_SUB); }
	@Override public RandomVariable bus(final RandomVariable rv) {
		if (rv.getTypePriority() > 2) return rv.sub(this);
		final double t = Math.max(time, rv.getFiltrationTime());
		if (handle == 0 && rv.isDeterministic()) return new RandomVariableCuda(t, rv.doubleValue() - valueIfNonStochastic);
		return vec(t, binary(B_SUB, handleOf(rv), scalarOf(rv), handle, valueIfNonStochastic));
	}
	@Override public RandomVariable mult(final RandomVariable rv) { return rv.getTypePriority() > 2 ? rv.mult(this) : bin(B_MULT, rv, (a, b) -> a * b, U_MULT); }
	@Override public RandomVariable div(final RandomVariable rv) { return rv.getTypePriority() > 2 ? rv.vid(this) : bin(B_DIV, rv, (a, b) -> a / b, -1); }
	@Override public RandomVariable vid(final RandomVariable rv) {
		if (rv.getTypePriority() > 2) return rv.div(this);
		final double t = Math.max(time, rv.getFiltrationTime());
		if (handle == 0 && rv.isDeterministic()) return new RandomVariableCuda(t, rv.doubleValue() / valueIfNonStochastic);
		return vec(t, binary(B_DIV, handleOf(rv), scalarOf(rv), handle, valueIfNonStochastic));
	}
	@Override public RandomVariable cap(final RandomVariable rv) { return rv.getTypePriority() > 2 ? rv.cap(this) : bin(B_CAP, rv, Math::min, -1); }
	@Override public RandomVariable floor(final RandomVariable rv) { return rv.getTypePriority() > 2 ? rv.floor(this) : bin(B_FLOOR, rv, Math::max, U_FLOOR); }

	// ---- ternary (:1278-1479)
	@Override public RandomVariable accrue(final RandomVariable rate, final double pl) {
		if (rate.getTypePriority() > 2) return rate.mult(pl).add(1.0).mult(this);
		if (rate.isDeterministic()) return mult(1.0 + rate.doubleValue() * pl);
		return vec(Math.max(time, rate.getFiltrationTime()), ternary(T_ACCRUE, handle, valueIfNonStochastic, handleOf(rate), 0, 0, 0, pl));
	}
	@Override public RandomVariable discount(final RandomVariable rate, final double pl) {
		if (rate.getTypePriority() > 2) return rate.mult(pl).add(1.0).invert().mult(this);
		if (rate.isDeterministic()) return div(1.0 + rate.doubleValue() * pl);
		return vec(Math.max(time, rate.getFiltrationTime()), ternary(T_DISCOUNT, handle, valueIfNonStochastic, handleOf(rate), 0, 0, 0, pl));
	}
	@Override public RandomVariable choose(final RandomVariable nonNeg, final RandomVariable neg) {
		if (handle == 0) return valueIfNonStochastic >= 0 ? nonNeg : neg;
		final double t = Math.max(Math.max(time, nonNeg.getFiltrationTime()), neg.getFiltrationTime());
		return vec(t, ternary(T_CHOOSE, handle, 0, handleOf(nonNeg), scalarOf(nonNeg), handleOf(neg), scalarOf(neg), 0));
	}
	@Override public RandomVariable addProduct(final RandomVariable f1, final double f2) {
		if (f1.getTypePriority() > 2) return f1.mult(f2).add(this);
		if (f1.isDeterministic()) return add(f1.doubleValue() * f2);
		return vec(Math.max(time, f1.getFiltrationTime()), ternary(T_ADD_PRODUCT_D, handle, valueIfNonStochastic, handleOf(f1), 0, 0, 0, f2));
	}
	@Override public RandomVariable addProduct(final RandomVariable f1, final RandomVariable f2) {
		if (f1.getTypePriority() > 2 || f2.getTypePriority() > 2) return f1.mult(f2).add(this);
		final double t = Math.max(Math.max(time, f1.getFiltrationTime()), f2.getFiltrationTime());
		final boolean d1 = f1.isDeterministic(), d2 = f2.isDeterministic();
		if (handle == 0 && d1 && d2) return new RandomVariableCuda(t, valueIfNonStochastic + (f1.doubleValue() * f2.doubleValue()));
		if (d1 && d2) return add(f1.doubleValue() * f2.doubleValue());
		if (d2) return addProduct(f1, f2.doubleValue());
		if (d1) return addProduct(f2, f1.doubleValue());
		if (handle != 0) return vec(t, ternary(T_ADD_PRODUCT, handle, 0, handleOf(f1), 0, handleOf(f2), 0, 0));
		return add(f1.mult(f2));
	}
	@Override public RandomVariable addRatio(final RandomVariable n, final RandomVariable d) {
		if (n.getTypePriority() > 2 || d.getTypePriority() > 2) return n.div(d).add(this);
		final double t = Math.max(Math.max(time, n.getFiltrationTime()), d.getFiltrationTime());
		if (handle == 0 && n.isDeterministic() && d.isDeterministic()) return new RandomVariableCuda(t, valueIfNonStochastic + (n.doubleValue() / d.doubleValue()));
		return vec(t, ternary(T_ADD_RATIO, handle, valueIfNonStochastic, handleOf(n), scalarOf(n), handleOf(d), scalarOf(d), 0));
	}
	@Override public RandomVariable subRatio(final RandomVariable n, final RandomVariable d) {
		if (n.getTypePriority() > 2 || d.getTypePriority() > 2) return n.div(d).mult(-1).add(this);
		final double t = Math.max(Math.max(time, n.getFiltrationTime()), d.getFiltrationTime());
		if (handle == 0 && n.isDeterministic() && d.isDeterministic()) return new RandomVariableCuda(t, valueIfNonStochastic - (n.doubleValue() / d.doubleValue()));
		return vec(t, ternary(T_SUB_RATIO, handle, valueIfNonStochastic, handleOf(n), scalarOf(n), handleOf(d), scalarOf(d), 0));
	}

	long handle() { return handle; }
}
