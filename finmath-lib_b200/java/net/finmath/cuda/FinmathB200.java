package net.finmath.cuda;

/**
 * Static native entry points — one per function of include/finmath_b200.h (JNI shim: csrc/jni/finmath_b200_jni.c).
 * Handles are opaque 64-bit ids of device-resident double vectors; 0 means "no vector, use the scalar next to it".
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no JDK (see INTEGRATION.md).
 */
public final class FinmathB200 {
	static {
		System.loadLibrary("finmath_b200");      // the CUDA library (C ABI)
		System.loadLibrary("finmath_b200_jni");  // the shim
		init(Integer.getInteger("net.finmath.cuda.device", 0));
	}
	private FinmathB200() {}

	// op codes of include/finmath_b200.h
	public static final int U_SQUARED = 0, U_SQRT = 1, U_EXP = 2, U_LOG = 3, U_SIN = 4, U_COS = 5, U_INVERT = 6, U_ABS = 7, U_ISNAN = 8, U_EXPM1 = 9,
			U_ADD = 10, U_SUB = 11, U_BUS = 12, U_MULT = 13, U_DIV = 14, U_VID = 15, U_CAP = 16, U_FLOOR = 17, U_POW = 18;
	public static final int B_ADD = 0, B_SUB = 1, B_MULT = 2, B_DIV = 3, B_CAP = 4, B_FLOOR = 5;
	public static final int T_ADD_PRODUCT = 0, T_ADD_PRODUCT_D = 1, T_ADD_RATIO = 2, T_SUB_RATIO = 3, T_ACCRUE = 4, T_DISCOUNT = 5, T_CHOOSE = 6;
	public static final int R_SUM = 0, R_SUM_PRODUCT = 1, R_CENTERED_M2 = 2, R_CENTERED_M2_W = 3, R_MIN = 4, R_MAX = 5;

	public static native void init(int device);
	public static native long upload(double[] values);
	public static native double[] download(long handle);
	public static native double get(long handle, long index);
	public static native long size(long handle);
	public static native void free(long handle);
	public static native long unary(int op, long x, double a);
	public static native long binary(int op, long x, double sx, long y, double sy);
	public static native long ternary(int op, long x, double sx, long y, double sy, long z, double sz, double a);
	/** A chain of element-wise operations in one pass: code = 8 bytes per instruction (include/finmath_b200.h, fmb_rv_eval_chain). */
	public static native long evalChain(byte[] code, int startLeaf, long[] leaves, double[] scalars);
	public static native double reduce(int op, long x, long w, double a);
	public static native long[] brownianGenerate(int seed, int numberOfTimeSteps, int numberOfFactors, long paths, long pathOffset, double[] sqrtDt);
	public static native long[] eulerLmm(int scheme, int measure, int stateSpace, double liborCap, int T, int N, int F, long paths, double[] dt, long[] dW,
			double[] initialState, double[] periodLength, double[] factorLoading, double[] variance, int[] firstLive);
	public static native long[] eulerBlackScholes(int scheme, int T, int F, long paths, double[] dt, long[] dW, double initialValue, double riskFreeRate, double volatility);
	public static native double[] regressionMoments(long[] basis, double[] basisScalar, long y);
	public static native double[] solveSvd(int K, double[] A, double[] b);
	public static native long regressionPredict(long[] basis, double[] basisScalar, double[] x);
}
