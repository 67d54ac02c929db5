/*
 * JNI shim: net.finmath.cuda.FinmathB200 (static native methods) -> the C ABI of include/finmath_b200.h.
 * Purely mechanical: argument marshalling and error translation, no logic.  Built only where a JDK is present
 * (cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux finmath_b200_jni.c -L.. -lfinmath_b200 -o libfinmath_b200_jni.so);
 * this image has no jni.h, so the file is not part of the default build.
 */
#include <jni.h>
#include <stdint.h>
#include "../../../include/finmath_b200.h"

static void throwFor(JNIEnv* env, int rc) {
	const char* cls = "java/lang/RuntimeException";
	if (rc == FMB_EINVAL) cls = "java/lang/IllegalArgumentException";
	else if (rc == FMB_EUNSUPPORTED) cls = "java/lang/UnsupportedOperationException";
	else if (rc == FMB_ENOMEM) cls = "java/lang/OutOfMemoryError";
	(*env)->ThrowNew(env, (*env)->FindClass(env, cls), fmb_last_error());
}
#define CHECK(expr) do { int rc__ = (expr); if (rc__ != FMB_OK) { throwFor(env, rc__); } } while (0)

JNIEXPORT void JNICALL Java_net_finmath_cuda_FinmathB200_init(JNIEnv* env, jclass c, jint device) { CHECK(fmb_init(device)); }

JNIEXPORT jlong JNICALL Java_net_finmath_cuda_FinmathB200_upload(JNIEnv* env, jclass c, jdoubleArray values) {
	const jsize n = (*env)->GetArrayLength(env, values);
	jdouble* p = (*env)->GetPrimitiveArrayCritical(env, values, NULL);
	fmb_handle h = 0;
	const int rc = fmb_rv_upload(p, (uint64_t)n, &h);
	(*env)->ReleasePrimitiveArrayCritical(env, values, p, JNI_ABORT);
	if (rc != FMB_OK) throwFor(env, rc);
	return (jlong)h;
}
JNIEXPORT jdoubleArray JNICALL Java_net_finmath_cuda_FinmathB200_download(JNIEnv* env, jclass c, jlong h) {
	uint64_t n = 0;
	CHECK(fmb_rv_size((fmb_handle)h, &n));
	jdoubleArray out = (*env)->NewDoubleArray(env, (jsize)n);
	jdouble* p = (*env)->GetPrimitiveArrayCritical(env, out, NULL);
	const int rc = fmb_rv_download((fmb_handle)h, p, n);
	(*env)->ReleasePrimitiveArrayCritical(env, out, p, 0);
	if (rc != FMB_OK) throwFor(env, rc);
	return out;
}
JNIEXPORT jdouble JNICALL Java_net_finmath_cuda_FinmathB200_get(JNIEnv* env, jclass c, jlong h, jlong i) {
	double v = 0; CHECK(fmb_rv_get((fmb_handle)h, (uint64_t)i, &v)); return v;
}
JNIEXPORT jlong JNICALL Java_net_finmath_cuda_FinmathB200_size(JNIEnv* env, jclass c, jlong h) {
	uint64_t n = 0; CHECK(fmb_rv_size((fmb_handle)h, &n)); return (jlong)n;
}
JNIEXPORT void JNICALL Java_net_finmath_cuda_FinmathB200_free(JNIEnv* env, jclass c, jlong h) { fmb_rv_free((fmb_handle)h); }

JNIEXPORT jlong JNICALL Java_net_finmath_cuda_FinmathB200_unary(JNIEnv* env, jclass c, jint op, jlong x, jdouble a) {
	fmb_handle out = 0; CHECK(fmb_rv_unary(op, (fmb_handle)x, a, &out)); return (jlong)out;
}
JNIEXPORT jlong JNICALL Java_net_finmath_cuda_FinmathB200_binary(JNIEnv* env, jclass c, jint op, jlong x, jdouble sx, jlong y, jdouble sy) {
	fmb_handle out = 0; CHECK(fmb_rv_binary(op, (fmb_handle)x, sx, (fmb_handle)y, sy, &out)); return (jlong)out;
}
JNIEXPORT jlong JNICALL Java_net_finmath_cuda_FinmathB200_ternary(JNIEnv* env, jclass c, jint op, jlong x, jdouble sx, jlong y, jdouble sy,
		jlong z, jdouble sz, jdouble a) {
	fmb_handle out = 0; CHECK(fmb_rv_ternary(op, (fmb_handle)x, sx, (fmb_handle)y, sy, (fmb_handle)z, sz, a, &out)); return (jlong)out;
}
/* a chain of element-wise operations in one pass (a RandomVariableCuda that defers evaluation builds code / leaves / scalars) */
JNIEXPORT jlong JNICALL Java_net_finmath_cuda_FinmathB200_evalChain(JNIEnv* env, jclass c, jbyteArray code, jint startLeaf, jlongArray leaves,
		jdoubleArray scalars) {
	const jsize nCode = (*env)->GetArrayLength(env, code), nLeaves = (*env)->GetArrayLength(env, leaves);
	const jsize nScalars = (*env)->GetArrayLength(env, scalars);
	jbyte* pc = (*env)->GetByteArrayElements(env, code, NULL);
	jlong* pl = (*env)->GetLongArrayElements(env, leaves, NULL);
	jdouble* ps = (*env)->GetDoubleArrayElements(env, scalars, NULL);
	fmb_handle out = 0;
	const int rc = fmb_rv_eval_chain(nCode / 8, (const unsigned char*)pc, startLeaf, (const fmb_handle*)pl, nLeaves, ps, nScalars, &out);
	(*env)->ReleaseByteArrayElements(env, code, pc, JNI_ABORT);
	(*env)->ReleaseLongArrayElements(env, leaves, pl, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, scalars, ps, JNI_ABORT);
	CHECK(rc);
	return (jlong)out;
}
/* returns hi + lo of the double-double sum (single-GPU JVM); min / max in [0] */
JNIEXPORT jdouble JNICALL Java_net_finmath_cuda_FinmathB200_reduce(JNIEnv* env, jclass c, jint op, jlong x, jlong w, jdouble a) {
	double out[2] = {0, 0}; CHECK(fmb_rv_reduce(op, (fmb_handle)x, (fmb_handle)w, a, out)); return out[0] + out[1];
}

JNIEXPORT jlongArray JNICALL Java_net_finmath_cuda_FinmathB200_brownianGenerate(JNIEnv* env, jclass c, jint seed, jint T, jint F, jlong paths,
		jlong pathOffset, jdoubleArray sqrtDt) {
	jlongArray out = (*env)->NewLongArray(env, T * F);
	jdouble* sq = (*env)->GetDoubleArrayElements(env, sqrtDt, NULL);
	jlong* h = (*env)->GetLongArrayElements(env, out, NULL);
	const int rc = fmb_bm_generate(seed, T, F, (uint64_t)paths, (uint64_t)pathOffset, sq, (fmb_handle*)h);
	(*env)->ReleaseLongArrayElements(env, out, h, 0);
	(*env)->ReleaseDoubleArrayElements(env, sqrtDt, sq, JNI_ABORT);
	if (rc != FMB_OK) throwFor(env, rc);
	return out;
}

JNIEXPORT jlongArray JNICALL Java_net_finmath_cuda_FinmathB200_eulerLmm(JNIEnv* env, jclass c, jint scheme, jint measure, jint stateSpace,
		jdouble liborCap, jint T, jint N, jint F, jlong paths, jdoubleArray dt, jlongArray dW, jdoubleArray initialState, jdoubleArray periodLength,
		jdoubleArray factorLoading, jdoubleArray variance, jintArray firstLive) {
	jlongArray out = (*env)->NewLongArray(env, (T + 1) * N);
	jdouble* pdt = (*env)->GetDoubleArrayElements(env, dt, NULL);
	jlong* pdw = (*env)->GetLongArrayElements(env, dW, NULL);
	jdouble* py0 = (*env)->GetDoubleArrayElements(env, initialState, NULL);
	jdouble* ppl = (*env)->GetDoubleArrayElements(env, periodLength, NULL);
	jdouble* pfl = (*env)->GetDoubleArrayElements(env, factorLoading, NULL);
	jdouble* pva = (*env)->GetDoubleArrayElements(env, variance, NULL);
	jint* pfi = (*env)->GetIntArrayElements(env, firstLive, NULL);
	jlong* h = (*env)->GetLongArrayElements(env, out, NULL);
	const int rc = fmb_euler_lmm(scheme, measure, stateSpace, liborCap, T, N, F, (uint64_t)paths, pdt, (const fmb_handle*)pdw, py0, ppl, pfl, pva,
	                             (const int32_t*)pfi, (fmb_handle*)h);
	(*env)->ReleaseLongArrayElements(env, out, h, 0);
	(*env)->ReleaseIntArrayElements(env, firstLive, pfi, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, variance, pva, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, factorLoading, pfl, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, periodLength, ppl, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, initialState, py0, JNI_ABORT);
	(*env)->ReleaseLongArrayElements(env, dW, pdw, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, dt, pdt, JNI_ABORT);
	if (rc != FMB_OK) throwFor(env, rc);
	return out;
}

JNIEXPORT jlongArray JNICALL Java_net_finmath_cuda_FinmathB200_eulerBlackScholes(JNIEnv* env, jclass c, jint scheme, jint T, jint F, jlong paths,
		jdoubleArray dt, jlongArray dW, jdouble initialValue, jdouble riskFreeRate, jdouble volatility) {
	jlongArray out = (*env)->NewLongArray(env, T + 1);
	jdouble* pdt = (*env)->GetDoubleArrayElements(env, dt, NULL);
	jlong* pdw = (*env)->GetLongArrayElements(env, dW, NULL);
	jlong* h = (*env)->GetLongArrayElements(env, out, NULL);
	const int rc = fmb_euler_black_scholes(scheme, T, F, (uint64_t)paths, pdt, (const fmb_handle*)pdw, initialValue, riskFreeRate, volatility, (fmb_handle*)h);
	(*env)->ReleaseLongArrayElements(env, out, h, 0);
	(*env)->ReleaseLongArrayElements(env, dW, pdw, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, dt, pdt, JNI_ABORT);
	if (rc != FMB_OK) throwFor(env, rc);
	return out;
}

/* moments[0 .. K*K) = XtX sums (hi+lo), moments[K*K .. K*K+K) = Xty sums */
JNIEXPORT jdoubleArray JNICALL Java_net_finmath_cuda_FinmathB200_regressionMoments(JNIEnv* env, jclass c, jlongArray basis, jdoubleArray basisScalar, jlong y) {
	const jsize K = (*env)->GetArrayLength(env, basis);
	double xh[64], xl[64], yh[8], yl[8];
	jlong* pb = (*env)->GetLongArrayElements(env, basis, NULL);
	jdouble* ps = (*env)->GetDoubleArrayElements(env, basisScalar, NULL);
	const int rc = K <= 8 ? fmb_regression_moments(K, (const fmb_handle*)pb, ps, (fmb_handle)y, xh, xl, yh, yl) : FMB_EUNSUPPORTED;
	(*env)->ReleaseDoubleArrayElements(env, basisScalar, ps, JNI_ABORT);
	(*env)->ReleaseLongArrayElements(env, basis, pb, JNI_ABORT);
	if (rc != FMB_OK) { throwFor(env, rc); return NULL; }
	jdoubleArray out = (*env)->NewDoubleArray(env, K * K + K);
	jdouble* po = (*env)->GetDoubleArrayElements(env, out, NULL);
	for (int i = 0; i < K * K; i++) po[i] = xh[i] + xl[i];
	for (int i = 0; i < K; i++) po[K * K + i] = yh[i] + yl[i];
	(*env)->ReleaseDoubleArrayElements(env, out, po, 0);
	return out;
}
JNIEXPORT jdoubleArray JNICALL Java_net_finmath_cuda_FinmathB200_solveSvd(JNIEnv* env, jclass c, jint K, jdoubleArray A, jdoubleArray b) {
	jdoubleArray out = (*env)->NewDoubleArray(env, K);
	jdouble* pa = (*env)->GetDoubleArrayElements(env, A, NULL);
	jdouble* pb = (*env)->GetDoubleArrayElements(env, b, NULL);
	jdouble* po = (*env)->GetDoubleArrayElements(env, out, NULL);
	const int rc = fmb_regression_solve_svd(K, pa, pb, po, NULL);
	(*env)->ReleaseDoubleArrayElements(env, out, po, 0);
	(*env)->ReleaseDoubleArrayElements(env, b, pb, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, A, pa, JNI_ABORT);
	if (rc != FMB_OK) throwFor(env, rc);
	return out;
}
JNIEXPORT jlong JNICALL Java_net_finmath_cuda_FinmathB200_regressionPredict(JNIEnv* env, jclass c, jlongArray basis, jdoubleArray basisScalar, jdoubleArray x) {
	const jsize K = (*env)->GetArrayLength(env, basis);
	jlong* pb = (*env)->GetLongArrayElements(env, basis, NULL);
	jdouble* ps = (*env)->GetDoubleArrayElements(env, basisScalar, NULL);
	jdouble* px = (*env)->GetDoubleArrayElements(env, x, NULL);
	fmb_handle out = 0;
	const int rc = fmb_regression_predict(K, (const fmb_handle*)pb, ps, px, &out);
	(*env)->ReleaseDoubleArrayElements(env, x, px, JNI_ABORT);
	(*env)->ReleaseDoubleArrayElements(env, basisScalar, ps, JNI_ABORT);
	(*env)->ReleaseLongArrayElements(env, basis, pb, JNI_ABORT);
	if (rc != FMB_OK) throwFor(env, rc);
	return (jlong)out;
}
