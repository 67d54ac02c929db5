/*
 * Drives every Java_net_finmath_cuda_FinmathB200_* entry point of the JNI shim through a fake JNIEnv (arrays are malloc'ed blocks,
 * "strings" are C strings, a thrown exception is recorded as its class name).  No JVM, no JDK: compiled against tests/stubs/jni.h and
 * linked with the shim and libfinmath_b200.so by tests/test_cpu_jni.py.
 *
 *   jni_fake_env_test nodevice   every computing entry point must surface FMB_ENODEVICE as java/lang/RuntimeException (no CPU
 *                                fallback), bad arguments as IllegalArgumentException; host-only entry points must work.
 *   jni_fake_env_test gpu        (on a B200, -m gpu) a small end-to-end workflow through the shim with checked numbers.
 */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---- fake JNIEnv --------------------------------------------------------------------------------------------------------- */
struct _jobject { int kind; jsize len; void* data; char text[512]; };
enum { K_CLASS = 1, K_STRING, K_BYTES, K_INTS, K_LONGS, K_DOUBLES };
static char pending[128] = "";
static char pendingMessage[1024] = "";

static jobject mk(int kind, jsize len, size_t elem) {
	jobject o = calloc(1, sizeof(*o));
	o->kind = kind; o->len = len; o->data = len > 0 ? calloc((size_t)len, elem) : calloc(1, elem);
	return o;
}
static jclass fFindClass(JNIEnv* e, const char* name) { (void)e; jobject o = mk(K_CLASS, 0, 1); snprintf(o->text, sizeof(o->text), "%s", name); return o; }
static jint fThrowNew(JNIEnv* e, jclass c, const char* m) { (void)e; snprintf(pending, sizeof(pending), "%.100s", c->text); snprintf(pendingMessage, sizeof(pendingMessage), "%s", m ? m : ""); return 0; }
static jthrowable fExceptionOccurred(JNIEnv* e) { (void)e; return pending[0] ? (jthrowable)1 : NULL; }
static void fExceptionClear(JNIEnv* e) { (void)e; pending[0] = 0; }
static jstring fNewStringUTF(JNIEnv* e, const char* s) { (void)e; jobject o = mk(K_STRING, 0, 1); snprintf(o->text, sizeof(o->text), "%s", s); return o; }
static jsize fGetArrayLength(JNIEnv* e, jarray a) { (void)e; return a->len; }
static jbyteArray fNewByteArray(JNIEnv* e, jsize n) { (void)e; return mk(K_BYTES, n, 1); }
static jintArray fNewIntArray(JNIEnv* e, jsize n) { (void)e; return mk(K_INTS, n, 4); }
static jlongArray fNewLongArray(JNIEnv* e, jsize n) { (void)e; return mk(K_LONGS, n, 8); }
static jdoubleArray fNewDoubleArray(JNIEnv* e, jsize n) { (void)e; return mk(K_DOUBLES, n, 8); }
static jbyte* fGetBytes(JNIEnv* e, jbyteArray a, jboolean* c) { (void)e; if (c) *c = 0; return a->data; }
static jint* fGetInts(JNIEnv* e, jintArray a, jboolean* c) { (void)e; if (c) *c = 0; return a->data; }
static jlong* fGetLongs(JNIEnv* e, jlongArray a, jboolean* c) { (void)e; if (c) *c = 0; return a->data; }
static jdouble* fGetDoubles(JNIEnv* e, jdoubleArray a, jboolean* c) { (void)e; if (c) *c = 0; return a->data; }
static void fRelBytes(JNIEnv* e, jbyteArray a, jbyte* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void fRelInts(JNIEnv* e, jintArray a, jint* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void fRelLongs(JNIEnv* e, jlongArray a, jlong* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void fRelDoubles(JNIEnv* e, jdoubleArray a, jdouble* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void fSetBytes(JNIEnv* e, jbyteArray a, jsize s, jsize n, const jbyte* b) { (void)e; memcpy((jbyte*)a->data + s, b, (size_t)n); }
static void fSetLongs(JNIEnv* e, jlongArray a, jsize s, jsize n, const jlong* b) { (void)e; memcpy((jlong*)a->data + s, b, (size_t)n * 8); }
static void fSetDoubles(JNIEnv* e, jdoubleArray a, jsize s, jsize n, const jdouble* b) { (void)e; memcpy((jdouble*)a->data + s, b, (size_t)n * 8); }

static const struct JNINativeInterface_ table = {
	NULL, fFindClass, fThrowNew, fExceptionOccurred, fExceptionClear, fNewStringUTF, fGetArrayLength, fNewByteArray, fNewIntArray, fNewLongArray,
	fNewDoubleArray, fGetBytes, fGetInts, fGetLongs, fGetDoubles, fRelBytes, fRelInts, fRelLongs, fRelDoubles, fSetBytes, fSetLongs, fSetDoubles };
static JNIEnv envValue = &table;
static JNIEnv* env = &envValue;

static jdoubleArray doubles(int n, const double* v) { jdoubleArray a = fNewDoubleArray(env, n); if (v) memcpy(a->data, v, (size_t)n * 8); return a; }
static jlongArray longs(int n, const jlong* v) { jlongArray a = fNewLongArray(env, n); if (v) memcpy(a->data, v, (size_t)n * 8); return a; }
static jintArray ints(int n, const jint* v) { jintArray a = fNewIntArray(env, n); if (v) memcpy(a->data, v, (size_t)n * 4); return a; }

/* ---- the shim's entry points (prototypes as generated) -------------------------------------------------------------------- */
#define J(name) Java_net_finmath_cuda_FinmathB200_##name
void J(init)(JNIEnv*, jclass, jint); void J(shutdown)(JNIEnv*, jclass); jint J(isInitialized)(JNIEnv*, jclass); jstring J(lastError)(JNIEnv*, jclass);
jint J(deviceCount)(JNIEnv*, jclass); jstring J(deviceName)(JNIEnv*, jclass); void J(synchronize)(JNIEnv*, jclass); void J(setFpMode)(JNIEnv*, jclass, jint);
jint J(getFpMode)(JNIEnv*, jclass); void J(timerStart)(JNIEnv*, jclass); jdouble J(timerStopMs)(JNIEnv*, jclass); jlong J(kernelLaunchCount)(JNIEnv*, jclass);
jlong J(create)(JNIEnv*, jclass, jlong); jlong J(upload)(JNIEnv*, jclass, jdoubleArray); jlong J(fill)(JNIEnv*, jclass, jdouble, jlong);
jdoubleArray J(download)(JNIEnv*, jclass, jlong); jdouble J(get)(JNIEnv*, jclass, jlong, jlong); jlong J(size)(JNIEnv*, jclass, jlong);
void J(retain)(JNIEnv*, jclass, jlong); void J(free)(JNIEnv*, jclass, jlong); void J(freeMany)(JNIEnv*, jclass, jlongArray); jlong J(devicePointer)(JNIEnv*, jclass, jlong); jlongArray J(poolStats)(JNIEnv*, jclass);
void J(poolTrim)(JNIEnv*, jclass); jlong J(unary)(JNIEnv*, jclass, jint, jlong, jdouble); jlong J(binary)(JNIEnv*, jclass, jint, jlong, jdouble, jlong, jdouble);
jlong J(ternary)(JNIEnv*, jclass, jint, jlong, jdouble, jlong, jdouble, jlong, jdouble, jdouble); jlong J(evalChain)(JNIEnv*, jclass, jbyteArray, jint, jlongArray, jdoubleArray); jlong J(accrueChain)(JNIEnv*, jclass, jlongArray, jdoubleArray, jdouble);
jdoubleArray J(reduce)(JNIEnv*, jclass, jint, jlong, jlong, jdouble); jdouble J(select)(JNIEnv*, jclass, jlong, jlong); jdoubleArray J(rangeSum)(JNIEnv*, jclass, jlong, jdouble, jdouble); jlongArray J(countLessOrEqual)(JNIEnv*, jclass, jlong, jdoubleArray);
jintArray J(mtWords)(JNIEnv*, jclass, jlong, jlong, jint); jdoubleArray J(mtUniforms)(JNIEnv*, jclass, jlong, jlong, jint); jdoubleArray J(icdf)(JNIEnv*, jclass, jdoubleArray);
jlongArray J(brownianGenerate)(JNIEnv*, jclass, jint, jint, jint, jlong, jlong, jdoubleArray);
jlongArray J(uniformsGenerate)(JNIEnv*, jclass, jlong, jint, jint, jlong, jlong);
jlongArray J(eulerBlackScholes)(JNIEnv*, jclass, jint, jint, jint, jlong, jdoubleArray, jlongArray, jdouble, jdouble, jdouble);
jlongArray J(eulerHeston)(JNIEnv*, jclass, jint, jint, jint, jlong, jdoubleArray, jlongArray, jdouble, jdoubleArray, jdouble, jdouble, jdouble, jdouble, jdouble);
jlongArray J(eulerLmm)(JNIEnv*, jclass, jint, jint, jint, jdouble, jint, jint, jint, jlong, jdoubleArray, jlongArray, jdoubleArray, jdoubleArray, jdoubleArray, jdoubleArray, jintArray);
jlongArray J(eulerHullWhite)(JNIEnv*, jclass, jint, jlong, jdoubleArray, jlongArray, jdoubleArray, jdoubleArray, jdoubleArray);
jdoubleArray J(regressionMoments)(JNIEnv*, jclass, jlongArray, jdoubleArray, jlong); jdoubleArray J(solveSvd)(JNIEnv*, jclass, jint, jdoubleArray, jdoubleArray);
jlong J(regressionPredict)(JNIEnv*, jclass, jlongArray, jdoubleArray, jdoubleArray); jlong J(regressionFit)(JNIEnv*, jclass, jlongArray, jdoubleArray, jlong, jlong, jlong);
jdoubleArray J(regressionFitGet)(JNIEnv*, jclass, jlong, jint); jlong J(regressionPredictFit)(JNIEnv*, jclass, jlongArray, jdoubleArray, jlong);
jlongArray J(regressionConditionalExpectation)(JNIEnv*, jclass, jlongArray, jdoubleArray, jlong, jlong, jlong, jlongArray, jdoubleArray);
jbyteArray J(commUniqueId)(JNIEnv*, jclass); void J(commInit)(JNIEnv*, jclass, jbyteArray, jint, jint); void J(commShutdown)(JNIEnv*, jclass);
jlongArray J(commInfo)(JNIEnv*, jclass); jbyteArray J(commPeerHandle)(JNIEnv*, jclass); void J(commPeerOpen)(JNIEnv*, jclass, jbyteArray);
jdouble J(benchDfmaTflops)(JNIEnv*, jclass); jdouble J(benchCopyGbs)(JNIEnv*, jclass, jlong);

static int failures = 0, calls = 0;
#define EXPECT(cond, what) do { if (!(cond)) { printf("FAIL %s (line %d), pending exception '%s' %s\n", what, __LINE__, pending, pendingMessage); failures++; } } while (0)
/* the call must have thrown exactly this exception class */
#define THROWS(cls, call) do { pending[0] = 0; call; calls++; EXPECT(strcmp(pending, cls) == 0, #call " should throw " cls); pending[0] = 0; } while (0)
#define OK(call) do { pending[0] = 0; call; calls++; if (pending[0]) { printf("FAIL %s threw %s: %s (line %d)\n", #call, pending, pendingMessage, __LINE__); fflush(stdout); exit(2); } } while (0)
#define RTE "java/lang/RuntimeException"
#define IAE "java/lang/IllegalArgumentException"
#define UOE "java/lang/UnsupportedOperationException"

static void hostOnly(void) {
	/* entry points that need no device */
	jstring s; OK(s = J(lastError)(env, NULL)); EXPECT(s && s->kind == K_STRING, "lastError returns a string");
	jint mode = -1; OK(J(setFpMode)(env, NULL, 1)); OK(mode = J(getFpMode)(env, NULL)); EXPECT(mode == 1, "fp mode round trip"); OK(J(setFpMode)(env, NULL, 0));
	THROWS(IAE, J(setFpMode)(env, NULL, 7));
	jint n = -1; OK(n = J(deviceCount)(env, NULL)); EXPECT(n >= 0, "device count");
	jlong k = -1; OK(k = J(kernelLaunchCount)(env, NULL)); EXPECT(k >= 0, "launch count");
	const double A[4] = { 4, 1, 1, 3 }, b[2] = { 1, 2 };
	jdoubleArray x; OK(x = J(solveSvd)(env, NULL, 2, doubles(4, A), doubles(2, b)));
	EXPECT(x && x->len == 3 && fabs(((double*)x->data)[0] - 1.0 / 11) < 1e-14 && fabs(((double*)x->data)[1] - 7.0 / 11) < 1e-14, "solveSvd solves a 2x2 system");
	THROWS(IAE, J(solveSvd)(env, NULL, 3, doubles(4, A), doubles(2, b)));
	jlongArray info; OK(info = J(commInfo)(env, NULL)); EXPECT(info && info->len == 3 && ((jlong*)info->data)[1] == 1, "commInfo without a communicator: world 1");
	OK(J(commShutdown)(env, NULL));
	OK(J(poolTrim)(env, NULL));
	jlongArray st; OK(st = J(poolStats)(env, NULL)); EXPECT(st && st->len == 3, "poolStats");
	OK(J(free)(env, NULL, 0));                                       /* free(0) and free(unknown) never throw (cleaner threads) */
	OK(J(free)(env, NULL, 12345));
	{ const jlong hs[3] = { 0, 12345, 777 }; OK(J(freeMany)(env, NULL, longs(3, hs))); }
	THROWS(RTE, J(size)(env, NULL, 12345));                          /* unknown handle: FMB_EHANDLE -> RuntimeException */
	THROWS(RTE, J(retain)(env, NULL, 12345));
}

static void noDevice(void) {
	const double v[3] = { 1, 2, 3 }, sq[2] = { 1, 1 };
	const jlong hs[2] = { 1, 2 };
	const jint fl[2] = { 0, 1 };
	unsigned char code[8] = { 0 };
	jbyteArray bc = fNewByteArray(env, 8); memcpy(bc->data, code, 8);
	THROWS(RTE, J(init)(env, NULL, 0));
	THROWS(RTE, J(deviceName)(env, NULL)); THROWS(RTE, J(synchronize)(env, NULL)); THROWS(RTE, J(timerStart)(env, NULL)); THROWS(RTE, J(timerStopMs)(env, NULL));
	THROWS(RTE, J(create)(env, NULL, 10)); THROWS(RTE, J(upload)(env, NULL, doubles(3, v))); THROWS(RTE, J(fill)(env, NULL, 1.0, 10));
	THROWS(RTE, J(download)(env, NULL, 1)); THROWS(RTE, J(get)(env, NULL, 1, 0)); THROWS(RTE, J(devicePointer)(env, NULL, 1));
	THROWS(RTE, J(unary)(env, NULL, 0, 1, 0.0)); THROWS(RTE, J(binary)(env, NULL, 0, 1, 0.0, 2, 0.0)); THROWS(RTE, J(ternary)(env, NULL, 0, 1, 0, 2, 0, 3, 0, 0));
	THROWS(RTE, J(evalChain)(env, NULL, bc, 0, longs(2, hs), doubles(2, sq))); THROWS(RTE, J(accrueChain)(env, NULL, longs(2, hs), doubles(2, sq), 1.0)); THROWS(RTE, J(reduce)(env, NULL, 0, 1, 0, 0.0)); THROWS(RTE, J(select)(env, NULL, 1, 0)); THROWS(RTE, J(rangeSum)(env, NULL, 1, 0.0, 1.0));
	THROWS(RTE, J(countLessOrEqual)(env, NULL, 1, doubles(2, sq))); THROWS(RTE, J(mtWords)(env, NULL, 3141, 0, 4)); THROWS(RTE, J(mtUniforms)(env, NULL, 3141, 0, 4));
	THROWS(RTE, J(icdf)(env, NULL, doubles(3, v))); THROWS(RTE, J(brownianGenerate)(env, NULL, 3141, 2, 1, 10, 0, doubles(2, sq)));
	THROWS(RTE, J(uniformsGenerate)(env, NULL, 3141, 2, 1, 10, 0));
	THROWS(RTE, J(eulerBlackScholes)(env, NULL, 2, 2, 1, 10, doubles(2, sq), longs(2, hs), 1.0, 0.05, 0.3));
	THROWS(RTE, J(eulerHeston)(env, NULL, 2, 1, 1, 10, doubles(2, sq), longs(2, hs), 1.0, doubles(2, sq), 0.3, 0.09, 0.1, 0.5, 0.1));
	THROWS(RTE, J(eulerLmm)(env, NULL, 2, 0, 1, 1e5, 2, 1, 1, 10, doubles(2, sq), longs(2, hs), doubles(2, sq), doubles(2, sq), doubles(2, sq), doubles(2, sq), ints(2, fl)));
	THROWS(RTE, J(eulerHullWhite)(env, NULL, 1, 10, doubles(2, sq), longs(2, hs), doubles(2, sq), doubles(2, sq), doubles(4, NULL)));
	THROWS(RTE, J(regressionMoments)(env, NULL, longs(2, hs), doubles(2, sq), 3)); THROWS(RTE, J(regressionPredict)(env, NULL, longs(2, hs), doubles(2, sq), doubles(2, sq)));
	THROWS(RTE, J(regressionFit)(env, NULL, longs(2, hs), doubles(2, sq), 3, 0, 0)); THROWS(RTE, J(regressionFitGet)(env, NULL, 3, 2));
	THROWS(RTE, J(regressionPredictFit)(env, NULL, longs(2, hs), doubles(2, sq), 3));
	THROWS(RTE, J(regressionConditionalExpectation)(env, NULL, longs(2, hs), doubles(2, sq), 3, 0, 0, NULL, NULL));
	THROWS(RTE, J(commUniqueId)(env, NULL)); THROWS(RTE, J(commInit)(env, NULL, fNewByteArray(env, 128), 0, 2));
	THROWS(RTE, J(commPeerHandle)(env, NULL)); THROWS(RTE, J(commPeerOpen)(env, NULL, fNewByteArray(env, 128)));
	THROWS(RTE, J(benchDfmaTflops)(env, NULL)); THROWS(RTE, J(benchCopyGbs)(env, NULL, 1 << 20));
	jint init = 1; OK(init = J(isInitialized)(env, NULL)); EXPECT(init == 0, "not initialised without a device");
	OK(J(shutdown)(env, NULL));
}

static double* D(jdoubleArray a) { return (double*)a->data; }
static jlong* L(jlongArray a) { return (jlong*)a->data; }

static void onGpu(void) {
	OK(J(init)(env, NULL, 0));
	jint init = 0; OK(init = J(isInitialized)(env, NULL)); EXPECT(init == 1, "initialised");
	THROWS(UOE, J(commPeerHandle)(env, NULL));                      /* peer exchange needs a communicator of 2..8 ranks */
	THROWS(IAE, J(commPeerOpen)(env, NULL, fNewByteArray(env, 128)));
	jstring name; OK(name = J(deviceName)(env, NULL)); printf("device: %s\n", name ? name->text : "?");
	const double v[4] = { 1, 2, 3, 4 }, w[4] = { 0.5, 0.5, 2, 2 };
	jlong x = 0, y = 0, z = 0; OK(x = J(upload)(env, NULL, doubles(4, v))); OK(y = J(upload)(env, NULL, doubles(4, w)));
	jlong n = 0; OK(n = J(size)(env, NULL, x)); EXPECT(n == 4, "size");
	OK(z = J(unary)(env, NULL, 13 /* U_MULT */, x, 2.0));
	jdoubleArray d; OK(d = J(download)(env, NULL, z)); EXPECT(d && d->len == 4 && D(d)[0] == 2 && D(d)[3] == 8, "x * 2");
	jlong b = 0; OK(b = J(binary)(env, NULL, 0 /* B_ADD */, x, 0, y, 0)); OK(d = J(download)(env, NULL, b)); EXPECT(D(d)[2] == 5, "x + y");
	jlong t = 0; OK(t = J(ternary)(env, NULL, 4 /* T_ACCRUE */, x, 0, y, 0, 0, 0, 2.0)); OK(d = J(download)(env, NULL, t)); EXPECT(D(d)[3] == 4 * (1 + 2 * 2.0), "accrue");
	jdouble g = 0; OK(g = J(get)(env, NULL, x, 2)); EXPECT(g == 3, "get");
	jdoubleArray r; OK(r = J(reduce)(env, NULL, 0 /* R_SUM */, x, 0, 0)); EXPECT(D(r)[0] + D(r)[1] == 10, "sum");
	OK(r = J(reduce)(env, NULL, 5 /* R_MAX */, y, 0, 0)); EXPECT(D(r)[0] == 2, "max");
	jdouble sel = 0; OK(sel = J(select)(env, NULL, y, 2)); EXPECT(sel == 2.0, "third smallest of {0.5, 0.5, 2, 2}"); const double pts[2] = { 0.5, 1.0 };
	jlongArray cnt; OK(cnt = J(countLessOrEqual)(env, NULL, y, doubles(2, pts))); EXPECT(L(cnt)[0] == 2 && L(cnt)[1] == 2, "count <=");
	jdoubleArray rs; OK(rs = J(rangeSum)(env, NULL, x, 1.0, 4.0)); EXPECT(D(rs)[0] + D(rs)[1] == 5.0 && D(rs)[2] == 1 && D(rs)[3] == 3, "sum of {2, 3}, one <= 1, three < 4");
	jlong f = 0; OK(f = J(fill)(env, NULL, 7.0, 4)); OK(d = J(download)(env, NULL, f)); EXPECT(D(d)[1] == 7, "fill");
	jlong c0 = 0; OK(c0 = J(create)(env, NULL, 16)); jlong ptr = 0; OK(ptr = J(devicePointer)(env, NULL, c0)); EXPECT(ptr != 0, "device pointer");
	OK(J(retain)(env, NULL, c0)); OK(J(free)(env, NULL, c0)); OK(n = J(size)(env, NULL, c0)); EXPECT(n == 16, "retained handle survives one free"); OK(J(free)(env, NULL, c0));
	THROWS(RTE, J(size)(env, NULL, c0));
	jlong five = 0; const double v5[5] = { 1, 2, 3, 4, 5 }; OK(five = J(upload)(env, NULL, doubles(5, v5)));
	THROWS(IAE, J(binary)(env, NULL, 0, x, 0, five, 0));                 /* size mismatch */
	/* MT19937 + AS241: seed 3141 -> first tempered words / uniforms of SURVEY.md 8c */
	jintArray mw; OK(mw = J(mtWords)(env, NULL, 3141, 0, 4)); EXPECT(((uint32_t*)mw->data)[0] == 2066431349u && ((uint32_t*)mw->data)[3] == 2939336508u, "MT words");
	jdoubleArray mu; OK(mu = J(mtUniforms)(env, NULL, 3141, 0, 2)); EXPECT(D(mu)[0] == 0.48112854170930563 && D(mu)[1] == 0.8554104072506368, "MT uniforms");
	jdoubleArray ic; OK(ic = J(icdf)(env, NULL, mu)); EXPECT(fabs(D(ic)[0] + 0.047321386241461816) < 1e-15, "AS241");
	/* Brownian motion -> Black-Scholes -> mean of S(T) ~ exp(r T) */
	const int T = 4, P = 20000;
	double dt[4], sq[4]; for (int i = 0; i < T; i++) { dt[i] = 0.25; sq[i] = 0.5; }
	jlongArray dW; OK(dW = J(brownianGenerate)(env, NULL, 3141, T, 1, P, 0, doubles(T, sq))); EXPECT(dW && dW->len == T, "T*F increments");
	jlongArray X; OK(X = J(eulerBlackScholes)(env, NULL, 2, T, 1, P, doubles(T, dt), dW, 1.0, 0.05, 0.2)); EXPECT(X && X->len == T + 1 && L(X)[0] == 0 && L(X)[T] != 0, "process handles");
	OK(r = J(reduce)(env, NULL, 0, L(X)[T], 0, 0)); EXPECT(fabs((D(r)[0] + D(r)[1]) / P - exp(0.05)) < 0.01, "E[S(1)] = exp(r)");
	{ jlongArray U; OK(U = J(uniformsGenerate)(env, NULL, 3141, 1, 2, 4, 0)); jdoubleArray u0; OK(u0 = J(download)(env, NULL, L(U)[0]));
	  EXPECT(u0 && D(u0)[0] == 0.48112854170930563 && D(u0)[1] == 0.7730656542235694, "uniforms in draw order (path-major)");
	  jlong z2 = 0; OK(z2 = J(unary)(env, NULL, 19 /* U_ICDF_NORMAL */, L(U)[0], 0.0)); OK(u0 = J(download)(env, NULL, z2)); EXPECT(fabs(D(u0)[0] + 0.047321386241461816) < 1e-15, "ICDF op"); }
	/* Heston / Hull-White / LMM through the shim: shapes only (numerics are covered by the Python-driven parity tests) */
	jlongArray dW2; OK(dW2 = J(brownianGenerate)(env, NULL, 31415, T, 2, P, 0, doubles(T, sq)));
	double rate[4] = { 0.05, 0.05, 0.05, 0.05 };
	jlongArray H; OK(H = J(eulerHeston)(env, NULL, 2, 1, T, P, doubles(T, dt), dW2, 1.0, doubles(T, rate), 0.3, 0.09, 0.1, 0.5, 0.1)); EXPECT(H && H->len == (T + 1) * 2, "Heston handles");
	double d0[4] = { -0.1, -0.1, -0.1, -0.1 }, d1[4] = { 1, 1, 1, 1 }, flh[16]; for (int i = 0; i < 16; i++) flh[i] = (i % 4 == 0 || i % 4 == 3) ? 0.01 : 0.0;
	jlongArray W; OK(W = J(eulerHullWhite)(env, NULL, T, P, doubles(T, dt), dW2, doubles(T, d0), doubles(T, d1), doubles(16, flh))); EXPECT(W && W->len == (T + 1) * 2, "Hull-White handles");
	{
		const int N = 3; double y0[3], pl[3] = { 0.25, 0.25, 0.25 }, fll[4 * 3 * 2], var[4 * 3]; jint first[4] = { 1, 2, 3, 3 };
		for (int j = 0; j < N; j++) y0[j] = log(0.05);
		for (int i = 0; i < 24; i++) fll[i] = 0.1;
		for (int i = 0; i < 12; i++) var[i] = 0.02;
		jlongArray Lm; OK(Lm = J(eulerLmm)(env, NULL, 2, 0, 1, 1e5, T, N, 2, P, doubles(T, dt), dW2, doubles(N, y0), doubles(N, pl), doubles(24, fll), doubles(12, var), ints(T, first)));
		EXPECT(Lm && Lm->len == (T + 1) * N && L(Lm)[1 * N + 0] == 0 && L(Lm)[1 * N + 1] != 0 && L(Lm)[2 * N + 1] == L(Lm)[1 * N + 1], "LMM handles alias frozen rates");
	}
	/* regression: y = 2 + 3 b exactly */
	double bv[4] = { 0.1, 0.4, 0.7, 1.3 }, yv[4]; for (int i = 0; i < 4; i++) yv[i] = 2 + 3 * bv[i];
	jlong hb = 0, hy = 0; OK(hb = J(upload)(env, NULL, doubles(4, bv))); OK(hy = J(upload)(env, NULL, doubles(4, yv)));
	jlong basis[2] = { 0, hb }; double bs[2] = { 1.0, 0.0 };
	jdoubleArray mom; OK(mom = J(regressionMoments)(env, NULL, longs(2, basis), doubles(2, bs), hy)); EXPECT(mom && mom->len == 12 && D(mom)[0] == 4.0, "moments: sum 1*1 = n");
	jlongArray ce; OK(ce = J(regressionConditionalExpectation)(env, NULL, longs(2, basis), doubles(2, bs), hy, 0, 0, NULL, NULL)); EXPECT(ce && ce->len == 2, "fit + ce handles");
	jdoubleArray fit; OK(fit = J(regressionFitGet)(env, NULL, L(ce)[0], 2)); EXPECT(fit && fabs(D(fit)[4 + 2] - 2.0) < 1e-9 && fabs(D(fit)[4 + 3] - 3.0) < 1e-9, "coefficients 2, 3");
	OK(d = J(download)(env, NULL, L(ce)[1])); EXPECT(fabs(D(d)[3] - yv[3]) < 1e-9, "fitted value");
	jlong f2 = 0; OK(f2 = J(regressionFit)(env, NULL, longs(2, basis), doubles(2, bs), hy, 4, L(ce)[0])); jlong p2 = 0; OK(p2 = J(regressionPredictFit)(env, NULL, longs(2, basis), doubles(2, bs), f2));
	OK(d = J(download)(env, NULL, p2)); EXPECT(fabs(D(d)[0] - yv[0]) < 1e-9, "predict from a cached fit");
	double xs[2] = { 2, 3 }; jlong p3 = 0; OK(p3 = J(regressionPredict)(env, NULL, longs(2, basis), doubles(2, bs), doubles(2, xs)));
	OK(d = J(download)(env, NULL, p3)); EXPECT(fabs(D(d)[1] - yv[1]) < 1e-12, "predict from host coefficients");
	/* chain: (x * 2) + y in one pass */
	unsigned char code[16] = { 0, 13, 0, 128, 0, 0, 0, 0,   1, 0, 0, 1, 0, 0, 0, 0 };
	jbyteArray bc = fNewByteArray(env, 16); memcpy(bc->data, code, 16);
	jlong leaves[2]; leaves[0] = x; leaves[1] = y; double sc[1] = { 2.0 };
	jlong ch = 0; OK(ch = J(evalChain)(env, NULL, bc, 0, longs(2, leaves), doubles(1, sc))); OK(d = J(download)(env, NULL, ch)); EXPECT(D(d)[2] == 3 * 2 + 2.0, "chain");
	{ jlong two[2]; two[0] = x; two[1] = y; const double dl[2] = { 0.5, 0.5 }; jlong ac = 0; OK(ac = J(accrueChain)(env, NULL, longs(2, two), doubles(2, dl), 1.0));
	  OK(d = J(download)(env, NULL, ac)); EXPECT(D(d)[0] == ((1 * 0.5 + 1.0) * (1 + 0.5 * 0.5) - 1.0) / 1.0, "accrue chain"); }
	OK(J(timerStart)(env, NULL)); jdouble ms = -1; OK(ms = J(timerStopMs)(env, NULL)); EXPECT(ms >= 0, "timer");
	jdouble tf = 0; OK(tf = J(benchDfmaTflops)(env, NULL)); EXPECT(tf > 1, "DFMA peak"); jdouble gb = 0; OK(gb = J(benchCopyGbs)(env, NULL, 1 << 26)); EXPECT(gb > 100, "copy bandwidth");
	OK(J(synchronize)(env, NULL));
	OK(J(shutdown)(env, NULL));
}

int main(int argc, char** argv) {
	setvbuf(stdout, NULL, _IONBF, 0);
	const char* mode = argc > 1 ? argv[1] : "nodevice";
	hostOnly();
	if (strcmp(mode, "gpu") == 0) onGpu(); else noDevice();
	printf("%s: %d shim calls, %d failures\n", mode, calls, failures);
	return failures ? 1 : 0;
}
