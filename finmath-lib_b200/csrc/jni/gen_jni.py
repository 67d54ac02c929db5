#!/usr/bin/env python
"""Generator of the JNI shim (finmath_b200_jni.c) and of the Java class holding the native methods (FinmathB200.java).

One table, one line per function of include/finmath_b200.h: Java name, return kind, arguments, the C call.  The shim is purely
mechanical (argument marshalling + error translation), so it is generated - and a test (tests/test_cpu_jni.py) checks that every
function the header declares has a row here, compiles the result against a JNI header and drives every Java_* entry point through a
fake JNIEnv.  Run this script after changing the table; both outputs are committed.

Argument kinds:  i jint   l jlong (uint64_t / int64_t)   d jdouble   h jlong (fmb_handle)
                 D double[] in   H long[] of handles in   I int[] in   B byte[] in       (arrays may be null where the C ABI allows NULL)
Return kinds:    void | int (out int) | long (out uint64) | double | handle (trailing fmb_handle* out) | handles:<count> (trailing fmb_handle* array)
                 | doubles:<count> (trailing double* array) | custom (hand-written body below)
"""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))

# (java name, C function, return kind, [(kind, name)...], C argument list with OUT standing for the generated output pointer)
TABLE = [
    ("init", "fmb_init", "void", [("i", "device")], "device"),
    ("shutdown", "fmb_shutdown", "void", [], ""),
    ("isInitialized", "fmb_is_initialized", "rawint", [], ""),
    ("lastError", "fmb_last_error", "custom", [], ""),
    ("deviceCount", "fmb_device_count", "int", [], "OUT"),
    ("deviceName", "fmb_device_name", "custom", [], ""),
    ("synchronize", "fmb_synchronize", "void", [], ""),
    ("setFpMode", "fmb_set_fp_mode", "void", [("i", "mode")], "mode"),
    ("getFpMode", "fmb_get_fp_mode", "int", [], "OUT"),
    ("timerStart", "fmb_timer_start", "void", [], ""),
    ("timerStopMs", "fmb_timer_stop_ms", "custom", [], ""),
    ("kernelLaunchCount", "fmb_kernel_launch_count", "long", [], "OUT"),
    ("create", "fmb_rv_create", "handle", [("l", "n")], "(uint64_t)n, OUT"),
    ("upload", "fmb_rv_upload", "custom", [("D", "values")], ""),
    ("fill", "fmb_rv_fill", "handle", [("d", "value"), ("l", "n")], "value, (uint64_t)n, OUT"),
    ("download", "fmb_rv_download", "custom", [("h", "handle")], ""),
    ("get", "fmb_rv_get", "double", [("h", "handle"), ("l", "index")], "(fmb_handle)handle, (uint64_t)index, OUT"),
    ("size", "fmb_rv_size", "long", [("h", "handle")], "(fmb_handle)handle, OUT"),
    ("retain", "fmb_rv_retain", "void", [("h", "handle")], "(fmb_handle)handle"),
    ("free", "fmb_rv_free", "nothrow", [("h", "handle")], "(fmb_handle)handle"),
    ("freeMany", "fmb_rv_free_many", "void", [("H", "handles")], "(const fmb_handle*)handles_p, (uint64_t)handles_n"),
    ("devicePointer", "fmb_rv_device_ptr", "custom", [("h", "handle")], ""),
    ("poolStats", "fmb_pool_stats", "custom", [], ""),
    ("poolTrim", "fmb_pool_trim", "void", [], ""),
    ("unary", "fmb_rv_unary", "handle", [("i", "op"), ("h", "x"), ("d", "a")], "op, (fmb_handle)x, a, OUT"),
    ("binary", "fmb_rv_binary", "handle", [("i", "op"), ("h", "x"), ("d", "sx"), ("h", "y"), ("d", "sy")], "op, (fmb_handle)x, sx, (fmb_handle)y, sy, OUT"),
    ("ternary", "fmb_rv_ternary", "handle", [("i", "op"), ("h", "x"), ("d", "sx"), ("h", "y"), ("d", "sy"), ("h", "z"), ("d", "sz"), ("d", "a")],
     "op, (fmb_handle)x, sx, (fmb_handle)y, sy, (fmb_handle)z, sz, a, OUT"),
    ("accrueChain", "fmb_rv_accrue_chain", "handle", [("H", "rates"), ("D", "periodLengths"), ("d", "divisor")],
     "rates_n, (const fmb_handle*)rates_p, periodLengths_p, divisor, OUT"),
    ("accruePrefix", "fmb_rv_accrue_prefix", "handles:rates_n", [("d", "start"), ("H", "rates"), ("D", "periodLengths")],
     "rates_n, start, (const fmb_handle*)rates_p, periodLengths_p, OUT"),
    ("evalChain", "fmb_rv_eval_chain", "handle", [("B", "code"), ("i", "startLeaf"), ("H", "leaves"), ("D", "scalars")],
     "code_n / 8, (const unsigned char*)code_p, startLeaf, (const fmb_handle*)leaves_p, leaves_n, scalars_p, scalars_n, OUT"),
    ("reduce", "fmb_rv_reduce", "doubles:2", [("i", "op"), ("h", "x"), ("h", "w"), ("d", "a")], "op, (fmb_handle)x, (fmb_handle)w, a, OUT"),
    ("reduceMany", "fmb_rv_reduce_many", "doubles:2 * x_n", [("i", "op"), ("H", "x"), ("d", "a")], "op, x_n, (const fmb_handle*)x_p, a, OUT"),
    ("select", "fmb_rv_select", "double", [("h", "x"), ("l", "rank")], "(fmb_handle)x, (uint64_t)rank, OUT"),
    ("countLessOrEqual", "fmb_rv_count_le", "custom", [("h", "x"), ("D", "points")], ""),
    ("rangeSum", "fmb_rv_range_sum", "doubles:4", [("h", "x"), ("d", "lo"), ("d", "hi")], "(fmb_handle)x, lo, hi, OUT"),
    ("mtWords", "fmb_mt_words", "custom", [("l", "seed"), ("l", "wordOffset"), ("i", "n")], ""),
    ("mtUniforms", "fmb_mt_uniforms", "doubles:n", [("l", "seed"), ("l", "uniformOffset"), ("i", "n")], "(int64_t)seed, (uint64_t)uniformOffset, (uint64_t)n, OUT"),
    ("icdf", "fmb_icdf", "doubles:p_n", [("D", "p")], "p_p, (uint64_t)p_n, OUT"),
    ("brownianGenerate", "fmb_bm_generate", "handles:T * F", [("i", "seed"), ("i", "T"), ("i", "F"), ("l", "paths"), ("l", "pathOffset"), ("D", "sqrtDt")],
     "seed, T, F, (uint64_t)paths, (uint64_t)pathOffset, sqrtDt_p, OUT"),
    ("uniformsGenerate", "fmb_uniforms_generate", "handles:T * F", [("l", "seed"), ("i", "T"), ("i", "F"), ("l", "paths"), ("l", "pathOffset")],
     "(int64_t)seed, T, F, (uint64_t)paths, (uint64_t)pathOffset, OUT"),
    ("eulerBlackScholes", "fmb_euler_black_scholes", "handles:T + 1",
     [("i", "scheme"), ("i", "T"), ("i", "F"), ("l", "paths"), ("D", "dt"), ("H", "dW"), ("d", "initialValue"), ("d", "riskFreeRate"), ("d", "volatility")],
     "scheme, T, F, (uint64_t)paths, dt_p, (const fmb_handle*)dW_p, initialValue, riskFreeRate, volatility, OUT"),
    ("eulerHeston", "fmb_euler_heston", "handles:(T + 1) * 2",
     [("i", "scheme"), ("i", "hestonScheme"), ("i", "T"), ("l", "paths"), ("D", "dt"), ("H", "dW"), ("d", "initialValue"), ("D", "riskFreeRate"), ("d", "volatility"),
      ("d", "theta"), ("d", "kappa"), ("d", "xi"), ("d", "rho")],
     "scheme, hestonScheme, T, (uint64_t)paths, dt_p, (const fmb_handle*)dW_p, initialValue, riskFreeRate_p, volatility, theta, kappa, xi, rho, OUT"),
    ("eulerLmm", "fmb_euler_lmm", "handles:(T + 1) * N",
     [("i", "scheme"), ("i", "measure"), ("i", "stateSpace"), ("d", "liborCap"), ("i", "T"), ("i", "N"), ("i", "F"), ("l", "paths"), ("D", "dt"), ("H", "dW"),
      ("D", "initialState"), ("D", "periodLength"), ("D", "factorLoading"), ("D", "variance"), ("I", "firstLive")],
     "scheme, measure, stateSpace, liborCap, T, N, F, (uint64_t)paths, dt_p, (const fmb_handle*)dW_p, initialState_p, periodLength_p, factorLoading_p, variance_p, "
     "(const int32_t*)firstLive_p, OUT"),
    ("eulerHullWhite", "fmb_euler_hull_white", "handles:(T + 1) * 2",
     [("i", "T"), ("l", "paths"), ("D", "dt"), ("H", "dW"), ("D", "drift0"), ("D", "drift1"), ("D", "factorLoadings")],
     "T, (uint64_t)paths, dt_p, (const fmb_handle*)dW_p, drift0_p, drift1_p, factorLoadings_p, OUT"),
    ("regressionMoments", "fmb_regression_moments", "custom", [("H", "basis"), ("D", "basisScalar"), ("h", "y")], ""),
    ("solveSvd", "fmb_regression_solve_svd", "custom", [("i", "K"), ("D", "A"), ("D", "b")], ""),
    ("regressionPredict", "fmb_regression_predict", "handle", [("H", "basis"), ("D", "basisScalar"), ("D", "x")],
     "basis_n, (const fmb_handle*)basis_p, basisScalar_p, x_p, OUT"),
    ("regressionFit", "fmb_regression_fit", "handle", [("H", "basis"), ("D", "basisScalar"), ("h", "y"), ("l", "nGlobal"), ("h", "cachedFit")],
     "basis_n, (const fmb_handle*)basis_p, basisScalar_p, (fmb_handle)y, (uint64_t)nGlobal, (fmb_handle)cachedFit, OUT"),
    ("regressionFitGet", "fmb_regression_fit_get", "custom", [("h", "fit"), ("i", "K")], ""),
    ("regressionPredictFit", "fmb_regression_predict_fit", "handle", [("H", "basis"), ("D", "basisScalar"), ("h", "fit")],
     "basis_n, (const fmb_handle*)basis_p, basisScalar_p, (fmb_handle)fit, OUT"),
    ("regressionConditionalExpectation", "fmb_regression_conditional_expectation", "custom",
     [("H", "basis"), ("D", "basisScalar"), ("h", "y"), ("l", "nGlobal"), ("h", "cachedFit"), ("H", "basisPredictor"), ("D", "basisPredictorScalar")], ""),
    ("commUniqueId", "fmb_comm_unique_id", "custom", [], ""),
    ("commInit", "fmb_comm_init", "void", [("B", "id"), ("i", "rank"), ("i", "world")], "(const unsigned char*)id_p, id_n, rank, world"),
    ("commShutdown", "fmb_comm_shutdown", "void", [], ""),
    ("commInfo", "fmb_comm_info", "custom", [], ""),
    ("commPeerHandle", "fmb_comm_peer_handle", "custom", [], ""),
    ("commPeerOpen", "fmb_comm_peer_open", "void", [("B", "handles")], "(const unsigned char*)handles_p, handles_n"),
    ("benchDfmaTflops", "fmb_bench_dfma_tflops", "double", [], "OUT"),
    ("benchCopyGbs", "fmb_bench_copy_gbs", "double", [("l", "bytes")], "(uint64_t)bytes, OUT"),
]

JTYPE = {"i": "jint", "l": "jlong", "d": "jdouble", "h": "jlong", "D": "jdoubleArray", "H": "jlongArray", "I": "jintArray", "B": "jbyteArray"}
JAVATYPE = {"i": "int", "l": "long", "d": "double", "h": "long", "D": "double[]", "H": "long[]", "I": "int[]", "B": "byte[]"}
ELEM = {"D": ("jdouble", "Double"), "H": ("jlong", "Long"), "I": ("jint", "Int"), "B": ("jbyte", "Byte")}

CUSTOM_C = {
    "lastError": ("jstring", "", "\treturn (*env)->NewStringUTF(env, fmb_last_error());\n"),
    "deviceName": ("jstring", "", "\tchar buf[256] = \"\";\n\tCHECK(fmb_device_name(buf, (int)sizeof(buf)));\n\treturn (*env)->NewStringUTF(env, buf);\n"),
    "timerStopMs": ("jdouble", "", "\tfloat ms = 0;\n\tCHECK(fmb_timer_stop_ms(&ms));\n\treturn (jdouble)ms;\n"),
    "upload": ("jlong", None,
               "\tfmb_handle out = 0;\n\tconst int rc = fmb_rv_upload(values_p, (uint64_t)values_n, &out);\n@RELEASE@\tif (rc != FMB_OK) { throwFor(env, rc); return 0; }\n\treturn (jlong)out;\n"),
    "download": ("jdoubleArray", "",
                 "\tuint64_t n = 0;\n\tCHECK(fmb_rv_size((fmb_handle)handle, &n));\n\tjdoubleArray out = (*env)->NewDoubleArray(env, (jsize)n);\n\tif (!out) return NULL;\n"
                 "\tjdouble* p = (*env)->GetDoubleArrayElements(env, out, NULL);\n\tconst int rc = fmb_rv_download((fmb_handle)handle, p, n);\n"
                 "\t(*env)->ReleaseDoubleArrayElements(env, out, p, 0);\n\tif (rc != FMB_OK) { throwFor(env, rc); return NULL; }\n\treturn out;\n"),
    "devicePointer": ("jlong", "", "\tvoid* p = NULL;\n\tCHECK(fmb_rv_device_ptr((fmb_handle)handle, &p));\n\treturn (jlong)(uintptr_t)p;\n"),
    "poolStats": ("jlongArray", "",
                  "\tuint64_t v[3] = {0, 0, 0};\n\tCHECK(fmb_pool_stats(&v[0], &v[1], &v[2]));\n\tjlongArray out = (*env)->NewLongArray(env, 3);\n\tif (!out) return NULL;\n"
                  "\tjlong w[3] = { (jlong)v[0], (jlong)v[1], (jlong)v[2] };\n\t(*env)->SetLongArrayRegion(env, out, 0, 3, w);\n\treturn out;\n"),
    "countLessOrEqual": ("jlongArray", None,
                         "\tjlongArray out = (*env)->NewLongArray(env, points_n);\n\tjlong* o = out ? (*env)->GetLongArrayElements(env, out, NULL) : NULL;\n"
                         "\tconst int rc = o ? fmb_rv_count_le((fmb_handle)x, points_p, points_n, (uint64_t*)o) : FMB_ENOMEM;\n"
                         "\tif (o) (*env)->ReleaseLongArrayElements(env, out, o, 0);\n@RELEASE@\tif (rc != FMB_OK) { throwFor(env, rc); return NULL; }\n\treturn out;\n"),
    "mtWords": ("jintArray", "",
                "\tjintArray out = (*env)->NewIntArray(env, n);\n\tif (!out) return NULL;\n\tjint* o = (*env)->GetIntArrayElements(env, out, NULL);\n"
                "\tconst int rc = fmb_mt_words((int64_t)seed, (uint64_t)wordOffset, (uint64_t)n, (uint32_t*)o);\n\t(*env)->ReleaseIntArrayElements(env, out, o, 0);\n"
                "\tif (rc != FMB_OK) { throwFor(env, rc); return NULL; }\n\treturn out;\n"),
    # moments[0 .. K*K) = XtX sums (hi), [K*K .. 2K*K) = lo, then Xty hi[K], lo[K]: the caller merges shards in double-double and divides by n
    "regressionMoments": ("jdoubleArray", None,
                          "\tconst int K = basis_n;\n\tdouble xh[64], xl[64], yh[8], yl[8];\n"
                          "\tconst int rc = K <= 8 ? fmb_regression_moments(K, (const fmb_handle*)basis_p, basisScalar_p, (fmb_handle)y, xh, xl, yh, yl) : FMB_EUNSUPPORTED;\n"
                          "@RELEASE@\tif (rc != FMB_OK) { throwFor(env, rc); return NULL; }\n\tjdoubleArray out = (*env)->NewDoubleArray(env, 2 * K * K + 2 * K);\n\tif (!out) return NULL;\n"
                          "\t(*env)->SetDoubleArrayRegion(env, out, 0, K * K, xh);\n\t(*env)->SetDoubleArrayRegion(env, out, K * K, K * K, xl);\n"
                          "\t(*env)->SetDoubleArrayRegion(env, out, 2 * K * K, K, yh);\n\t(*env)->SetDoubleArrayRegion(env, out, 2 * K * K + K, K, yl);\n\treturn out;\n"),
    # x[0 .. K) then the condition number
    "solveSvd": ("jdoubleArray", None,
                 "\tdouble x[64], cond = 0;\n\tconst int rc = (K >= 1 && K <= 64 && A_n >= K * K && b_n >= K) ? fmb_regression_solve_svd(K, A_p, b_p, x, &cond) : FMB_EINVAL;\n"
                 "@RELEASE@\tif (rc != FMB_OK) { throwFor(env, rc); return NULL; }\n\tjdoubleArray out = (*env)->NewDoubleArray(env, K + 1);\n\tif (!out) return NULL;\n"
                 "\t(*env)->SetDoubleArrayRegion(env, out, 0, K, x);\n\t(*env)->SetDoubleArrayRegion(env, out, K, 1, &cond);\n\treturn out;\n"),
    # XtX[K*K], Xty[K], x[K], cond
    "regressionFitGet": ("jdoubleArray", "",
                         "\tdouble v[64 + 8 + 8 + 1];\n\tif (K < 1 || K > 8) { throwFor(env, FMB_EINVAL); return NULL; }\n"
                         "\tCHECK(fmb_regression_fit_get((fmb_handle)fit, K, v, v + K * K, v + K * K + K, v + K * K + 2 * K));\n"
                         "\tjdoubleArray out = (*env)->NewDoubleArray(env, K * K + 2 * K + 1);\n\tif (!out) return NULL;\n"
                         "\t(*env)->SetDoubleArrayRegion(env, out, 0, K * K + 2 * K + 1, v);\n\treturn out;\n"),
    # returns { fit handle, conditional expectation handle }; basisPredictor == null: predict on the estimator's basis functions
    "regressionConditionalExpectation": ("jlongArray", None,
                                         "\tfmb_handle fit = 0, ce = 0;\n"
                                         "\tconst int rc = fmb_regression_conditional_expectation(basis_n, (const fmb_handle*)basis_p, basisScalar_p, (fmb_handle)y, (uint64_t)nGlobal, (fmb_handle)cachedFit,\n"
                                         "\t\tbasisPredictor_p ? basisPredictor_n : basis_n, (const fmb_handle*)basisPredictor_p, basisPredictorScalar_p, &fit, &ce);\n"
                                         "@RELEASE@\tif (rc != FMB_OK) { throwFor(env, rc); return NULL; }\n\tjlongArray out = (*env)->NewLongArray(env, 2);\n\tif (!out) return NULL;\n"
                                         "\tjlong w[2] = { (jlong)fit, (jlong)ce };\n\t(*env)->SetLongArrayRegion(env, out, 0, 2, w);\n\treturn out;\n"),
    "commUniqueId": ("jbyteArray", "",
                     "\tunsigned char id[128];\n\tCHECK(fmb_comm_unique_id(id, 128));\n\tjbyteArray out = (*env)->NewByteArray(env, 128);\n\tif (!out) return NULL;\n"
                     "\t(*env)->SetByteArrayRegion(env, out, 0, 128, (const jbyte*)id);\n\treturn out;\n"),
    "commPeerHandle": ("jbyteArray", "",
                       "\tunsigned char h[64];\n\tCHECK(fmb_comm_peer_handle(h, 64));\n\tjbyteArray out = (*env)->NewByteArray(env, 64);\n\tif (!out) return NULL;\n"
                       "\t(*env)->SetByteArrayRegion(env, out, 0, 64, (const jbyte*)h);\n\treturn out;\n"),
    # { rank, world, exchanges }
    "commInfo": ("jlongArray", "",
                 "\tint rank = 0, world = 1;\n\tuint64_t ex = 0;\n\tCHECK(fmb_comm_info(&rank, &world, &ex));\n\tjlongArray out = (*env)->NewLongArray(env, 3);\n\tif (!out) return NULL;\n"
                 "\tjlong w[3] = { rank, world, (jlong)ex };\n\t(*env)->SetLongArrayRegion(env, out, 0, 3, w);\n\treturn out;\n"),
}
CUSTOM_JAVA = {"lastError": "String", "deviceName": "String", "timerStopMs": "double", "upload": "long", "download": "double[]", "devicePointer": "long",
               "poolStats": "long[]", "countLessOrEqual": "long[]", "mtWords": "int[]", "regressionMoments": "double[]", "solveSvd": "double[]",
               "regressionFitGet": "double[]", "regressionConditionalExpectation": "long[]", "commUniqueId": "byte[]", "commInfo": "long[]", "commPeerHandle": "byte[]"}


def acquire(args):
    pre, post = "", ""
    for kind, name in args:
        if kind in ELEM:
            ctype, jn = ELEM[kind]
            pre += "\tconst jsize %s_n = %s ? (*env)->GetArrayLength(env, %s) : 0;\n" % (name, name, name)
            pre += "\t%s* %s_p = %s ? (*env)->Get%sArrayElements(env, %s, NULL) : NULL;\n" % (ctype, name, name, jn, name)
            pre += "\t(void)%s_n;\n" % name
            post = "\tif (%s_p) (*env)->Release%sArrayElements(env, %s, %s_p, JNI_ABORT);\n" % (name, jn, name, name) + post
    return pre, post


def zero_of(jret):
    return "" if jret == "void" else (" NULL" if jret.endswith("Array") or jret == "jstring" else " 0")


def gen_c():
    out = ['''/*
 * JNI shim: net.finmath.cuda.FinmathB200 (static native methods) -> the C ABI of include/finmath_b200.h.
 * GENERATED by gen_jni.py - do not edit.  Purely mechanical: argument marshalling and error translation, no logic.
 * Build where a JDK is present:
 *   cc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux finmath_b200_jni.c -L.. -lfinmath_b200 -o libfinmath_b200_jni.so
 * This image has no JDK: tests/test_cpu_jni.py compiles the file against tests/stubs/jni.h (written from the JNI specification) with
 * -Wall -Werror and drives every entry point through a fake JNIEnv (tests/stubs/jni_fake_env_test.c).
 */
#include <jni.h>
#include <stdint.h>
#include <stddef.h>
#include "../../../include/finmath_b200.h"

/* error translation (SURVEY.md 8b): IllegalArgumentException / UnsupportedOperationException / OutOfMemoryError / RuntimeException */
static void throwFor(JNIEnv* env, int rc) {
	const char* cls = "java/lang/RuntimeException";
	if (rc == FMB_EINVAL) cls = "java/lang/IllegalArgumentException";
	else if (rc == FMB_EUNSUPPORTED) cls = "java/lang/UnsupportedOperationException";
	else if (rc == FMB_ENOMEM) cls = "java/lang/OutOfMemoryError";
	jclass c = (*env)->FindClass(env, cls);
	if (c) (*env)->ThrowNew(env, c, fmb_last_error());
}
#define CHECK(expr) do { int rc__ = (expr); if (rc__ != FMB_OK) { throwFor(env, rc__); return RETZERO; } } while (0)
''']
    for java, cfn, ret, args, call in TABLE:
        params = "".join(", %s %s" % (JTYPE[k], n) for k, n in args)
        pre, post = acquire(args)
        if ret == "custom":
            jret, _, body = CUSTOM_C[java]
            body = body.replace("@RELEASE@", post)
            needs_arrays = "@RELEASE@" in CUSTOM_C[java][2]
            code = (pre if needs_arrays else "") + body
        elif ret == "void":
            jret = "void"
            code = pre + "\tconst int rc = %s(%s);\n" % (cfn, call) + post + "\tif (rc != FMB_OK) throwFor(env, rc);\n"
        elif ret == "nothrow":
            jret = "void"
            code = "\t%s(%s);\n" % (cfn, call)
        elif ret == "rawint":
            jret = "jint"
            code = "\treturn (jint)%s(%s);\n" % (cfn, call)
        elif ret in ("int", "long", "double", "handle"):
            jret = {"int": "jint", "long": "jlong", "double": "jdouble", "handle": "jlong"}[ret]
            ctype = {"int": "int", "long": "uint64_t", "double": "double", "handle": "fmb_handle"}[ret]
            code = pre + "\t%s out = 0;\n\tconst int rc = %s(%s);\n" % (ctype, cfn, call.replace("OUT", "&out")) + post
            code += "\tif (rc != FMB_OK) { throwFor(env, rc); return 0; }\n\treturn (%s)out;\n" % jret
        elif ret.startswith("handles:") or ret.startswith("doubles:"):
            count = ret.split(":", 1)[1]
            is_h = ret.startswith("handles")
            jret, jn, ctype = ("jlongArray", "Long", "jlong") if is_h else ("jdoubleArray", "Double", "jdouble")
            code = pre + "\t%s out = (*env)->New%sArray(env, (jsize)(%s));\n" % (jret, jn, count)
            code += "\t%s* o = out ? (*env)->Get%sArrayElements(env, out, NULL) : NULL;\n" % (ctype, jn)
            code += "\tconst int rc = o ? %s(%s) : FMB_ENOMEM;\n" % (cfn, call.replace("OUT", "(fmb_handle*)o" if is_h else "o"))
            code += "\tif (o) (*env)->Release%sArrayElements(env, out, o, 0);\n" % jn + post
            code += "\tif (rc != FMB_OK) { throwFor(env, rc); return NULL; }\n\treturn out;\n"
        else:
            raise ValueError(ret)
        out.append("#undef RETZERO\n#define RETZERO%s\nJNIEXPORT %s JNICALL Java_net_finmath_cuda_FinmathB200_%s(JNIEnv* env, jclass cls%s) {\n\t(void)env; (void)cls;\n%s}\n"
                   % (zero_of(jret), jret, java, params, code))
    return "\n".join(out)


def gen_java():
    hdr = open(os.path.join(ROOT, "include", "finmath_b200.h")).read()
    lines = []
    for java, cfn, ret, args, call in TABLE:
        if ret == "custom":
            jr = CUSTOM_JAVA[java]
        else:
            jr = {"void": "void", "nothrow": "void", "rawint": "int", "int": "int", "long": "long", "double": "double", "handle": "long"}.get(ret)
            if jr is None:
                jr = "long[]" if ret.startswith("handles") else "double[]"
        lines.append("\t/** {@code %s} */\n\tpublic static native %s %s(%s);" % (cfn, jr, java, ", ".join("%s %s" % (JAVATYPE[k], n) for k, n in args)))
    enums = []
    for m in re.finditer(r"\b(FMB_[A-Z]_[A-Z0-9_]+) = (\d+)", hdr):
        enums.append("%s = %s" % (m.group(1)[4:], m.group(2)))
    return '''package net.finmath.cuda;

/**
 * Static native entry points - one per function of include/finmath_b200.h (JNI shim: csrc/jni/finmath_b200_jni.c).
 * GENERATED by csrc/jni/gen_jni.py - do not edit.
 * Handles are opaque 64-bit ids of device-resident double vectors; 0 means "no vector, use the scalar next to it".
 * Failures surface as IllegalArgumentException / UnsupportedOperationException / OutOfMemoryError / RuntimeException.
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no JDK (see INTEGRATION.md); the shim itself is compile-checked and
 * driven through a fake JNIEnv by tests/test_cpu_jni.py.
 */
public final class FinmathB200 {
	static {
		System.loadLibrary("finmath_b200");      // the CUDA library (C ABI)
		System.loadLibrary("finmath_b200_jni");  // the shim
		init(Integer.getInteger("net.finmath.cuda.device", Integer.parseInt(System.getenv().getOrDefault("LOCAL_RANK", "0"))));
	}
	private FinmathB200() {}

	// op codes of include/finmath_b200.h
	public static final int %s;

%s
}
''' % (",\n\t\t\t".join(", ".join(enums[i:i + 6]) for i in range(0, len(enums), 6)), "\n".join(lines))


def main():
    with open(os.path.join(HERE, "finmath_b200_jni.c"), "w") as f:
        f.write(gen_c())
    with open(os.path.join(os.path.dirname(HERE), "..", "java", "net", "finmath", "cuda", "FinmathB200.java"), "w") as f:
        f.write(gen_java())


if __name__ == "__main__":
    main()
