/*
 * tests/stubs/jni.h — a MINIMAL stand-in for the JDK's <jni.h>, written from the public JNI specification (Java Native Interface
 * Specification, chapters 3 "JNI Types and Data Structures" and 4 "JNI Functions"), for two purposes only:
 *   1. compile-checking finmath-lib_b200/csrc/jni/finmath_b200_jni.c with -Wall -Werror in an image that has no JDK,
 *   2. driving the shim's Java_* entry points through a fake JNIEnv (jni_fake_env_test.c).
 * It declares the primitive and reference types of the specification and ONLY the JNIEnv functions the shim uses, with the
 * specification's names and signatures; the table layout is NOT the JDK's (a production build uses $JAVA_HOME/include/jni.h, where
 * the same source compiles unchanged because it calls the functions by name through (*env)->).
 */
#ifndef FMB_TEST_STUB_JNI_H
#define FMB_TEST_STUB_JNI_H
#include <stdint.h>

typedef uint8_t jboolean;
typedef int8_t jbyte;
typedef uint16_t jchar;
typedef int16_t jshort;
typedef int32_t jint;
typedef int64_t jlong;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jthrowable;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jdoubleArray;

#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_OK 0
#define JNI_COMMIT 1
#define JNI_ABORT 2
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNIIMPORT

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
	void* reserved0;
	jclass (JNICALL *FindClass)(JNIEnv* env, const char* name);
	jint (JNICALL *ThrowNew)(JNIEnv* env, jclass clazz, const char* message);
	jthrowable (JNICALL *ExceptionOccurred)(JNIEnv* env);
	void (JNICALL *ExceptionClear)(JNIEnv* env);
	jstring (JNICALL *NewStringUTF)(JNIEnv* env, const char* bytes);
	jsize (JNICALL *GetArrayLength)(JNIEnv* env, jarray array);
	jbyteArray (JNICALL *NewByteArray)(JNIEnv* env, jsize length);
	jintArray (JNICALL *NewIntArray)(JNIEnv* env, jsize length);
	jlongArray (JNICALL *NewLongArray)(JNIEnv* env, jsize length);
	jdoubleArray (JNICALL *NewDoubleArray)(JNIEnv* env, jsize length);
	jbyte* (JNICALL *GetByteArrayElements)(JNIEnv* env, jbyteArray array, jboolean* isCopy);
	jint* (JNICALL *GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
	jlong* (JNICALL *GetLongArrayElements)(JNIEnv* env, jlongArray array, jboolean* isCopy);
	jdouble* (JNICALL *GetDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jboolean* isCopy);
	void (JNICALL *ReleaseByteArrayElements)(JNIEnv* env, jbyteArray array, jbyte* elems, jint mode);
	void (JNICALL *ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
	void (JNICALL *ReleaseLongArrayElements)(JNIEnv* env, jlongArray array, jlong* elems, jint mode);
	void (JNICALL *ReleaseDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jdouble* elems, jint mode);
	void (JNICALL *SetByteArrayRegion)(JNIEnv* env, jbyteArray array, jsize start, jsize len, const jbyte* buf);
	void (JNICALL *SetLongArrayRegion)(JNIEnv* env, jlongArray array, jsize start, jsize len, const jlong* buf);
	void (JNICALL *SetDoubleArrayRegion)(JNIEnv* env, jdoubleArray array, jsize start, jsize len, const jdouble* buf);
};
#endif
