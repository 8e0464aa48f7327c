/* Minimal declarations of the JNI names jni/cco_jni.c uses, so that the shim gets a syntax/type check in an image
 * without a JDK (tests/test_abi.py::test_jni_shim_compiles).  NOT a JNI implementation; never linked. */
#ifndef CCO_TEST_JNI_MIN_H
#define CCO_TEST_JNI_MIN_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef jint jsize;
typedef unsigned char jboolean;
typedef double jdouble;
typedef struct _jobject *jobject;
typedef jobject jclass, jarray, jobjectArray, jintArray, jbooleanArray, jdoubleArray, jthrowable;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv *, const char *);
  jint (*ThrowNew)(JNIEnv *, jclass, const char *);
  void (*DeleteLocalRef)(JNIEnv *, jobject);
  jsize (*GetArrayLength)(JNIEnv *, jarray);
  jobjectArray (*NewObjectArray)(JNIEnv *, jsize, jclass, jobject);
  jobject (*GetObjectArrayElement)(JNIEnv *, jobjectArray, jsize);
  void (*SetObjectArrayElement)(JNIEnv *, jobjectArray, jsize, jobject);
  jint *(*GetIntArrayElements)(JNIEnv *, jintArray, jboolean *);
  jboolean *(*GetBooleanArrayElements)(JNIEnv *, jbooleanArray, jboolean *);
  jdouble *(*GetDoubleArrayElements)(JNIEnv *, jdoubleArray, jboolean *);
  void (*ReleaseIntArrayElements)(JNIEnv *, jintArray, jint *, jint);
  void (*ReleaseBooleanArrayElements)(JNIEnv *, jbooleanArray, jboolean *, jint);
  void (*ReleaseDoubleArrayElements)(JNIEnv *, jdoubleArray, jdouble *, jint);
  jobject (*NewDirectByteBuffer)(JNIEnv *, void *, jlong);
  void *(*GetDirectBufferAddress)(JNIEnv *, jobject);
};
#endif
