/*
 * cco_jni.c -- JNI shim between com.actionml.b200.B200SimilarityAnalysis (jni/B200SimilarityAnalysis.scala) and the
 * C ABI of include/cco_b200.h.  It replaces the two Mahout calls of the reference,
 * /root/reference/src/main/scala/URAlgorithm.scala:323-329 and :343-346.
 *
 * SOURCE ONLY: this image has no JDK (no jni.h), so the file is compiled here only against the minimal declarations of
 * tests/abi/jni_min.h (tests/test_abi.py::test_jni_shim_compiles) -- a syntax/type check, not a run.  Build on a box with
 * a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *       -o libcco_b200_jni.so cco_jni.c -L../universal_recommender_b200/csrc -lcco_b200
 *
 * Threading: URAlgorithm.train runs on the single Spark-driver thread (URAlgorithm.scala:292-307), so one context per JVM
 * is enough.  The context is a GROUP context over every B200 of the box (cco_create_group): the library runs one host
 * thread per GPU internally, so the single JVM thread reaches all 8 GPUs without a second process.
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>

#include "cco_b200.h"

#define CCO_JNI_MAX_MATS 64 /* event types per train; the reference's engine.json files name 1-10 */

static cco_ctx_t *g_ctx; /* created on first use, lives as long as the JVM */

static void throw_cco(JNIEnv *env, int st) {
  /* Mahout: Preconditions.checkArgument -> IllegalArgumentException; everything else aborts `pio train` as RuntimeException */
  const char *cls = (st == CCO_E_INVALID_ARG || st == CCO_E_SHAPE_MISMATCH) ? "java/lang/IllegalArgumentException"
                                                                             : "java/lang/RuntimeException";
  jclass c = (*env)->FindClass(env, cls);
  if (c) (*env)->ThrowNew(env, c, cco_last_error());
}

static int ensure_ctx(JNIEnv *env) {
  if (g_ctx) return 0;
  int n = cco_device_count();
  if (n < 1) {
    throw_cco(env, n < 0 ? n : CCO_E_CUDA);
    return -1;
  }
  int32_t devices[64];
  if (n > 64) n = 64;
  for (int i = 0; i < n; ++i) devices[i] = i;
  int st = cco_create_group(n, devices, &g_ctx);
  if (st) {
    throw_cco(env, st);
    return -1;
  }
  return 0;
}

/* pinned staging for the Scala side: returns a direct ByteBuffer over cco_host_alloc memory (PCIe-speed H2D) */
JNIEXPORT jobject JNICALL Java_com_actionml_b200_B200SimilarityAnalysis_00024_hostAlloc(JNIEnv *env, jobject self, jlong bytes) {
  (void)self;
  if (ensure_ctx(env)) return NULL;
  void *p = NULL;
  int st = cco_host_alloc(g_ctx, (size_t)bytes, &p);
  if (st) {
    throw_cco(env, st);
    return NULL;
  }
  return (*env)->NewDirectByteBuffer(env, p, bytes);
}

JNIEXPORT void JNICALL Java_com_actionml_b200_B200SimilarityAnalysis_00024_hostFree(JNIEnv *env, jobject self, jobject buf) {
  (void)self;
  if (g_ctx && buf) cco_host_free(g_ctx, (*env)->GetDirectBufferAddress(env, buf));
}

/* train: direct ByteBuffers (row_ptr int64 / col_idx int32, native order) per event type -> cco_result_t* as jlong */
JNIEXPORT jlong JNICALL Java_com_actionml_b200_B200SimilarityAnalysis_00024_train(JNIEnv *env, jobject self, jobjectArray rowPtr,
                                                                                  jobjectArray colIdx, jlong nRows, jintArray nCols,
                                                                                  jintArray m, jintArray k, jbooleanArray hasMin,
                                                                                  jdoubleArray minLlr, jint seed, jint flags) {
  (void)self;
  if (ensure_ctx(env)) return 0;
  const jsize n = (*env)->GetArrayLength(env, rowPtr);
  if (n < 1 || n > CCO_JNI_MAX_MATS || (*env)->GetArrayLength(env, colIdx) != n || (*env)->GetArrayLength(env, nCols) != n ||
      (*env)->GetArrayLength(env, m) != n || (*env)->GetArrayLength(env, k) != n || (*env)->GetArrayLength(env, hasMin) != n ||
      (*env)->GetArrayLength(env, minLlr) != n) {
    jclass c = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
    if (c) (*env)->ThrowNew(env, c, "B200SimilarityAnalysis.train: between 1 and 64 event types, all arrays of equal length");
    return 0;
  }
  cco_csr_t mats[CCO_JNI_MAX_MATS];
  cco_indicator_params_t prm[CCO_JNI_MAX_MATS];
  jint *nc = (*env)->GetIntArrayElements(env, nCols, NULL);
  jint *mm = (*env)->GetIntArrayElements(env, m, NULL);
  jint *kk = (*env)->GetIntArrayElements(env, k, NULL);
  jboolean *hm = (*env)->GetBooleanArrayElements(env, hasMin, NULL);
  jdouble *ml = (*env)->GetDoubleArrayElements(env, minLlr, NULL);
  int bad = !nc || !mm || !kk || !hm || !ml;
  for (jsize i = 0; i < n && !bad; ++i) {
    jobject rp = (*env)->GetObjectArrayElement(env, rowPtr, i), ci = (*env)->GetObjectArrayElement(env, colIdx, i);
    mats[i].n_rows = nRows;
    mats[i].n_cols = nc[i];
    mats[i].row_ptr = rp ? (const int64_t *)(*env)->GetDirectBufferAddress(env, rp) : NULL;
    mats[i].col_idx = ci ? (const int32_t *)(*env)->GetDirectBufferAddress(env, ci) : NULL;
    if (!mats[i].row_ptr) bad = 1; /* not a direct buffer */
    prm[i].max_interactions = mm[i];
    prm[i].top_k = kk[i];
    prm[i].has_min_llr = hm[i] ? 1 : 0;
    prm[i].min_llr = ml[i];
    (*env)->DeleteLocalRef(env, rp);
    (*env)->DeleteLocalRef(env, ci);
  }
  cco_result_t *res = NULL;
  int st = bad ? CCO_E_INVALID_ARG : cco_train(g_ctx, n, mats, prm, seed, (uint32_t)flags, &res);
  /* the library never keeps host pointers: everything can be released as soon as cco_train returns */
  if (nc) (*env)->ReleaseIntArrayElements(env, nCols, nc, JNI_ABORT);
  if (mm) (*env)->ReleaseIntArrayElements(env, m, mm, JNI_ABORT);
  if (kk) (*env)->ReleaseIntArrayElements(env, k, kk, JNI_ABORT);
  if (hm) (*env)->ReleaseBooleanArrayElements(env, hasMin, hm, JNI_ABORT);
  if (ml) (*env)->ReleaseDoubleArrayElements(env, minLlr, ml, JNI_ABORT);
  if (bad) {
    jclass c = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
    if (c) (*env)->ThrowNew(env, c, "B200SimilarityAnalysis.train: row_ptr/col_idx must be direct ByteBuffers");
    return 0;
  }
  if (st) {
    throw_cco(env, st);
    return 0;
  }
  return (jlong)(intptr_t)res;
}

/* indicator i of a result as three direct ByteBuffers over the library's pinned arrays: row_ptr (int64, n_rows + 1),
 * col_idx (int32, nnz), llr (float64, nnz; absent -> null when the train ran with CCO_FLAG_RESULT_NO_LLR).
 * The buffers are views: they die with resultFree. */
JNIEXPORT jobjectArray JNICALL Java_com_actionml_b200_B200SimilarityAnalysis_00024_resultMatrix(JNIEnv *env, jobject self, jlong res,
                                                                                                jint i) {
  (void)self;
  const cco_result_t *r = (const cco_result_t *)(intptr_t)res;
  int64_t n_rows = 0;
  int32_t n_cols = 0;
  const int64_t *rp = NULL;
  const int32_t *ci = NULL, *cnt = NULL;
  const double *llr = NULL;
  int st = cco_result_matrix(r, i, &n_rows, &n_cols, &rp, &ci, &llr, &cnt);
  if (st) {
    throw_cco(env, st);
    return NULL;
  }
  (void)cnt; /* k11 of the kept cells: the reference consumer (package.scala:100-108) has no use for it */
  const int64_t nnz = rp[n_rows];
  jclass bb = (*env)->FindClass(env, "java/nio/ByteBuffer");
  if (!bb) return NULL;
  jobjectArray out = (*env)->NewObjectArray(env, 3, bb, NULL);
  if (!out) return NULL;
  (*env)->SetObjectArrayElement(env, out, 0, (*env)->NewDirectByteBuffer(env, (void *)rp, (jlong)(8 * (n_rows + 1))));
  (*env)->SetObjectArrayElement(env, out, 1, (*env)->NewDirectByteBuffer(env, (void *)ci, (jlong)(4 * nnz)));
  if (llr) (*env)->SetObjectArrayElement(env, out, 2, (*env)->NewDirectByteBuffer(env, (void *)llr, (jlong)(8 * nnz)));
  return out;
}

JNIEXPORT jint JNICALL Java_com_actionml_b200_B200SimilarityAnalysis_00024_resultNumMatrices(JNIEnv *env, jobject self, jlong res) {
  (void)env;
  (void)self;
  return cco_result_num_matrices((const cco_result_t *)(intptr_t)res);
}

JNIEXPORT void JNICALL Java_com_actionml_b200_B200SimilarityAnalysis_00024_resultFree(JNIEnv *env, jobject self, jlong res) {
  (void)env;
  (void)self;
  cco_result_free((cco_result_t *)(intptr_t)res);
}

/* JVM shutdown hook (Runtime.addShutdownHook in the Scala object) */
JNIEXPORT void JNICALL Java_com_actionml_b200_B200SimilarityAnalysis_00024_shutdown(JNIEnv *env, jobject self) {
  (void)env;
  (void)self;
  if (g_ctx) {
    cco_destroy(g_ctx);
    g_ctx = NULL;
  }
}
