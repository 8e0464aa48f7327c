/* SOURCE ONLY -- not compiled or tested in this image (no jni.h).  See INTEGRATION.md.
 * Build on a box with a JDK: gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *   -o libcco_b200_jni.so cco_jni.c -L../universal_recommender_b200/csrc -lcco_b200 */
#include <jni.h>
#include "cco_b200.h"
static cco_ctx_t *g_ctx;                       /* one context per JVM (single driver thread) */
static void throw_cco(JNIEnv *env, int st) {
  const char *cls = (st == CCO_E_INVALID_ARG || st == CCO_E_SHAPE_MISMATCH) ? "java/lang/IllegalArgumentException"
                                                                             : "java/lang/RuntimeException";
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), cco_last_error());
}
JNIEXPORT jlong JNICALL Java_com_actionml_b200_B200SimilarityAnalysis_00024_train(JNIEnv *env, jobject self,
    jobjectArray rowPtr, jobjectArray colIdx, jlong nRows, jintArray nCols, jintArray m, jintArray k,
    jbooleanArray hasMin, jdoubleArray minLlr, jint seed) {
  if (!g_ctx) { cco_config_t cfg = {0, 0, 1, 0, NULL}; int st = cco_create(&cfg, &g_ctx); if (st) { throw_cco(env, st); return 0; } }
  jsize n = (*env)->GetArrayLength(env, rowPtr);
  cco_csr_t mats[64]; cco_indicator_params_t prm[64];
  jint *nc = (*env)->GetIntArrayElements(env, nCols, NULL), *mm = (*env)->GetIntArrayElements(env, m, NULL),
       *kk = (*env)->GetIntArrayElements(env, k, NULL);
  jboolean *hm = (*env)->GetBooleanArrayElements(env, hasMin, NULL);
  jdouble *ml = (*env)->GetDoubleArrayElements(env, minLlr, NULL);
  for (jsize i = 0; i < n; ++i) {
    mats[i].n_rows = nRows; mats[i].n_cols = nc[i];
    mats[i].row_ptr = (const int64_t *)(*env)->GetDirectBufferAddress(env, (*env)->GetObjectArrayElement(env, rowPtr, i));
    mats[i].col_idx = (const int32_t *)(*env)->GetDirectBufferAddress(env, (*env)->GetObjectArrayElement(env, colIdx, i));
    prm[i].max_interactions = mm[i]; prm[i].top_k = kk[i]; prm[i].has_min_llr = hm[i]; prm[i].min_llr = ml[i];
  }
  cco_result_t *res = NULL;
  int st = cco_train(g_ctx, n, mats, prm, seed, 0, &res);
  /* ... ReleaseXxxArrayElements ... */
  if (st) { throw_cco(env, st); return 0; }
  return (jlong)(intptr_t)res;
}
/* resultMatrix: cco_result_matrix(...) + three NewDirectByteBuffer over the pinned result arrays; resultFree: cco_result_free */
