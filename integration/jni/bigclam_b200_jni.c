/*
 * bigclam_b200_jni.c — the JNI side of INTEGRATION.md §2: every entry of the Scala facade `object BigclamNative`
 * is a 1:1 forward to the C ABI (include/bigclam_b200.h).  Build on a box with a JDK:
 *
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude integration/jni/bigclam_b200_jni.c \
 *       -o libbigclam_b200_jni.so -Lbigclam_apachespark_b200 -l:libbigclam_b200.so
 *
 * There is no JDK in the build image: tests/test_jni_shim.py compiles this file against a stand-in <jni.h>
 * (tests/jni_stub/jni.h: the types and the JNIEnv functions used here, same names and signatures as the JNI
 * specification) and drives the entry points from a fake JVM written in C (tests/jni_stub/fake_jvm.c) — that checks
 * the forwarding code, not a JVM.  `Java_BigclamNative_00024_*` is the JNI name mangling of a method of the Scala
 * `object BigclamNative` (class `BigclamNative$`).
 *
 * What each entry replaces in codes/bigclam4-7.scala:
 *   create   Neightborbc broadcast (:50-51) + the script variables (:22-43)        -> bigclam_create
 *   setF     F / sumF after initNeighborComF (:105-107)                             -> bigclam_set_F
 *   step     backtrackingLineSearchs(uset) (:152-223), null uset = all (:227)       -> bigclam_step
 *   nUpdated Sx.size of the last call (:186)                                        -> (kept from bigclam_step)
 *   run      SGDFindC / MBSGD outer loop (:225-243; v3 :206-222; v2 :203-219)       -> bigclam_run
 *   getF     F.collect (:36)                                                        -> bigclam_get_F
 *   setFCsr / getFNnz / getFCsr   F as RDD[(Long, BSV[Double])] rows (:36, :97-104)   -> bigclam_set_F_csr, _get_F_nnz, _get_F_csr
 *   createMulti / the same calls on a multi handle: all GPUs of the box             -> bigclam_multi_*
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>

#include "bigclam_b200.h"

typedef struct {
    bigclam_ctx *ctx;        /* one GPU ...                          */
    bigclam_multi *multi;    /* ... or all GPUs behind one handle    */
    int64_t n_updated;       /* of the most recent step              */
} jni_handle;

static const char *last_error(const jni_handle *h) {
    if (h == NULL) return bigclam_last_error(NULL);
    return h->multi != NULL ? bigclam_multi_last_error(h->multi) : bigclam_last_error(h->ctx);
}

/* The reference reports failures as JVM exceptions; the C ABI returns codes: re-throw. */
static int throw_on(JNIEnv *env, int rc, const char *msg) {
    if (rc != BIGCLAM_OK) {
        jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
        if (cls != NULL) (*env)->ThrowNew(env, cls, msg != NULL ? msg : "bigclam_b200 error");
    }
    return rc;
}

static jlong create_common(JNIEnv *env, jlong n, jlongArray rowptr, jintArray col, jint k, jdouble alpha, jdouble beta,
                           jint maxInter, jint device, jint world) {
    bigclam_params p;
    if (throw_on(env, bigclam_default_params(&p, k), "bad K")) return 0;     /* MIN_P_/MAX_P_/MIN_F_/MAX_F_ of :39-43 */
    p.alpha = alpha;
    p.beta = beta;
    p.max_inter = maxInter;
    p.device = device;
    p.flags = BIGCLAM_F_SPARSE_ROWS;                                          /* F as the reference keeps it: sparse rows */
    jni_handle *h = (jni_handle *)calloc(1, sizeof(jni_handle));
    if (h == NULL) {
        throw_on(env, BIGCLAM_ENOMEM, "out of host memory");
        return 0;
    }
    jlong *rp = (*env)->GetLongArrayElements(env, rowptr, NULL);
    jint *cl = (*env)->GetIntArrayElements(env, col, NULL);
    int rc = BIGCLAM_ENOMEM;
    if (rp != NULL && cl != NULL) {
        if (world <= 1) rc = bigclam_create(&h->ctx, n, (const int64_t *)rp, (const int32_t *)cl, &p);
        else rc = bigclam_multi_create(&h->multi, n, (const int64_t *)rp, (const int32_t *)cl, &p, world, NULL);
    }
    if (rp != NULL) (*env)->ReleaseLongArrayElements(env, rowptr, rp, JNI_ABORT);
    if (cl != NULL) (*env)->ReleaseIntArrayElements(env, col, cl, JNI_ABORT);
    if (throw_on(env, rc, world <= 1 ? bigclam_last_error(NULL) : bigclam_multi_last_error(NULL))) {
        free(h);
        return 0;
    }
    return (jlong)(intptr_t)h;
}

JNIEXPORT jlong JNICALL Java_BigclamNative_00024_create(JNIEnv *env, jobject self, jlong n, jlongArray rowptr, jintArray col,
                                                        jint k, jdouble alpha, jdouble beta, jint maxInter, jint device) {
    (void)self;
    return create_common(env, n, rowptr, col, k, alpha, beta, maxInter, device, 1);
}

JNIEXPORT jlong JNICALL Java_BigclamNative_00024_createMulti(JNIEnv *env, jobject self, jlong n, jlongArray rowptr, jintArray col,
                                                             jint k, jdouble alpha, jdouble beta, jint maxInter, jint world) {
    (void)self;
    return create_common(env, n, rowptr, col, k, alpha, beta, maxInter, 0, world);
}

JNIEXPORT void JNICALL Java_BigclamNative_00024_setF(JNIEnv *env, jobject self, jlong handle, jdoubleArray F) {
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    jdouble *f = (*env)->GetDoubleArrayElements(env, F, NULL);               /* row-major n x k */
    int rc = BIGCLAM_ENOMEM;
    if (f != NULL) {
        rc = h->multi != NULL ? bigclam_multi_set_F(h->multi, f) : bigclam_set_F(h->ctx, f);
        (*env)->ReleaseDoubleArrayElements(env, F, f, JNI_ABORT);
    }
    throw_on(env, rc, last_error(h));
}

JNIEXPORT jdouble JNICALL Java_BigclamNative_00024_step(JNIEnv *env, jobject self, jlong handle, jbyteArray usetMask) {
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    jbyte *m = usetMask != NULL ? (*env)->GetByteArrayElements(env, usetMask, NULL) : NULL;   /* null == all vertices (:227) */
    double llh = 0.0;
    int rc = h->multi != NULL ? bigclam_multi_step(h->multi, (const uint8_t *)m, &llh, &h->n_updated)
                              : bigclam_step(h->ctx, (const uint8_t *)m, &llh, &h->n_updated);        /* replaces :154-222 */
    if (m != NULL) (*env)->ReleaseByteArrayElements(env, usetMask, m, JNI_ABORT);
    throw_on(env, rc, last_error(h));
    return llh;
}

JNIEXPORT jlong JNICALL Java_BigclamNative_00024_nUpdated(JNIEnv *env, jobject self, jlong handle) {
    (void)env;
    (void)self;
    return (jlong)((jni_handle *)(intptr_t)handle)->n_updated;
}

/* Returns the LLH the reference's loop returns (:242); calls[0] receives the number of hot-path calls. */
JNIEXPORT jdouble JNICALL Java_BigclamNative_00024_run(JNIEnv *env, jobject self, jlong handle, jint variant, jdouble relTol,
                                                       jlong maxOuter, jlongArray calls) {
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    double llh = 0.0;
    int64_t ncalls = 0;
    int rc = h->multi != NULL ? bigclam_multi_run(h->multi, variant, relTol, maxOuter, &llh, &ncalls, NULL, 0)
                              : bigclam_run(h->ctx, variant, relTol, maxOuter, &llh, &ncalls, NULL, 0);
    if (throw_on(env, rc, last_error(h)) == BIGCLAM_OK && calls != NULL && (*env)->GetArrayLength(env, calls) > 0) {
        jlong *c = (*env)->GetLongArrayElements(env, calls, NULL);
        if (c != NULL) {
            c[0] = (jlong)ncalls;
            (*env)->ReleaseLongArrayElements(env, calls, c, 0);              /* 0: copy back */
        }
    }
    return llh;
}

JNIEXPORT void JNICALL Java_BigclamNative_00024_getF(JNIEnv *env, jobject self, jlong handle, jdoubleArray out) {
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    jdouble *f = (*env)->GetDoubleArrayElements(env, out, NULL);
    int rc = BIGCLAM_ENOMEM;
    if (f != NULL) {
        rc = h->multi != NULL ? bigclam_multi_get_F(h->multi, 0, f) : bigclam_get_F(h->ctx, f);
        (*env)->ReleaseDoubleArrayElements(env, out, f, rc == BIGCLAM_OK ? 0 : JNI_ABORT);
    }
    throw_on(env, rc, last_error(h));
}

JNIEXPORT void JNICALL Java_BigclamNative_00024_getSumF(JNIEnv *env, jobject self, jlong handle, jdoubleArray out) {
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    jdouble *s = (*env)->GetDoubleArrayElements(env, out, NULL);
    int rc = BIGCLAM_ENOMEM;
    if (s != NULL) {
        rc = h->multi != NULL ? bigclam_multi_get_sumF(h->multi, 0, s) : bigclam_get_sumF(h->ctx, s);
        (*env)->ReleaseDoubleArrayElements(env, out, s, rc == BIGCLAM_OK ? 0 : JNI_ABORT);
    }
    throw_on(env, rc, last_error(h));
}

/*
 * F in the reference's own shape, RDD[(Long, BSV[Double])] (:36, :97-104): per row the BSV's (index, data) arrays,
 * concatenated in dense-id order with an indptr — no dense n x K array on either side (bigclam_set_F_csr / _get_F_csr).
 */
JNIEXPORT void JNICALL Java_BigclamNative_00024_setFCsr(JNIEnv *env, jobject self, jlong handle, jlongArray indptr, jintArray indices,
                                                        jdoubleArray values) {
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    jlong *ip = (*env)->GetLongArrayElements(env, indptr, NULL);
    jint *ix = (*env)->GetIntArrayElements(env, indices, NULL);
    jdouble *v = (*env)->GetDoubleArrayElements(env, values, NULL);
    int rc = BIGCLAM_ENOMEM;
    if (ip != NULL && ix != NULL && v != NULL)
        rc = h->multi != NULL ? bigclam_multi_set_F_csr(h->multi, (const int64_t *)ip, (const int32_t *)ix, v)
                              : bigclam_set_F_csr(h->ctx, (const int64_t *)ip, (const int32_t *)ix, v);
    if (ip != NULL) (*env)->ReleaseLongArrayElements(env, indptr, ip, JNI_ABORT);
    if (ix != NULL) (*env)->ReleaseIntArrayElements(env, indices, ix, JNI_ABORT);
    if (v != NULL) (*env)->ReleaseDoubleArrayElements(env, values, v, JNI_ABORT);
    throw_on(env, rc, last_error(h));
}

JNIEXPORT jlong JNICALL Java_BigclamNative_00024_getFNnz(JNIEnv *env, jobject self, jlong handle) {
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    int64_t nnz = 0;
    int rc = h->multi != NULL ? bigclam_multi_get_F_nnz(h->multi, &nnz) : bigclam_get_F_nnz(h->ctx, &nnz);
    throw_on(env, rc, last_error(h));
    return (jlong)nnz;
}

/* indptr: n + 1 longs; indices / values: getFNnz() entries (ascending component inside a row, zeros never stored). */
JNIEXPORT void JNICALL Java_BigclamNative_00024_getFCsr(JNIEnv *env, jobject self, jlong handle, jlongArray indptr, jintArray indices,
                                                        jdoubleArray values) {
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    jlong *ip = (*env)->GetLongArrayElements(env, indptr, NULL);
    jint *ix = (*env)->GetIntArrayElements(env, indices, NULL);
    jdouble *v = (*env)->GetDoubleArrayElements(env, values, NULL);
    int rc = BIGCLAM_ENOMEM;
    if (ip != NULL && ix != NULL && v != NULL)
        rc = h->multi != NULL ? bigclam_multi_get_F_csr(h->multi, (int64_t *)ip, (int32_t *)ix, v)
                              : bigclam_get_F_csr(h->ctx, (int64_t *)ip, (int32_t *)ix, v);
    const jint mode = rc == BIGCLAM_OK ? 0 : JNI_ABORT;
    if (ip != NULL) (*env)->ReleaseLongArrayElements(env, indptr, ip, mode);
    if (ix != NULL) (*env)->ReleaseIntArrayElements(env, indices, ix, mode);
    if (v != NULL) (*env)->ReleaseDoubleArrayElements(env, values, v, mode);
    throw_on(env, rc, last_error(h));
}

JNIEXPORT void JNICALL Java_BigclamNative_00024_destroy(JNIEnv *env, jobject self, jlong handle) {
    (void)env;
    (void)self;
    jni_handle *h = (jni_handle *)(intptr_t)handle;
    if (h == NULL) return;
    if (h->multi != NULL) bigclam_multi_destroy(h->multi);
    if (h->ctx != NULL) bigclam_destroy(h->ctx);
    free(h);
}
