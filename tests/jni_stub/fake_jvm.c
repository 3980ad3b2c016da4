/*
 * fake_jvm.c — TEST INFRASTRUCTURE: plays the JVM for integration/jni/bigclam_b200_jni.c (no JDK in the image).
 * Implements the JNIEnv functions of tests/jni_stub/jni.h the way a copying JVM does — Get<T>ArrayElements hands out
 * a COPY, Release with mode 0 copies it back, JNI_ABORT drops it — so a forwarder that releases with the wrong mode
 * loses its result here as it would there; a pending exception is a recorded (class, message).  main() is the
 * changed part of the reference's driver script in C:
 *
 *   fake_jvm <edge list> <K> <max calls> <F0.f64> <out.bin> <world> <run | step | csr>
 *     run : create(Multi) -> setF -> run(4, 1e-4, max calls) -> getF, getSumF          (SGDFindC, bigclam4-7.scala:225-243)
 *     step: create(Multi) -> setF -> max calls x { step(null); nUpdated } -> getF, getSumF   (:152-223 per call)
 *     csr : like run, but F travels as rows of (index, data) both ways: setFCsr, getFNnz + getFCsr   (the BSV rows of :36)
 *   out.bin: int64 n, k, calls, ntrace | double llh | double trace[ntrace] | double sumF[k] | double F[n*k]
 *            (same layout as tests/c_host/sgd_find_c.c; trace = the LLH of every step in `step` mode)
 * Exit code 3 with "EXCEPTION <class>: <message>" on stderr when the shim threw.
 */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bigclam_b200.h"

/* the shim's entry points (integration/jni/bigclam_b200_jni.c) */
jlong Java_BigclamNative_00024_create(JNIEnv *, jobject, jlong, jlongArray, jintArray, jint, jdouble, jdouble, jint, jint);
jlong Java_BigclamNative_00024_createMulti(JNIEnv *, jobject, jlong, jlongArray, jintArray, jint, jdouble, jdouble, jint, jint);
void Java_BigclamNative_00024_setF(JNIEnv *, jobject, jlong, jdoubleArray);
jdouble Java_BigclamNative_00024_step(JNIEnv *, jobject, jlong, jbyteArray);
jlong Java_BigclamNative_00024_nUpdated(JNIEnv *, jobject, jlong);
jdouble Java_BigclamNative_00024_run(JNIEnv *, jobject, jlong, jint, jdouble, jlong, jlongArray);
void Java_BigclamNative_00024_getF(JNIEnv *, jobject, jlong, jdoubleArray);
void Java_BigclamNative_00024_getSumF(JNIEnv *, jobject, jlong, jdoubleArray);
void Java_BigclamNative_00024_setFCsr(JNIEnv *, jobject, jlong, jlongArray, jintArray, jdoubleArray);
jlong Java_BigclamNative_00024_getFNnz(JNIEnv *, jobject, jlong);
void Java_BigclamNative_00024_getFCsr(JNIEnv *, jobject, jlong, jlongArray, jintArray, jdoubleArray);
void Java_BigclamNative_00024_destroy(JNIEnv *, jobject, jlong);

typedef struct {
    void *data;
    jsize len;
    size_t elem;
    int outstanding;            /* copies handed out and not yet released */
} fake_array;

static char g_exc_class[128], g_exc_msg[512];
static int g_exc_pending = 0;
static long g_outstanding = 0;

static jobject new_array(size_t elem, jsize len, const void *init) {
    fake_array *a = (fake_array *)calloc(1, sizeof(fake_array));
    a->data = calloc((size_t)len > 0 ? (size_t)len : 1, elem);
    a->len = len;
    a->elem = elem;
    if (init != NULL) memcpy(a->data, init, (size_t)len * elem);
    return (jobject)a;
}
static void free_array(jobject o) {
    fake_array *a = (fake_array *)o;
    free(a->data);
    free(a);
}
static void *get_elems(jarray arr, jboolean *is_copy) {
    fake_array *a = (fake_array *)arr;
    void *c = malloc((size_t)a->len * a->elem + 1);
    memcpy(c, a->data, (size_t)a->len * a->elem);
    if (is_copy != NULL) *is_copy = 1;
    a->outstanding++;
    g_outstanding++;
    return c;
}
static void release_elems(jarray arr, void *elems, jint mode) {
    fake_array *a = (fake_array *)arr;
    if (mode != JNI_ABORT) memcpy(a->data, elems, (size_t)a->len * a->elem);     /* 0 (and JNI_COMMIT): copy back */
    free(elems);
    a->outstanding--;
    g_outstanding--;
}

static jclass f_FindClass(JNIEnv *env, const char *name) {
    (void)env;
    static char cls[128];
    snprintf(cls, sizeof(cls), "%s", name);
    return (jclass)cls;
}
static jint f_ThrowNew(JNIEnv *env, jclass clazz, const char *msg) {
    (void)env;
    snprintf(g_exc_class, sizeof(g_exc_class), "%s", (const char *)clazz);
    snprintf(g_exc_msg, sizeof(g_exc_msg), "%s", msg);
    g_exc_pending = 1;
    return 0;
}
static jsize f_GetArrayLength(JNIEnv *env, jarray a) { (void)env; return ((fake_array *)a)->len; }
static jbyte *f_GetByte(JNIEnv *env, jbyteArray a, jboolean *c) { (void)env; return (jbyte *)get_elems(a, c); }
static jint *f_GetInt(JNIEnv *env, jintArray a, jboolean *c) { (void)env; return (jint *)get_elems(a, c); }
static jlong *f_GetLong(JNIEnv *env, jlongArray a, jboolean *c) { (void)env; return (jlong *)get_elems(a, c); }
static jdouble *f_GetDouble(JNIEnv *env, jdoubleArray a, jboolean *c) { (void)env; return (jdouble *)get_elems(a, c); }
static void f_RelByte(JNIEnv *env, jbyteArray a, jbyte *e, jint m) { (void)env; release_elems(a, e, m); }
static void f_RelInt(JNIEnv *env, jintArray a, jint *e, jint m) { (void)env; release_elems(a, e, m); }
static void f_RelLong(JNIEnv *env, jlongArray a, jlong *e, jint m) { (void)env; release_elems(a, e, m); }
static void f_RelDouble(JNIEnv *env, jdoubleArray a, jdouble *e, jint m) { (void)env; release_elems(a, e, m); }

static const struct JNINativeInterface_ g_table = {
    f_FindClass, f_ThrowNew, f_GetArrayLength, f_GetByte, f_GetInt, f_GetLong, f_GetDouble, f_RelByte, f_RelInt, f_RelLong, f_RelDouble,
};

static int pending(void) {
    if (!g_exc_pending) return 0;
    fprintf(stderr, "EXCEPTION %s: %s\n", g_exc_class, g_exc_msg);
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 8) {
        fprintf(stderr, "usage: %s <edge list> <K> <max calls> <F0.f64> <out.bin> <world> <run | step>\n", argv[0]);
        return 2;
    }
    JNIEnv env_obj = &g_table;
    JNIEnv *env = &env_obj;
    const jint k = (jint)atoi(argv[2]);
    const jlong max_calls = (jlong)atoll(argv[3]);
    const jint world = (jint)atoi(argv[6]);
    const int step_mode = strcmp(argv[7], "step") == 0;
    const int csr_mode = strcmp(argv[7], "csr") == 0;

    char errbuf[256] = {0};
    bigclam_graph g;
    memset(&g, 0, sizeof(g));
    if (bigclam_graph_read_edgelist(argv[1], 1, &g, errbuf, (int64_t)sizeof(errbuf)) != BIGCLAM_OK) {
        fprintf(stderr, "reader: %s\n", errbuf);
        return 1;
    }
    const size_t nk = (size_t)g.n * (size_t)k;
    double *F0 = (double *)malloc(nk * sizeof(double));
    FILE *fh = fopen(argv[4], "rb");
    if (fh == NULL || fread(F0, sizeof(double), nk, fh) != nk) {
        fprintf(stderr, "F0 file %s\n", argv[4]);
        return 1;
    }
    fclose(fh);

    /* val ctx = BigclamNative.create(ids.length, rowptr, col, i, alpha, beta, MaxInter, 0) */
    jobject rowptr = new_array(sizeof(jlong), (jsize)(g.n + 1), g.rowptr);
    jobject col = new_array(sizeof(jint), (jsize)g.nnz, g.col);
    jlong ctx = world <= 1 ? Java_BigclamNative_00024_create(env, NULL, (jlong)g.n, rowptr, col, k, 0.05, 0.1, 15, 0)
                           : Java_BigclamNative_00024_createMulti(env, NULL, (jlong)g.n, rowptr, col, k, 0.05, 0.1, 15, world);
    if (pending()) return 3;
    /* BigclamNative.setF(ctx, denseRowMajor(F, ids, i)) */
    jobject F = new_array(sizeof(jdouble), (jsize)nk, F0);
    if (csr_mode) {
        /* rows.flatMap(_._2.index) / rows.flatMap(_._2.data): the non-zeros of every row, in dense-id order */
        jlong *ip = (jlong *)calloc((size_t)g.n + 1, sizeof(jlong));
        jint *ix = (jint *)malloc((nk > 0 ? nk : 1) * sizeof(jint));
        jdouble *v = (jdouble *)malloc((nk > 0 ? nk : 1) * sizeof(jdouble));
        jlong cnt = 0;
        for (int64_t u = 0; u < g.n; ++u) {
            for (jint c = k - 1; c >= 0; --c)                                   /* any order inside a row is allowed: descending */
                if (F0[(size_t)u * (size_t)k + (size_t)c] != 0.0) {
                    ix[cnt] = c;
                    v[cnt++] = F0[(size_t)u * (size_t)k + (size_t)c];
                }
            ip[u + 1] = cnt;
        }
        jobject a_ip = new_array(sizeof(jlong), (jsize)(g.n + 1), ip), a_ix = new_array(sizeof(jint), (jsize)cnt, ix),
                a_v = new_array(sizeof(jdouble), (jsize)cnt, v);
        Java_BigclamNative_00024_setFCsr(env, NULL, ctx, a_ip, a_ix, a_v);
        free_array(a_ip);
        free_array(a_ix);
        free_array(a_v);
        free(ip);
        free(ix);
        free(v);
    } else {
        Java_BigclamNative_00024_setF(env, NULL, ctx, F);
    }
    if (pending()) return 3;

    double llh = 0.0;
    jlong calls = 0, ntrace = 0;
    double *trace = (double *)calloc((size_t)(max_calls > 0 ? max_calls : 1), sizeof(double));
    if (step_mode) {
        for (jlong it = 0; it < max_calls; ++it) {
            llh = Java_BigclamNative_00024_step(env, NULL, ctx, NULL);       /* def backtrackingLineSearchs(uset) = step(ctx, null) */
            if (pending()) return 3;
            trace[ntrace++] = llh;
            printf("step %lld: LLH %.17g, %lld rows updated\n", (long long)it, llh,
                   (long long)Java_BigclamNative_00024_nUpdated(env, NULL, ctx));
        }
        calls = max_calls;
    } else {
        jlong zero = 0;
        jobject ncalls = new_array(sizeof(jlong), 1, &zero);
        llh = Java_BigclamNative_00024_run(env, NULL, ctx, 4, 1e-4, max_calls, ncalls);
        if (pending()) return 3;
        calls = ((jlong *)((fake_array *)ncalls)->data)[0];
        free_array(ncalls);
    }
    jobject sumF = new_array(sizeof(jdouble), (jsize)k, NULL);
    if (csr_mode) {
        const jlong nnz = Java_BigclamNative_00024_getFNnz(env, NULL, ctx);
        if (pending()) return 3;
        jobject a_ip = new_array(sizeof(jlong), (jsize)(g.n + 1), NULL), a_ix = new_array(sizeof(jint), (jsize)nnz, NULL),
                a_v = new_array(sizeof(jdouble), (jsize)nnz, NULL);
        Java_BigclamNative_00024_getFCsr(env, NULL, ctx, a_ip, a_ix, a_v);
        if (pending()) return 3;
        const jlong *ip = (const jlong *)((fake_array *)a_ip)->data;
        const jint *ix = (const jint *)((fake_array *)a_ix)->data;
        const jdouble *v = (const jdouble *)((fake_array *)a_v)->data;
        double *dense = (double *)((fake_array *)F)->data;
        memset(dense, 0, nk * sizeof(double));
        if (ip[g.n] != nnz) {
            fprintf(stderr, "getFCsr: indptr ends at %lld, getFNnz said %lld\n", (long long)ip[g.n], (long long)nnz);
            return 4;
        }
        for (int64_t u = 0; u < g.n; ++u)
            for (jlong e = ip[u]; e < ip[u + 1]; ++e) dense[(size_t)u * (size_t)k + (size_t)ix[e]] = v[e];
        free_array(a_ip);
        free_array(a_ix);
        free_array(a_v);
    } else {
        Java_BigclamNative_00024_getF(env, NULL, ctx, F);
        if (pending()) return 3;
    }
    Java_BigclamNative_00024_getSumF(env, NULL, ctx, sumF);
    if (pending()) return 3;
    Java_BigclamNative_00024_destroy(env, NULL, ctx);
    if (g_outstanding != 0) {
        fprintf(stderr, "%ld array copies were never released\n", g_outstanding);
        return 4;
    }

    const int64_t head[4] = {g.n, (int64_t)k, (int64_t)calls, (int64_t)ntrace};
    FILE *out = fopen(argv[5], "wb");
    if (out == NULL) return 1;
    fwrite(head, sizeof(int64_t), 4, out);
    fwrite(&llh, sizeof(double), 1, out);
    fwrite(trace, sizeof(double), (size_t)ntrace, out);
    fwrite(((fake_array *)sumF)->data, sizeof(double), (size_t)k, out);
    fwrite(((fake_array *)F)->data, sizeof(double), nk, out);
    fclose(out);
    free_array(rowptr);
    free_array(col);
    free_array(F);
    free_array(sumF);
    free(F0);
    free(trace);
    bigclam_graph_free(&g);
    return 0;
}
