/*
 * sgd_find_c.c — a plain C99 caller of the C ABI (include/bigclam_b200.h), no Python, no torch.
 *
 * TEST INFRASTRUCTURE: the shape of the host code a non-Python driver (the JNI shim of INTEGRATION.md) puts above
 * the library, written in the one compiled language the image has.  It walks the main body of the reference's
 * spark-shell script through the entry points that replace it:
 *
 *   GraphLoader.edgeListFile + collectNeighborIds   bigclam4-7.scala:45,50-51  -> bigclam_graph_read_edgelist
 *   conductanceLocalMin + initNeighborComF(K)       bigclam4-7.scala:58-108    -> bigclam_conductance_seeds_gpu,
 *                                                                                 bigclam_init_neighbor_com_F
 *   (or an explicit F0 file: parity runs always share F0)
 *   SGDFindC: backtrackingLineSearchs until |1 - new/old| < 1e-4   :225-243    -> bigclam_run(variant 4)
 *   F, sumF back to the driver                                      :36,38     -> bigclam_get_F, bigclam_get_sumF
 *
 * usage: sgd_find_c <edge list> <K> <max calls, 0 = unbounded> <F0.f64 | init> <out.bin> [world, default 1]
 *   F0.f64: n x K doubles, row-major, rows in ascending vertex id (the reader's dense relabelling)
 *   out.bin: int64 n, k, calls, ntrace | double llh | double trace[ntrace] | double sumF[k] | double F[n*k]
 * world > 1 drives all the GPUs behind one handle (bigclam_multi_*).  Exit code 0, or 1 with the library's message
 * on stderr — without a CUDA device that is BIGCLAM_ECUDA: there is no CPU path to fall back to.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bigclam_b200.h"

#define TRACE_CAP 4096

static int die(const char *what, int rc, const char *msg) {
    fprintf(stderr, "sgd_find_c: %s failed (%d): %s\n", what, rc, msg != NULL ? msg : "?");
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s <edge list> <K> <max calls> <F0.f64 | init> <out.bin> [world]\n", argv[0]);
        return 2;
    }
    const int32_t k = (int32_t)atoi(argv[2]);
    const int64_t max_calls = (int64_t)atoll(argv[3]);
    const int32_t world = argc > 6 ? (int32_t)atoi(argv[6]) : 1;
    char errbuf[256] = {0};
    bigclam_graph g;
    memset(&g, 0, sizeof(g));
    int rc = bigclam_graph_read_edgelist(argv[1], 1 /* dedup: simple undirected graph */, &g, errbuf, (int64_t)sizeof(errbuf));
    if (rc != BIGCLAM_OK) return die("bigclam_graph_read_edgelist", rc, errbuf);

    const size_t nk = (size_t)g.n * (size_t)k;
    double *F = (double *)malloc(nk * sizeof(double));
    double *sumF = (double *)malloc((size_t)k * sizeof(double));
    double *trace = (double *)malloc(TRACE_CAP * sizeof(double));
    if (F == NULL || sumF == NULL || trace == NULL) return die("malloc", BIGCLAM_ENOMEM, "host buffers");

    if (strcmp(argv[4], "init") == 0) {
        double *cond = (double *)malloc((size_t)g.n * sizeof(double));
        int32_t *seeds = (int32_t *)malloc((size_t)g.n * sizeof(int32_t));
        int64_t n_seeds = 0;
        if (cond == NULL || seeds == NULL) return die("malloc", BIGCLAM_ENOMEM, "seed buffers");
        rc = bigclam_conductance_seeds_gpu(g.n, g.rowptr, g.col, 0, cond, seeds, &n_seeds);
        if (rc != BIGCLAM_OK) return die("bigclam_conductance_seeds_gpu", rc, bigclam_last_error(NULL));
        rc = bigclam_init_neighbor_com_F(g.n, g.rowptr, g.col, k, seeds, n_seeds, 0, 1u, F);
        if (rc != BIGCLAM_OK) return die("bigclam_init_neighbor_com_F", rc, bigclam_last_error(NULL));
        free(cond);
        free(seeds);
    } else {
        FILE *fh = fopen(argv[4], "rb");
        if (fh == NULL || fread(F, sizeof(double), nk, fh) != nk) return die("F0 file", BIGCLAM_EIO, argv[4]);
        fclose(fh);
    }

    bigclam_params p;
    rc = bigclam_default_params(&p, k);
    if (rc != BIGCLAM_OK) return die("bigclam_default_params", rc, "bad K");
    p.device = 0;
    p.flags = BIGCLAM_F_SPARSE_ROWS;

    double llh = 0.0;
    int64_t calls = 0;
    if (world <= 1) {
        bigclam_ctx *ctx = NULL;
        rc = bigclam_create(&ctx, g.n, g.rowptr, g.col, &p);
        if (rc != BIGCLAM_OK) return die("bigclam_create", rc, bigclam_last_error(NULL));
        rc = bigclam_set_F(ctx, F);
        if (rc == BIGCLAM_OK) rc = bigclam_run(ctx, 4, 1e-4, max_calls, &llh, &calls, trace, TRACE_CAP);
        if (rc == BIGCLAM_OK) rc = bigclam_get_F(ctx, F);
        if (rc == BIGCLAM_OK) rc = bigclam_get_sumF(ctx, sumF);
        if (rc != BIGCLAM_OK) return die("single-GPU run", rc, bigclam_last_error(ctx));
        bigclam_destroy(ctx);
    } else {
        bigclam_multi *m = NULL;
        rc = bigclam_multi_create(&m, g.n, g.rowptr, g.col, &p, world, NULL);
        if (rc != BIGCLAM_OK) return die("bigclam_multi_create", rc, bigclam_multi_last_error(NULL));
        rc = bigclam_multi_set_F(m, F);
        if (rc == BIGCLAM_OK) rc = bigclam_multi_run(m, 4, 1e-4, max_calls, &llh, &calls, trace, TRACE_CAP);
        if (rc == BIGCLAM_OK) rc = bigclam_multi_get_F(m, world - 1, F);          /* the replicas are identical: read the last one */
        if (rc == BIGCLAM_OK) rc = bigclam_multi_get_sumF(m, 0, sumF);
        if (rc != BIGCLAM_OK) return die("multi-GPU run", rc, bigclam_multi_last_error(m));
        bigclam_multi_destroy(m);
    }

    const int64_t ntrace = calls < TRACE_CAP ? calls : TRACE_CAP;
    const int64_t head[4] = {g.n, (int64_t)k, calls, ntrace};
    FILE *out = fopen(argv[5], "wb");
    if (out == NULL) return die("fopen", BIGCLAM_EIO, argv[5]);
    int ok = fwrite(head, sizeof(int64_t), 4, out) == 4 && fwrite(&llh, sizeof(double), 1, out) == 1 &&
             fwrite(trace, sizeof(double), (size_t)ntrace, out) == (size_t)ntrace &&
             fwrite(sumF, sizeof(double), (size_t)k, out) == (size_t)k && fwrite(F, sizeof(double), nk, out) == nk;
    fclose(out);
    if (!ok) return die("fwrite", BIGCLAM_EIO, argv[5]);
    /* the reference's progress line (bigclam4-7.scala:238-241 prints the LLH of every call) */
    printf("SGDFindC: n=%lld K=%d world=%d calls=%lld LLH=%.17g (%s)\n", (long long)g.n, (int)k, (int)world,
           (long long)calls, llh, bigclam_version());
    free(F);
    free(sumF);
    free(trace);
    bigclam_graph_free(&g);
    return 0;
}
