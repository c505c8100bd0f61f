/*
 * batch.c — oracle restatement of src/batch.rs (TEST ONLY): data-parallel
 * fan-out over independent streams with one codec state per worker
 * (rayon `par_iter().map_init`, src/batch.rs:34-37,79-83), order-preserving
 * results (:57,100), per-stream failure reported in-band (:52-53,95-96).
 * pthreads with a shared atomic work index stand in for rayon's pool.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int orc_num_cores(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    int op; /* 0 compress, 1 decompress, 2 checksum */
    int level, format;
    const uint8_t *in;
    const uint64_t *in_off;
    size_t n;
    uint8_t *out;
    const uint64_t *out_off;
    const uint64_t *max_out;
    uint64_t *out_size;
    uint32_t *sums;
    int32_t *status;
    size_t next; /* atomic */
} job;

static void *worker(void *arg)
{
    job *j = (job *)arg;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n)
            return 0;
        const uint8_t *src = j->in + j->in_off[i];
        size_t len = (size_t)(j->in_off[i + 1] - j->in_off[i]);
        if (j->op == 2) {
            j->sums[i] = j->level ? orc_crc32(0, src, len) : orc_adler32(1, src, len);
            continue;
        }
        size_t sz = 0;
        int st;
        if (j->op == 0) {
            /* bound = *_compress_bound(len), src/batch.rs:39 */
            st = orc_compress(j->level, j->format, src, len, j->out + j->out_off[i],
                              orc_compress_bound(j->format, len), &sz);
        } else {
            size_t used = 0;
            st = orc_decompress(j->format, src, len, j->out + j->out_off[i],
                                (size_t)j->max_out[i], &used, &sz);
        }
        j->status[i] = st;
        j->out_size[i] = st == ORC_OK ? sz : 0;
    }
}

static int run(job *j, int nthreads)
{
    if (nthreads <= 0)
        nthreads = orc_num_cores();
    if ((size_t)nthreads > j->n)
        nthreads = j->n ? (int)j->n : 1;
    if (nthreads == 1) {
        worker(j);
        return 0;
    }
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (; started < nthreads; started++)
        if (pthread_create(&t[started], 0, worker, j))
            break;
    if (started == 0)
        worker(j);
    for (int k = 0; k < started; k++)
        pthread_join(t[k], 0);
    free(t);
    return 0;
}

int orc_compress_batch(int level, int format, const uint8_t *in, const uint64_t *in_off,
                       size_t n, uint8_t *out, const uint64_t *out_off, uint64_t *out_size,
                       int32_t *status, int nthreads)
{
    job j = {0, level, format, in, in_off, n, out, out_off, 0, out_size, 0, status, 0};
    return run(&j, nthreads);
}

int orc_decompress_batch(int format, const uint8_t *in, const uint64_t *in_off, size_t n,
                         uint8_t *out, const uint64_t *out_off, const uint64_t *max_out,
                         uint64_t *out_size, int32_t *status, int nthreads)
{
    job j = {1, 0, format, in, in_off, n, out, out_off, max_out, out_size, 0, status, 0};
    return run(&j, nthreads);
}

int orc_checksum_batch(int kind, const uint8_t *in, const uint64_t *in_off, size_t n,
                       uint32_t *out, int nthreads)
{
    job j = {2, kind, 0, in, in_off, n, 0, 0, 0, 0, out, 0, 0};
    return run(&j, nthreads);
}
