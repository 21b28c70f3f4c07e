/*
 * oracle/cpu_baseline.c -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * All-core CPU baseline driver for bench.py: one restated Convproc per stream,
 * one stream per thread (the way folve uses its cores: one Convproc per open
 * file, README.md:361-362), PCM already in memory, block loop as in
 * SoundProcessor::Process (sound-processor.cc:98-127).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "zita_oracle.h"

typedef struct {
    int inp, out;
    int ind0, len;
    const float *data;
} zb_impulse;

typedef struct {
    int ninp, nout;
    unsigned size, fragm;
    int nimp;
    const zb_impulse *imp;
    int nblocks;
    float amplitude;
    unsigned seed;
    double checksum;
    int rc;
    pthread_barrier_t *start, *stop;
} zb_job;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *worker(void *arg) {
    zb_job *j = (zb_job *)arg;
    zo_convproc *p = zo_new();
    j->rc = zo_configure(p, (uint32_t)j->ninp, (uint32_t)j->nout, j->size, j->fragm, j->fragm, j->fragm, 0.0f);
    for (int k = 0; !j->rc && k < j->nimp; k++) {
        const zb_impulse *m = &j->imp[k];
        j->rc = zo_impdata_create(p, (uint32_t)m->inp, (uint32_t)m->out, 1, m->data, m->ind0, m->ind0 + m->len);
    }
    if (!j->rc) { zo_reset(p); j->rc = zo_start_process(p, 0, 0); }
    /* one block of interleaved synthetic PCM per stream, reused for every block */
    const size_t n = (size_t)j->fragm;
    float *pcm = (float *)malloc(sizeof(float) * n * (size_t)(j->ninp > j->nout ? j->ninp : j->nout));
    uint32_t s = j->seed * 2654435761u + 12345u;
    for (size_t i = 0; i < n * (size_t)j->ninp; i++) {
        s = s * 1664525u + 1013904223u;
        pcm[i] = j->amplitude * ((float)(s >> 8) * (1.0f / 8388608.0f) - 1.0f);
    }
    double acc = 0.0;
    pthread_barrier_wait(j->start);
    if (!j->rc) {
        for (int b = 0; b < j->nblocks; b++) {
            for (int ch = 0; ch < j->ninp; ch++) {
                float *dst = zo_inpdata(p, (uint32_t)ch);
                for (size_t f = 0; f < n; f++) dst[f] = pcm[f * (size_t)j->ninp + (size_t)ch];
            }
            zo_process(p);
            for (int ch = 0; ch < j->nout; ch++) {
                const float *src = zo_outdata(p, (uint32_t)ch);
                float mx = 0.0f;
                for (size_t f = 0; f < n; f++) if (src[f] > mx) mx = src[f];
                acc += mx;
            }
        }
    }
    pthread_barrier_wait(j->stop);
    j->checksum = acc;
    free(pcm);
    zo_delete(p);
    return NULL;
}

/* Runs nthreads independent streams of nblocks blocks each; returns the wall
 * seconds of the processing phase (filter loading excluded), <0 on error. */
double zb_run(int nthreads, int nblocks, int ninp, int nout, unsigned size, unsigned fragm,
              int nimp, const zb_impulse *imp, float amplitude, double *checksum_out) {
    if (nthreads < 1) return -1.0;
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(*th));
    zb_job *jobs = (zb_job *)calloc((size_t)nthreads, sizeof(*jobs));
    pthread_barrier_t start, stop;
    pthread_barrier_init(&start, NULL, (unsigned)nthreads + 1);
    pthread_barrier_init(&stop, NULL, (unsigned)nthreads + 1);
    for (int t = 0; t < nthreads; t++) {
        zb_job *j = &jobs[t];
        j->ninp = ninp; j->nout = nout; j->size = size; j->fragm = fragm;
        j->nimp = nimp; j->imp = imp; j->nblocks = nblocks; j->amplitude = amplitude;
        j->seed = (unsigned)t + 1u; j->start = &start; j->stop = &stop;
        pthread_create(&th[t], NULL, worker, j);
    }
    pthread_barrier_wait(&start);
    const double t0 = now_s();
    pthread_barrier_wait(&stop);
    const double t1 = now_s();
    double cs = 0.0;
    int rc = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        cs += jobs[t].checksum;
        rc |= jobs[t].rc;
    }
    pthread_barrier_destroy(&start);
    pthread_barrier_destroy(&stop);
    free(th);
    free(jobs);
    if (checksum_out) *checksum_out = cs;
    return rc ? -1.0 : (t1 - t0);
}
