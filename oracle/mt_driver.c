/*
 * oracle/mt_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Threaded drivers for the CPU baseline (bench.py --impl reference and the cpu_baseline leg):
 *   - decode: one HapDecode whose HapDecodeCallback fans hap_decode_chunk out over a pthread pool,
 *     i.e. the reference's own parallel mechanism (/root/reference/source/hap.c:861, hap.h:113-128);
 *   - encode: HapEncode has no callback and a serial chunk loop (hap.c:448-476), so "all cores"
 *     means frame-parallel: T threads each encoding their own frames.
 * Built twice: into oracle/_ref/libhap_ref.so against the unmodified reference (DRV(x) = x) and into
 * oracle/liboracle.so against the restatement (DRV(x) = orc_##x).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>

#ifdef DRV_ORACLE
#include "hap_oracle.h"
#define DRV(x) orc_##x
#define DRVNAME(x) orcdrv_##x
typedef orc_work_fn work_fn;
typedef orc_decode_cb decode_cb;
#else
#include "hap.h"
#define DRV(x) x
#define DRVNAME(x) refdrv_##x
typedef HapDecodeWorkFunction work_fn;
typedef HapDecodeCallback decode_cb;
#endif

typedef struct {
    work_fn fn;
    void *p;
    unsigned count, threads;
    volatile unsigned next;
} fanout;

static void *fanout_worker(void *arg)
{
    fanout *f = (fanout *)arg;
    for (;;) {
        unsigned i = __atomic_fetch_add(&f->next, 1, __ATOMIC_RELAXED);
        if (i >= f->count) break;
        f->fn(f->p, i);
    }
    return NULL;
}

static void pool_callback(work_fn fn, void *p, unsigned count, void *info)
{
    fanout f = {fn, p, count, *(unsigned *)info, 0};
    unsigned t = f.threads < count ? f.threads : count;
    pthread_t tid[256];
    if (t > 256) t = 256;
    for (unsigned i = 1; i < t; i++) pthread_create(&tid[i], NULL, fanout_worker, &f);
    fanout_worker(&f);
    for (unsigned i = 1; i < t; i++) pthread_join(tid[i], NULL);
}

/* one frame, chunks fanned out over `threads` workers */
unsigned DRVNAME(decode_mt)(const void *frame, unsigned long bytes, unsigned index, void *out,
                            unsigned long cap, unsigned long *used, unsigned *fmt, unsigned threads)
{
    return DRV(HapDecode)(frame, bytes, index, pool_callback, &threads, out, cap, used, fmt);
}

typedef struct {
    unsigned tid, threads, frames;
    unsigned count;
    const void **ins;            /* frames * count input pointers */
    unsigned long *lens;         /* count */
    unsigned *fmts, *compressors, *chunks;
    uint8_t *out;                /* frames * out_stride */
    unsigned long out_stride;
    unsigned long *used;         /* frames */
    unsigned result;
} enc_job;

static void *enc_worker(void *arg)
{
    enc_job *j = (enc_job *)arg;
    for (unsigned f = j->tid; f < j->frames; f += j->threads) {
        unsigned r = DRV(HapEncode)(j->count, j->ins + (size_t)f * j->count, j->lens, j->fmts,
                                    j->compressors, j->chunks, j->out + (size_t)f * j->out_stride,
                                    j->out_stride, &j->used[f]);
        if (r) j->result = r;
    }
    return NULL;
}

/* `frames` frames, frame-parallel over `threads` workers */
unsigned DRVNAME(encode_frames_mt)(unsigned frames, unsigned count, const void **ins, unsigned long *lens,
                                   unsigned *fmts, unsigned *compressors, unsigned *chunks, void *out,
                                   unsigned long out_stride, unsigned long *used, unsigned threads)
{
    if (threads == 0) threads = 1;
    if (threads > 256) threads = 256;
    enc_job jobs[256];
    pthread_t tid[256];
    for (unsigned t = 0; t < threads; t++) {
        enc_job j = {t, threads, frames, count, ins, lens, fmts, compressors, chunks, (uint8_t *)out, out_stride, used, 0};
        jobs[t] = j;
        if (t) pthread_create(&tid[t], NULL, enc_worker, &jobs[t]);
    }
    enc_worker(&jobs[0]);
    unsigned r = jobs[0].result;
    for (unsigned t = 1; t < threads; t++) {
        pthread_join(tid[t], NULL);
        if (jobs[t].result) r = jobs[t].result;
    }
    return r;
}

/* `frames` frames decoded frame-parallel (each frame's chunks serial inside its thread) */
typedef struct {
    unsigned tid, threads, frames;
    const uint8_t *in;
    unsigned long in_stride;
    const unsigned long *in_bytes;
    uint8_t *out;
    unsigned long out_stride;
    unsigned result;
} dec_job;

static void serial_callback(work_fn fn, void *p, unsigned count, void *info)
{
    (void)info;
    for (unsigned i = 0; i < count; i++) fn(p, i);
}

static void *dec_worker(void *arg)
{
    dec_job *j = (dec_job *)arg;
    for (unsigned f = j->tid; f < j->frames; f += j->threads) {
        unsigned long used;
        unsigned fmt;
        unsigned r = DRV(HapDecode)(j->in + (size_t)f * j->in_stride, j->in_bytes[f], 0, serial_callback, NULL,
                                    j->out + (size_t)f * j->out_stride, j->out_stride, &used, &fmt);
        if (r) j->result = r;
    }
    return NULL;
}

unsigned DRVNAME(decode_frames_mt)(unsigned frames, const void *in, unsigned long in_stride,
                                   const unsigned long *in_bytes, void *out, unsigned long out_stride,
                                   unsigned threads)
{
    if (threads == 0) threads = 1;
    if (threads > 256) threads = 256;
    dec_job jobs[256];
    pthread_t tid[256];
    for (unsigned t = 0; t < threads; t++) {
        dec_job j = {t, threads, frames, (const uint8_t *)in, in_stride, in_bytes, (uint8_t *)out, out_stride, 0};
        jobs[t] = j;
        if (t) pthread_create(&tid[t], NULL, dec_worker, &jobs[t]);
    }
    dec_worker(&jobs[0]);
    unsigned r = jobs[0].result;
    for (unsigned t = 1; t < threads; t++) {
        pthread_join(tid[t], NULL);
        if (jobs[t].result) r = jobs[t].result;
    }
    return r;
}
