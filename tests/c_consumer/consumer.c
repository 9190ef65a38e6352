/*
 * tests/c_consumer/consumer.c -- a plain C host application of the Hap API, the way the reference's consumers are
 * written (/root/reference/source/hap.h:31-33: "include hap.h and link against the library").  It is compiled
 * against the REFERENCE's own header when that is present (-I/root/reference/source) and against include/hap.h
 * otherwise, and linked against libhap_b200.so: the drop-in claim taken literally.  Nothing here knows about CUDA.
 *
 *   consumer            container-only calls (no GPU needed): size bound, HapCompressorNone round trip, queries
 *   consumer --gpu      also the Snappy paths (GPU): 1 and 8 chunks, chunk callback on a pthread pool, error codes
 */
#include <hap.h>

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "consumer: line %d: %s\n", __LINE__, #c); return 1; } } while (0)

typedef struct { HapDecodeWorkFunction fn; void *p; unsigned next, count; pthread_mutex_t mu; } pool_t;
static int g_calls, g_count;

static void *pool_worker(void *arg)
{
    pool_t *q = (pool_t *)arg;
    for (;;) {
        pthread_mutex_lock(&q->mu);
        unsigned i = q->next++;
        pthread_mutex_unlock(&q->mu);
        if (i >= q->count) break;
        q->fn(q->p, i);
    }
    return NULL;
}

/* hap.h:113-128: the callback runs `function` once per chunk, from any threads, and returns when all are done */
static void threaded_callback(HapDecodeWorkFunction function, void *p, unsigned int count, void *info)
{
    (void)info;
    pool_t q = {function, p, 0, count, PTHREAD_MUTEX_INITIALIZER};
    pthread_t t[4];
    g_calls++;
    g_count = (int)count;
    for (int i = 0; i < 4; i++) pthread_create(&t[i], NULL, pool_worker, &q);
    for (int i = 0; i < 4; i++) pthread_join(t[i], NULL);
}

static void fill(unsigned char *p, size_t n)
{
    /* block-structured, partly repetitive bytes */
    for (size_t i = 0; i < n; i++) p[i] = (unsigned char)(((i / 16) % 37 == 0 ? i : (i % 16) * 7 + (i / 4096)) & 0xFF);
}

int main(int argc, char **argv)
{
    int gpu = argc > 1 && strcmp(argv[1], "--gpu") == 0;
    enum { N = 1920 * 1080 / 2 };   /* 1080p DXT1: 1 036 800 bytes */
    unsigned long lens[1] = {N};
    unsigned int fmts[1] = {HapTextureFormat_RGB_DXT1}, chunks[1] = {1}, comp[1] = {HapCompressorNone};
    unsigned long cap = HapMaxEncodedLength(1, lens, fmts, chunks);
    CHECK(cap == 1209665ul);   /* SURVEY.md 3.3 */
    unsigned char *tex = malloc(N), *frame = malloc(cap), *back = malloc(N);
    const void *ins[1] = {tex};
    unsigned long used = 0, got = 0;
    unsigned int fmt = 0, count = 0;
    int nchunks = -1;
    fill(tex, N);

    CHECK(HapEncode(1, ins, lens, fmts, comp, chunks, frame, cap, &used) == HapResult_No_Error);
    CHECK(used == N + 4 && frame[3] == 0xAB);
    CHECK(HapGetFrameTextureCount(frame, used, &count) == HapResult_No_Error && count == 1);
    CHECK(HapGetFrameTextureFormat(frame, used, 0, &fmt) == HapResult_No_Error && fmt == HapTextureFormat_RGB_DXT1);
    CHECK(HapGetFrameTextureChunkCount(frame, used, 0, &nchunks) == HapResult_No_Error && nchunks == 1);
    CHECK(HapDecode(frame, used, 0, threaded_callback, NULL, back, N, &got, &fmt) == HapResult_No_Error);
    CHECK(got == N && memcmp(back, tex, N) == 0 && g_calls == 0);
    CHECK(HapDecode(frame, used, 0, NULL, NULL, back, N, &got, &fmt) == HapResult_Bad_Arguments);   /* hap.c:1010 */
    CHECK(HapDecode(frame, used, 1, threaded_callback, NULL, back, N, &got, &fmt) == HapResult_Bad_Arguments);
    CHECK(HapDecode(frame, used, 0, threaded_callback, NULL, back, N - 1, &got, &fmt) == HapResult_Buffer_Too_Small);
    CHECK(HapEncode(1, ins, lens, fmts, comp, chunks, frame, 16, &used) == HapResult_Buffer_Too_Small);
    if (gpu) {
        for (unsigned k = 1; k <= 8; k += 7) {
            comp[0] = HapCompressorSnappy;
            chunks[0] = k;
            cap = HapMaxEncodedLength(1, lens, fmts, chunks);
            frame = realloc(frame, cap);
            CHECK(HapEncode(1, ins, lens, fmts, comp, chunks, frame, cap, &used) == HapResult_No_Error);
            CHECK(used < N && frame[3] == 0xCB);   /* complex storage, DXT1 */
            CHECK(HapGetFrameTextureChunkCount(frame, used, 0, &nchunks) == HapResult_No_Error && nchunks == (int)k);
            memset(back, 0, N);
            g_calls = 0;
            CHECK(HapDecode(frame, used, 0, threaded_callback, NULL, back, N, &got, &fmt) == HapResult_No_Error);
            CHECK(got == N && fmt == HapTextureFormat_RGB_DXT1 && memcmp(back, tex, N) == 0);
            CHECK(g_calls == (k > 1 ? 1 : 0) && (k == 1 || g_count == (int)k));   /* hap.c:852-861 */
            CHECK(HapDecode(frame, used - 7, 0, threaded_callback, NULL, back, N, &got, &fmt) == HapResult_Bad_Frame);
        }
    }
    printf("consumer ok (%s)\n", gpu ? "container + GPU paths" : "container paths");
    free(tex); free(frame); free(back);
    return 0;
}
