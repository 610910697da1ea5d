/* host/shim_bench.c -- measurement driver for the reference-named object API (bench.py's second
 * end-to-end leg). It is the steady-state loop of reference src/simple.c:445-505 with the per-stream
 * pull chain replaced by coolmic_b200_batch_tick(): sources are memory iohandles (the role snddev
 * plays, snddev.c:87), consumers read each transform's handle (the role enc_vorbis.c:91 plays through
 * tee[0]) and the meters are fused. Nothing here touches cmgpu_* directly: only public objects.
 */
#include "shim_internal.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define SHIM_BENCH_STREAMING 1
#endif

typedef struct cyc_src {
    const unsigned char *data;
    size_t len, pos;
    int streaming;
} cyc_src_t;

/* What a capture device does when it fills the buffer it was handed: the destination is a pinned ring
 * slot that the CPU never reads again (the copy engine does), so the bulk goes out with non-temporal
 * stores -- no read-for-ownership of the destination lines, no cache pollution. COOLMIC_B200_BENCH_PLAIN_COPY=1
 * falls back to memcpy (A/B). */
static void capture_copy(void *dst, const void *src, size_t n, int streaming)
{
#ifdef SHIM_BENCH_STREAMING
    unsigned char *d = dst;
    const unsigned char *s = src;
    if (streaming && n >= 256) {
        const size_t head = (size_t)(-(uintptr_t)d & 15u);
        size_t i;
        memcpy(d, s, head);
        d += head; s += head; n -= head;
        for (i = 0; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(s + i));
            const __m128i b = _mm_loadu_si128((const __m128i *)(s + i + 16));
            const __m128i c = _mm_loadu_si128((const __m128i *)(s + i + 32));
            const __m128i e = _mm_loadu_si128((const __m128i *)(s + i + 48));
            _mm_stream_si128((__m128i *)(d + i), a);
            _mm_stream_si128((__m128i *)(d + i + 16), b);
            _mm_stream_si128((__m128i *)(d + i + 32), c);
            _mm_stream_si128((__m128i *)(d + i + 48), e);
        }
        _mm_sfence();
        memcpy(d + i, s + i, n - i);
        return;
    }
#else
    (void)streaming;
#endif
    memcpy(dst, src, n);
}

#ifdef SHIM_BENCH_SELFTEST
/* test hook (tests/stub/san_main.c): every size and both misalignments against memcpy; 0 = all equal */
int shim_bench_copy_selftest(void)
{
    enum { N = 4096 + 128 };
    unsigned char *src = malloc(N + 64), *a = malloc(N + 64), *b = malloc(N + 64);
    size_t n, so, dof;
    int bad = !src || !a || !b;
    for (n = 0; !bad && n < N + 64; n++)
        src[n] = (unsigned char)(n * 131u + 7u);
    for (n = 0; !bad && n <= N; n += (n < 600 ? 1 : 97)) {
        for (so = 0; so < 3 && !bad; so++) {
            for (dof = 0; dof < 33 && !bad; dof += (dof < 17 ? 1 : 15)) {
                memset(a, 0xee, N + 64);
                memset(b, 0xee, N + 64);
                capture_copy(a + dof, src + so, n, 1);
                memcpy(b + dof, src + so, n);
                bad = memcmp(a, b, N + 64) != 0;
            }
        }
    }
    free(src); free(a); free(b);
    return bad;
}
#endif

static ssize_t cyc_read(void *userdata, void *buffer, size_t len)
{
    cyc_src_t *m = userdata;
    size_t n = m->len - m->pos;
    if (n > len)
        n = len;
    capture_copy(buffer, m->data + m->pos, n, m->streaming);
    m->pos += n;
    if (m->pos == m->len)
        m->pos = 0;                     /* endless, like a capture device */
    return (ssize_t)n;
}

typedef struct consumers {
    coolmic_iohandle_t **rd;
    unsigned char *sink;                /* [threads][block bytes] */
    size_t block_bytes;
    unsigned streams, threads;
    pthread_t *tid;
    pthread_mutex_t mu;
    pthread_cond_t cv_go, cv_done;
    uint64_t job;
    unsigned left;
    int quit, failed;
    uint64_t bytes;
} consumers_t;

typedef struct consumer_arg {
    consumers_t *c;
    unsigned index;
} consumer_arg_t;

static uint64_t drain_range(consumers_t *c, unsigned index)
{
    const unsigned lo = (unsigned)((uint64_t)c->streams * index / c->threads);
    const unsigned hi = (unsigned)((uint64_t)c->streams * (index + 1) / c->threads);
    unsigned char *sink = c->sink + (size_t)index * c->block_bytes;
    uint64_t got = 0;
    unsigned s;
    for (s = lo; s < hi; s++) {
        /* one tick's worth: the handle returns what the oldest unread tick holds */
        ssize_t r = coolmic_iohandle_read(c->rd[s], sink, c->block_bytes);
        if (r < 0) {
            c->failed = 1;
            continue;
        }
        got += (uint64_t)r;
    }
    return got;
}

static void *consumer_main(void *p)
{
    consumer_arg_t *a = p;
    consumers_t *c = a->c;
    const unsigned index = a->index;
    uint64_t seen = 0;
    free(a);
    pthread_mutex_lock(&c->mu);
    for (;;) {
        uint64_t got;
        while (!c->quit && c->job == seen)
            pthread_cond_wait(&c->cv_go, &c->mu);
        if (c->quit)
            break;
        seen = c->job;
        pthread_mutex_unlock(&c->mu);
        got = drain_range(c, index);
        pthread_mutex_lock(&c->mu);
        c->bytes += got;
        if (--c->left == 0)
            pthread_cond_signal(&c->cv_done);
    }
    pthread_mutex_unlock(&c->mu);
    return NULL;
}

/* Consumers read one tick's worth from every handle: started here, joined by consume_wait(). In between
 * the caller queues the next tick -- readers of different handles and the driver may overlap as long as
 * the driver's tick does not need the slot they read (include/coolmic_b200_shim.h). */
static void consume_start(consumers_t *c)
{
    if (c->threads > 1) {
        pthread_mutex_lock(&c->mu);
        c->left = c->threads - 1;
        c->job++;
        pthread_cond_broadcast(&c->cv_go);
        pthread_mutex_unlock(&c->mu);
    }
}

static void consume_wait(consumers_t *c)
{
    const uint64_t got = drain_range(c, c->threads - 1);       /* the caller's own share */
    pthread_mutex_lock(&c->mu);
    c->bytes += got;
    while (c->threads > 1 && c->left)
        pthread_cond_wait(&c->cv_done, &c->mu);
    pthread_mutex_unlock(&c->mu);
}

static void consume_one_tick(consumers_t *c)
{
    consume_start(c);
    consume_wait(c);
}

int coolmic_b200_bench_objects(int device, unsigned int channels, unsigned int streams, unsigned int block_frames,
                               unsigned int n_ticks, unsigned int ring_slots, unsigned int threads,
                               unsigned int bytes_per_stream, const void *pcm, double *seconds,
                               uint64_t *frames_metered)
{
    coolmic_b200_batch_t *batch;
    cyc_src_t *src;
    coolmic_transform_t **tr;
    coolmic_vumeter_t **vu;
    coolmic_vumeter_result_t *results;
    int *rcs;
    consumers_t c;
    struct timespec t0, t1;
    unsigned s, t, i, lag;
    int rc = COOLMIC_ERROR_NONE;
    const char *env;
    const int streaming = !((env = getenv("COOLMIC_B200_BENCH_PLAIN_COPY")) && *env == '1');

    if (!pcm || !seconds || !frames_metered || !streams || !threads || !ring_slots || !n_ticks ||
        bytes_per_stream % (2u * channels))
        return COOLMIC_ERROR_INVAL;
    if (threads > streams)
        threads = streams;
    batch = coolmic_b200_batch_new_ring(device, channels, streams, block_frames, ring_slots, threads);
    if (!batch)
        return COOLMIC_ERROR_GENERIC;
    memset(&c, 0, sizeof(c));
    src = calloc(streams, sizeof(*src));
    tr = calloc(streams, sizeof(*tr));
    vu = calloc(streams, sizeof(*vu));
    results = calloc(streams, sizeof(*results));
    rcs = calloc(streams, sizeof(*rcs));
    c.rd = calloc(streams, sizeof(*c.rd));
    c.block_bytes = (size_t)block_frames * 2u * channels;
    c.sink = malloc(c.block_bytes * threads);
    c.tid = calloc(threads, sizeof(*c.tid));
    c.streams = streams;
    c.threads = threads;
    pthread_mutex_init(&c.mu, NULL);
    pthread_cond_init(&c.cv_go, NULL);
    pthread_cond_init(&c.cv_done, NULL);
    if (!src || !tr || !vu || !results || !rcs || !c.rd || !c.sink || !c.tid) {
        rc = COOLMIC_ERROR_NOMEM;
        goto out;
    }
    for (s = 0; s < streams; s++) {
        coolmic_iohandle_t *in;
        uint16_t gain[COOLMIC_B200_MAX_CHANNELS];
        const uint16_t scale = (uint16_t)(1000u + s % 9000u);
        for (i = 0; i < channels; i++)
            gain[i] = (uint16_t)((uint32_t)scale * 3u / 4u + 37u * ((s + i) % 64u));
        src[s].data = (const unsigned char *)pcm + (size_t)s * bytes_per_stream;
        src[s].len = bytes_per_stream;
        src[s].streaming = streaming;
        tr[s] = coolmic_b200_batch_transform_new(batch, "tr", SHIM_RO_NULL, 48000);
        in = coolmic_iohandle_new("mem", SHIM_RO_NULL, &src[s], NULL, cyc_read, NULL);
        if (!tr[s] || !in) {
            rc = COOLMIC_ERROR_GENERIC;
            shim_unref(in);
            goto out;
        }
        coolmic_transform_set_master_gain(tr[s], channels, scale, gain);
        coolmic_transform_attach_iohandle(tr[s], in);
        shim_unref(in);
        c.rd[s] = coolmic_transform_get_iohandle(tr[s]);
        vu[s] = coolmic_b200_batch_vumeter_new(batch, tr[s], "vu", SHIM_RO_NULL);
        if (!c.rd[s] || !vu[s]) {
            rc = COOLMIC_ERROR_GENERIC;
            goto out;
        }
    }
    for (i = 0; i + 1 < threads; i++) {
        consumer_arg_t *a = malloc(sizeof(*a));
        if (!a || (a->c = &c, a->index = i, pthread_create(&c.tid[i], NULL, consumer_main, a)) != 0) {
            free(a);
            c.threads = i + 1;          /* carry on with the threads that did start */
            break;
        }
    }
    /* warm-up outside the clock: one tick through the whole pipeline */
    if (coolmic_b200_batch_tick(batch) < 0) {
        rc = COOLMIC_ERROR_GENERIC;
        goto out;
    }
    consume_one_tick(&c);
    coolmic_b200_batch_results(batch, results, rcs);

    /* Steady state: while the consumers read the output of tick t - lag, the driver pulls and queues
     * tick t + 1 (its slot was read one round earlier). lag = ring_slots - 2 leaves the slot of the tick
     * being read and the slot being refilled distinct. */
    lag = ring_slots > 2 ? ring_slots - 2 : ring_slots - 1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < n_ticks; t++) {
        const int overlap = ring_slots > 2 && t >= lag;
        int fr;
        if (overlap)
            consume_start(&c);          /* output of tick t - lag: it has had `lag` ticks to come down */
        fr = coolmic_b200_batch_tick(batch);
        if (overlap)
            consume_wait(&c);
        if (fr < 0) {
            rc = fr;
            goto out;
        }
        if (!overlap && t >= lag)
            consume_one_tick(&c);
    }
    for (t = 0; t < lag && t < n_ticks; t++)
        consume_one_tick(&c);
    if (coolmic_b200_batch_results(batch, results, rcs) != COOLMIC_ERROR_NONE)
        rc = COOLMIC_ERROR_GENERIC;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    *frames_metered = 0;
    for (s = 0; s < streams; s++)
        if (rcs[s] == COOLMIC_ERROR_NONE)
            *frames_metered += results[s].frames;
    if (c.failed || c.bytes != (uint64_t)(n_ticks + 1) * streams * c.block_bytes)
        rc = COOLMIC_ERROR_GENERIC;     /* every byte that went in must have come out through the handles */
out:
    if (c.tid) {
        pthread_mutex_lock(&c.mu);
        c.quit = 1;
        pthread_cond_broadcast(&c.cv_go);
        pthread_mutex_unlock(&c.mu);
        for (i = 0; i + 1 < c.threads; i++)
            if (c.tid[i])
                pthread_join(c.tid[i], NULL);
    }
    for (s = 0; s < streams; s++) {
        if (c.rd)
            shim_unref(c.rd[s]);
        if (vu)
            shim_unref(vu[s]);
        if (tr)
            shim_unref(tr[s]);
    }
    shim_unref(batch);
    free(src); free(tr); free(vu); free(results); free(rcs); free(c.rd); free(c.sink); free(c.tid);
    pthread_mutex_destroy(&c.mu);
    pthread_cond_destroy(&c.cv_go);
    pthread_cond_destroy(&c.cv_done);
    return rc;
}
