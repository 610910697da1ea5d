/* oracle/ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin driver around the UNMODIFIED reference objects (compiled by oracle/Makefile
 * straight from /root/reference/src into oracle/_ref/). Nothing in here computes
 * audio: it only wires reference objects together the way the reference does
 * (src/simple.c:183-236: source -> transform -> tee -> {consumer, vumeter}) and
 * copies what they produce into flat, ctypes-friendly structures.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load the resulting library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

#include <coolmic-dsp/coolmic-dsp.h>
#include <coolmic-dsp/iohandle.h>
#include <coolmic-dsp/transform.h>
#include <coolmic-dsp/vumeter.h>
#include <coolmic-dsp/tee.h>
#include <coolmic-dsp/snddev.h>

/* Flat copy of coolmic_vumeter_result_t (include/coolmic-dsp/vumeter.h:48-83) with
 * fixed-width members so Python does not depend on the host ABI's padding. */
typedef struct refh_result {
    int32_t  rc;               /* return code of coolmic_vumeter_result() */
    uint32_t rate;
    uint32_t channels;
    int32_t  global_peak;
    uint64_t frames;
    double   global_power;
    int32_t  channel_peak[16];
    double   channel_power[16];
} refh_result_t;

static void flatten(refh_result_t *dst, int rc, const coolmic_vumeter_result_t *src)
{
    unsigned c;
    memset(dst, 0, sizeof(*dst));
    dst->rc = rc;
    if (rc != COOLMIC_ERROR_NONE)
        return;
    dst->rate = src->rate;
    dst->channels = src->channels;
    dst->global_peak = src->global_peak;
    dst->frames = src->frames;
    dst->global_power = src->global_power;
    for (c = 0; c < 16; c++) {
        dst->channel_peak[c] = src->channel_peak[c];
        dst->channel_power[c] = src->channel_power[c];
    }
}

/* ---- memory source: an iohandle over a caller-owned byte range ---------------- */
typedef struct memsrc {
    const char *data;
    size_t len;
    size_t pos;
    size_t chunk;      /* max bytes handed out per read callback; 0 = unlimited */
} memsrc_t;

static ssize_t memsrc_read(void *userdata, void *buffer, size_t len)
{
    memsrc_t *m = userdata;
    size_t n = m->len - m->pos;
    if (n > len)
        n = len;
    if (m->chunk && n > m->chunk)
        n = m->chunk;
    memcpy(buffer, m->data + m->pos, n);
    m->pos += n;
    return (ssize_t)n;
}

static int memsrc_eof(void *userdata)
{
    memsrc_t *m = userdata;
    return m->pos >= m->len;
}

static coolmic_iohandle_t *memsrc_handle(memsrc_t *m, const void *data, size_t len, size_t chunk)
{
    m->data = data;
    m->len = len;
    m->pos = 0;
    m->chunk = chunk;
    return coolmic_iohandle_new("memsrc", igloo_RO_NULL, m, NULL, memsrc_read, memsrc_eof);
}

unsigned refh_sizeof_result(void) { return (unsigned)sizeof(coolmic_vumeter_result_t); }

/* ---- snddev_sine through the real driver --------------------------------------- */
long refh_sine(unsigned rate, unsigned channels, size_t nbytes, size_t pull, void *out)
{
    coolmic_snddev_t *dev = coolmic_snddev_new("sine", igloo_RO_NULL, COOLMIC_DSP_SNDDEV_DRIVER_SINE, NULL,
                                               rate, channels, COOLMIC_DSP_SNDDEV_RX, -1);
    coolmic_iohandle_t *h;
    size_t done = 0;

    if (!dev)
        return -1;
    h = coolmic_snddev_get_iohandle(dev);
    if (!h) {
        igloo_ro_unref(dev);
        return -1;
    }
    if (!pull)
        pull = 1024;
    while (done < nbytes) {
        size_t want = nbytes - done < pull ? nbytes - done : pull;
        ssize_t r = coolmic_iohandle_read(h, (char *)out + done, want);
        if (r <= 0)
            break;
        done += (size_t)r;
    }
    igloo_ro_unref(h);
    igloo_ro_unref(dev);
    return (long)done;
}

/* ---- mem -> transform -> caller ------------------------------------------------ */
long refh_transform(const void *in, size_t in_bytes, unsigned rate, unsigned channels,
                    int set_gain, unsigned gain_n, unsigned scale, const uint16_t *gain,
                    size_t src_chunk, size_t pull, void *out, size_t out_cap, int *gain_rc)
{
    memsrc_t mem;
    coolmic_iohandle_t *src, *h;
    coolmic_transform_t *tr;
    size_t done = 0;

    tr = coolmic_transform_new("tr", igloo_RO_NULL, rate, channels);
    if (!tr)
        return -1;
    if (set_gain) {
        int rc = coolmic_transform_set_master_gain(tr, gain_n, (uint16_t)scale, gain);
        if (gain_rc)
            *gain_rc = rc;
    }
    src = memsrc_handle(&mem, in, in_bytes, src_chunk);
    coolmic_transform_attach_iohandle(tr, src);
    igloo_ro_unref(src);
    h = coolmic_transform_get_iohandle(tr);
    if (!pull)
        pull = 1024;
    for (;;) {
        size_t want = out_cap - done < pull ? out_cap - done : pull;
        ssize_t r;
        if (!want)
            break;
        r = coolmic_iohandle_read(h, (char *)out + done, want);
        if (r <= 0)
            break;
        done += (size_t)r;
    }
    igloo_ro_unref(h);
    igloo_ro_unref(tr);
    return (long)done;
}

/* ---- mem -> vumeter ------------------------------------------------------------ */
long refh_vumeter(const void *in, size_t in_bytes, unsigned rate, unsigned channels,
                  size_t src_chunk, long maxlen, unsigned result_every,
                  refh_result_t *results, size_t results_cap)
{
    memsrc_t mem;
    coolmic_iohandle_t *src;
    coolmic_vumeter_t *vu;
    coolmic_vumeter_result_t res;
    size_t n_results = 0;
    unsigned good = 0;

    vu = coolmic_vumeter_new("vu", igloo_RO_NULL, rate, channels);
    if (!vu)
        return -1;
    src = memsrc_handle(&mem, in, in_bytes, src_chunk);
    coolmic_vumeter_attach_iohandle(vu, src);
    igloo_ro_unref(src);
    for (;;) {
        ssize_t r = coolmic_vumeter_read(vu, maxlen);
        if (r <= 0)
            break;
        if (result_every && ++good == result_every) {
            int rc = coolmic_vumeter_result(vu, &res);
            good = 0;
            if (n_results < results_cap)
                flatten(&results[n_results++], rc, &res);
        }
    }
    {
        int rc = coolmic_vumeter_result(vu, &res);
        if (n_results < results_cap)
            flatten(&results[n_results++], rc, &res);
    }
    igloo_ro_unref(vu);
    return (long)n_results;
}

/* ---- the reference wiring: mem -> transform -> tee -> {consumer, vumeter} ------- */
typedef struct pipeline_job {
    const void *in;
    size_t in_bytes;
    unsigned rate, channels;
    int set_gain;
    unsigned gain_n, scale;
    const uint16_t *gain;
    size_t src_chunk, pull;
    unsigned result_every;
    void *out;
    size_t out_cap;
    refh_result_t *results;
    size_t results_cap;
    int last_only;     /* keep only the newest result, in results[0] */
    /* outputs */
    size_t out_bytes;
    size_t n_results;
    int gain_rc;
} pipeline_job_t;

static void store_result(pipeline_job_t *j, int rc, const coolmic_vumeter_result_t *res)
{
    if (j->last_only) {
        if (j->results)
            flatten(&j->results[0], rc, res);
    } else if (j->results && j->n_results < j->results_cap) {
        flatten(&j->results[j->n_results], rc, res);
    }
    j->n_results++;
}

static int run_pipeline(pipeline_job_t *j)
{
    memsrc_t mem;
    coolmic_iohandle_t *src, *h;
    coolmic_transform_t *tr;
    coolmic_tee_t *tee;
    coolmic_vumeter_t *vu;
    coolmic_iohandle_t *consumer;
    coolmic_vumeter_result_t res;
    char scratch[8192];
    unsigned good = 0;
    size_t pull = j->pull ? j->pull : 1024;

    if (pull > sizeof(scratch))
        pull = sizeof(scratch);

    j->out_bytes = 0;
    j->n_results = 0;
    j->gain_rc = 0;

    tr = coolmic_transform_new("tr", igloo_RO_NULL, j->rate, j->channels);
    vu = coolmic_vumeter_new("vu", igloo_RO_NULL, j->rate, j->channels);
    tee = coolmic_tee_new("tee", igloo_RO_NULL, 2);
    if (!tr || !vu || !tee)
        return -1;
    if (j->set_gain)
        j->gain_rc = coolmic_transform_set_master_gain(tr, j->gain_n, (uint16_t)j->scale, j->gain);

    /* same order as src/simple.c:212-229 */
    src = memsrc_handle(&mem, j->in, j->in_bytes, j->src_chunk);
    coolmic_transform_attach_iohandle(tr, src);
    igloo_ro_unref(src);
    h = coolmic_transform_get_iohandle(tr);
    coolmic_tee_attach_iohandle(tee, h);
    igloo_ro_unref(h);
    consumer = coolmic_tee_get_iohandle(tee, 0);
    h = coolmic_tee_get_iohandle(tee, 1);
    coolmic_vumeter_attach_iohandle(vu, h);
    igloo_ro_unref(h);

    for (;;) {
        /* the encoder's pull (src/enc_vorbis.c:78,91 reads 1024-byte chunks) */
        ssize_t r = coolmic_iohandle_read(consumer, scratch, pull);
        ssize_t v;
        if (r > 0 && j->out) {
            size_t n = (size_t)r;
            if (j->out_bytes + n > j->out_cap)
                n = j->out_cap - j->out_bytes;
            memcpy((char *)j->out + j->out_bytes, scratch, n);
        }
        if (r > 0)
            j->out_bytes += (size_t)r;
        /* the meter's pull (src/simple.c:477) and cadence (src/simple.c:486-499) */
        v = coolmic_vumeter_read(vu, -1);
        if (v > 0 && j->result_every && ++good == j->result_every) {
            int rc = coolmic_vumeter_result(vu, &res);
            good = 0;
            store_result(j, rc, &res);
        }
        if (r <= 0 && v <= 0)
            break;
    }
    {
        int rc = coolmic_vumeter_result(vu, &res);
        /* a window that ended exactly on the cadence leaves nothing to report */
        if (!(j->last_only && rc != COOLMIC_ERROR_NONE && j->n_results))
            store_result(j, rc, &res);
    }

    igloo_ro_unref(consumer);
    igloo_ro_unref(vu);
    igloo_ro_unref(tee);
    igloo_ro_unref(tr);
    return 0;
}

long refh_pipeline(const void *in, size_t in_bytes, unsigned rate, unsigned channels,
                   int set_gain, unsigned gain_n, unsigned scale, const uint16_t *gain,
                   size_t src_chunk, size_t pull, unsigned result_every,
                   void *out, size_t out_cap, size_t *out_bytes,
                   refh_result_t *results, size_t results_cap, int *gain_rc)
{
    pipeline_job_t j;
    memset(&j, 0, sizeof(j));
    j.in = in; j.in_bytes = in_bytes; j.rate = rate; j.channels = channels;
    j.set_gain = set_gain; j.gain_n = gain_n; j.scale = scale; j.gain = gain;
    j.src_chunk = src_chunk; j.pull = pull; j.result_every = result_every;
    j.out = out; j.out_cap = out_cap; j.results = results; j.results_cap = results_cap;
    if (run_pipeline(&j) != 0)
        return -1;
    if (out_bytes)
        *out_bytes = j.out_bytes;
    if (gain_rc)
        *gain_rc = j.gain_rc;
    return (long)j.n_results;
}

/* ---- CPU baseline: many streams over P threads ---------------------------------
 * in:  [n_streams][bytes_per_stream] contiguous; out: same shape or NULL;
 * scale[s], gain[s*channels + c]; last result of each stream into last[s] (or NULL).
 * Streams are split into contiguous ranges, one per thread.
 * Returns elapsed wall seconds of the threaded region (CLOCK_MONOTONIC). */
typedef struct bench_arg {
    const char *in;
    char *out;
    size_t bytes_per_stream;
    unsigned rate, channels;
    const uint16_t *scale;
    const uint16_t *gain;
    size_t s_begin, s_end;
    size_t pull;
    unsigned result_every;
    refh_result_t *last;
    int failed;
} bench_arg_t;

static void *bench_thread(void *p)
{
    bench_arg_t *a = p;
    size_t s;
    for (s = a->s_begin; s < a->s_end; s++) {
        pipeline_job_t j;
        memset(&j, 0, sizeof(j));
        j.in = a->in + s * a->bytes_per_stream;
        j.in_bytes = a->bytes_per_stream;
        j.rate = a->rate;
        j.channels = a->channels;
        j.set_gain = 1;
        j.gain_n = a->channels;
        j.scale = a->scale[s];
        j.gain = a->gain + s * a->channels;
        j.pull = a->pull;
        j.result_every = a->result_every;
        j.out = a->out ? a->out + s * a->bytes_per_stream : NULL;
        j.out_cap = a->bytes_per_stream;
        j.results = a->last ? &a->last[s] : NULL;
        j.results_cap = 1;
        j.last_only = 1;
        if (run_pipeline(&j) != 0)
            a->failed = 1;
    }
    return NULL;
}

double refh_bench(const void *in, void *out, size_t n_streams, size_t bytes_per_stream,
                  unsigned rate, unsigned channels, const uint16_t *scale, const uint16_t *gain,
                  size_t pull, unsigned result_every, unsigned n_threads, refh_result_t *last)
{
    pthread_t *tid;
    bench_arg_t *args;
    struct timespec t0, t1;
    unsigned t;
    int failed = 0;

    if (!n_threads)
        n_threads = 1;
    if (n_threads > n_streams)
        n_threads = (unsigned)n_streams;
    tid = calloc(n_threads, sizeof(*tid));
    args = calloc(n_threads, sizeof(*args));
    if (!tid || !args)
        return -1.0;

    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < n_threads; t++) {
        bench_arg_t *a = &args[t];
        a->in = in; a->out = out; a->bytes_per_stream = bytes_per_stream;
        a->rate = rate; a->channels = channels; a->scale = scale; a->gain = gain;
        a->s_begin = n_streams * t / n_threads;
        a->s_end = n_streams * (t + 1) / n_threads;
        a->pull = pull; a->result_every = result_every; a->last = last;
        pthread_create(&tid[t], NULL, bench_thread, a);
    }
    for (t = 0; t < n_threads; t++) {
        pthread_join(tid[t], NULL);
        failed |= args[t].failed;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(tid);
    free(args);
    if (failed)
        return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
