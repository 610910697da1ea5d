/* oracle/coolmic_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the libcoolmic-dsp transform + vumeter hot path; see
 * coolmic_oracle.h. Written from the behaviour described in SURVEY.md section 8a /
 * Appendix A, each function citing the reference lines it restates. Validated against
 * the reference's own object code by tests/test_oracle.py.
 */
#include "coolmic_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------ */
/* transform.c:101-124. Per sample: widen, multiply by the channel's gain, divide by
 * the scale with C's truncation toward zero, saturate to int16. scale == 0 leaves the
 * buffer untouched (transform.c:107-108). */
void oracle_gain_process(int16_t *samples, size_t frames, unsigned channels,
                         uint16_t scale, const uint16_t *gain)
{
    size_t total, i;

    if (scale == 0)
        return;

    total = frames * channels;
    for (i = 0; i < total; i++) {
        int64_t v = (int64_t)samples[i] * (int64_t)gain[i % channels];
        v = v / (int64_t)scale;                 /* C99 6.5.5p6: truncates toward zero */
        if (v > 32767)
            v = 32767;
        if (v < -32768)
            v = -32768;
        samples[i] = (int16_t)v;
    }
}

/* transform.c:195-222. Four cases: disable, exact-width copy, mono broadcast, stereo
 * setting folded onto a mono signal (u32 average); anything else is INVAL and the
 * previous state is kept. */
int oracle_gain_adapt(unsigned stream_channels, unsigned n, uint16_t scale, const uint16_t *gain,
                      uint16_t *state_scale, uint16_t state_gain[ORACLE_MAX_CHANNELS])
{
    unsigned c;

    if (n == 0 || scale == 0 || gain == NULL) {
        *state_scale = 0;
        return ORACLE_ERROR_NONE;
    }
    if (n == stream_channels) {
        for (c = 0; c < n; c++)
            state_gain[c] = gain[c];
    } else if (n == 1) {
        for (c = 0; c < stream_channels; c++)
            state_gain[c] = gain[0];
    } else if (n == 2 && stream_channels == 1) {
        state_gain[0] = (uint16_t)(((uint32_t)gain[0] + (uint32_t)gain[1]) / 2u);
    } else {
        return ORACLE_ERROR_INVAL;
    }
    *state_scale = scale;
    return ORACLE_ERROR_NONE;
}

/* vumeter.c:161-177. abs() acts on the int-promoted sample, so |-32768| = 32768 beats
 * 32767; comparisons are strict, so the first sample reaching a magnitude keeps the
 * peak. The global peak is only examined when the channel peak moves (vumeter.c:163-168). */
void oracle_meter_accumulate(oracle_meter_t *m, const int16_t *samples, size_t frames, unsigned channels)
{
    size_t f;
    unsigned c;

    for (f = 0; f < frames; f++) {
        for (c = 0; c < channels; c++) {
            int x = samples[f * channels + c];
            int mag = x < 0 ? -x : x;
            int cur = m->channel_peak[c];
            if (mag > (cur < 0 ? -cur : cur)) {
                int g = m->global_peak;
                m->channel_peak[c] = (int16_t)x;
                if (mag > (g < 0 ? -g : g))
                    m->global_peak = (int16_t)x;
            }
            m->power[c] += (int64_t)x * (int64_t)x;
        }
    }
    m->frames += frames;
}

static double power_to_db(double mean_square)
{
    /* vumeter.c:204-205 / 210-211 */
    double p = 20. * log10(sqrt(mean_square) / 32768.);
    return fmin(p, 0.);
}

/* vumeter.c:189-218. Integer division happens BEFORE the conversion to double, signed for
 * the per-channel mean (vumeter.c:203) and unsigned for the global one (vumeter.c:209). */
int oracle_meter_finalise(oracle_meter_t *m, uint32_t rate, unsigned channels, oracle_result_t *out)
{
    unsigned c;
    int64_t all = 0;

    if (!m || !out)
        return ORACLE_ERROR_FAULT;
    if (m->frames == 0)
        return ORACLE_ERROR_INVAL;

    memset(out, 0, sizeof(*out));
    out->rc = ORACLE_ERROR_NONE;
    out->rate = rate;
    out->channels = channels;
    out->frames = m->frames;
    out->global_peak = m->global_peak;
    for (c = 0; c < channels; c++) {
        all += m->power[c];
        out->channel_peak[c] = m->channel_peak[c];
        out->channel_power[c] = power_to_db((double)(m->power[c] / (int64_t)m->frames));
    }
    out->global_power = power_to_db((double)((uint64_t)all / (uint64_t)(m->frames * (uint64_t)channels)));

    memset(m, 0, sizeof(*m));
    return ORACLE_ERROR_NONE;
}

/* ------------------------------------------------------------------------------ */
/* iohandle.c:74-104: keep calling the read callback until the request is filled, the
 * callback yields 0, or it fails; a failure after progress reports the progress. */
static ssize_t read_fully(oracle_read_cb cb, void *userdata, void *buffer, size_t len)
{
    size_t done = 0;

    if (!cb || !buffer)
        return ORACLE_ERROR_FAULT;
    while (done < len) {
        ssize_t r = cb(userdata, (char *)buffer + done, len - done);
        if (r < 0)
            return done ? (ssize_t)done : r;
        if (r == 0)
            break;
        done += (size_t)r;
    }
    return (ssize_t)done;
}

void oracle_transform_init(oracle_transform_t *t, unsigned channels, oracle_read_cb src, void *userdata)
{
    memset(t, 0, sizeof(*t));
    t->channels = channels;
    t->src = src;
    t->src_userdata = userdata;
}

/* transform.c:126-165. Requests are cut down to whole frames; bytes of a frame the source
 * has not finished yet are parked in `carry` and prepended to the next request. */
ssize_t oracle_transform_read(oracle_transform_t *t, void *buffer, size_t len)
{
    const size_t framesize = 2u * t->channels;
    size_t have = 0, rest;
    ssize_t r;

    len -= len % framesize;
    if (len == 0)
        return 0;

    if (t->carry_fill) {
        memcpy(buffer, t->carry, t->carry_fill);
        have = t->carry_fill;
        t->carry_fill = 0;
    }

    r = read_fully(t->src, t->src_userdata, (char *)buffer + have, len - have);
    if (r > 0)
        have += (size_t)r;

    rest = have % framesize;
    if (rest) {
        memcpy(t->carry, (char *)buffer + have - rest, rest);
        t->carry_fill = rest;
        have -= rest;
    }

    oracle_gain_process(buffer, have / framesize, t->channels, t->scale, t->gain);
    return (ssize_t)have;
}

void oracle_vumeter_init(oracle_vumeter_t *v, uint32_t rate, unsigned channels, oracle_read_cb src, void *userdata)
{
    memset(v, 0, sizeof(*v));
    v->rate = rate;
    v->channels = channels;
    v->src = src;
    v->src_userdata = userdata;
}

/* vumeter.c:112-187. One physical read of at most (1024 - fill) bytes, capped by maxlen when
 * maxlen >= 0; whole frames are metered, a trailing partial frame stays at the buffer's head.
 * The return value is what the physical read delivered (0 when it failed with data pending). */
ssize_t oracle_vumeter_read(oracle_vumeter_t *v, ssize_t maxlen)
{
    size_t want = sizeof(v->buffer) - v->fill;
    size_t framesize = 2u * v->channels;
    size_t frames, used;
    ssize_t r;

    if (maxlen >= 0 && want > (size_t)maxlen)
        want = (size_t)maxlen;

    /* iohandle.c:81-82: a zero-length request is a successful read of 0 bytes */
    r = want ? read_fully(v->src, v->src_userdata, v->buffer + v->fill, want) : 0;
    if (r < 0) {
        /* vumeter.c:127-131 tests for exactly -1; other negatives (e.g. FAULT from a missing
         * handle) would be added to the fill level -- undefined territory we do not restate. */
        if (v->fill == 0)
            return -1;
        r = 0;
    } else {
        v->fill += (size_t)r;
    }

    frames = v->fill / framesize;
    used = frames * framesize;
    oracle_meter_accumulate(&v->meter, (const int16_t *)(const void *)v->buffer, frames, v->channels);
    if (used < v->fill)
        memmove(v->buffer, v->buffer + used, v->fill - used);
    v->fill -= used;
    return r;
}

int oracle_vumeter_result(oracle_vumeter_t *v, oracle_result_t *out)
{
    if (!v || !out)
        return ORACLE_ERROR_FAULT;
    return oracle_meter_finalise(&v->meter, v->rate, v->channels, out);
}

/* ------------------------------------------------------------------------------ */
typedef struct memsrc {
    const unsigned char *data;
    size_t len, pos, chunk;
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

static ssize_t transform_as_source(void *userdata, void *buffer, size_t len)
{
    return oracle_transform_read(userdata, buffer, len);
}

long oracle_run_transform(const void *in, size_t in_bytes, unsigned channels,
                          int set_gain, unsigned gain_n, unsigned scale, const uint16_t *gain,
                          size_t src_chunk, size_t pull, void *out, size_t out_cap, int *gain_rc)
{
    memsrc_t mem = { in, in_bytes, 0, src_chunk };
    oracle_transform_t t;
    size_t done = 0;

    if (!channels)
        return -1;
    oracle_transform_init(&t, channels, memsrc_read, &mem);
    if (set_gain) {
        int rc = oracle_gain_adapt(channels, gain_n, (uint16_t)scale, gain, &t.scale, t.gain);
        if (gain_rc)
            *gain_rc = rc;
    }
    if (!pull)
        pull = 1024;
    while (done < out_cap) {
        size_t want = out_cap - done < pull ? out_cap - done : pull;
        /* the consumer reads through an iohandle, i.e. through the same retry loop */
        ssize_t r = read_fully(transform_as_source, &t, (char *)out + done, want);
        if (r <= 0)
            break;
        done += (size_t)r;
    }
    return (long)done;
}

long oracle_run_vumeter(const void *in, size_t in_bytes, uint32_t rate, unsigned channels,
                        size_t src_chunk, long maxlen, unsigned result_every,
                        oracle_result_t *results, size_t results_cap)
{
    memsrc_t mem = { in, in_bytes, 0, src_chunk };
    oracle_vumeter_t v;
    oracle_result_t res;
    size_t n = 0;
    unsigned good = 0;

    if (!channels || !rate)
        return -1;
    oracle_vumeter_init(&v, rate, channels, memsrc_read, &mem);
    for (;;) {
        ssize_t r = oracle_vumeter_read(&v, maxlen);
        if (r <= 0)
            break;
        if (result_every && ++good == result_every) {
            int rc = oracle_vumeter_result(&v, &res);
            good = 0;
            res.rc = rc;
            if (n < results_cap)
                results[n++] = res;
        }
    }
    {
        int rc = oracle_vumeter_result(&v, &res);
        if (rc != ORACLE_ERROR_NONE)
            memset(&res, 0, sizeof(res));
        res.rc = rc;
        if (n < results_cap)
            results[n++] = res;
    }
    return (long)n;
}

/* ------------------------------------------------------------------------------ */
void oracle_batch(int16_t *pcm, size_t n_streams, size_t stride_samples, const uint32_t *frames,
                  unsigned channels, const uint16_t *scale, const uint16_t *gain,
                  oracle_meter_t *meters)
{
    size_t s;
    for (s = 0; s < n_streams; s++) {
        int16_t *p = pcm + s * stride_samples;
        oracle_gain_process(p, frames[s], channels, scale[s], gain + s * channels);
        if (meters)
            oracle_meter_accumulate(&meters[s], p, frames[s], channels);
    }
}

/* enc_vorbis.c:108-117: de-interleave, one division by 32768.f per sample (binary32) */
void oracle_planar(const int16_t *in, size_t frames, unsigned channels, float *planes, size_t plane_stride)
{
    size_t f;
    unsigned c;
    for (f = 0; f < frames; f++)
        for (c = 0; c < channels; c++)
            planes[(size_t)c * plane_stride + f] = in[f * channels + c] / 32768.f;
}

/* EXTENSION CHECKER: see coolmic_oracle.h. No reference lines to cite for the mix itself. */
void oracle_mix_process(const int16_t *in, size_t frames, unsigned cin, unsigned cout, uint16_t scale,
                        const uint16_t *w, int16_t *out, oracle_meter_t *min, oracle_meter_t *mout)
{
    size_t f;
    unsigned m, c;

    for (f = 0; f < frames; f++) {
        for (m = 0; m < cout; m++) {
            int64_t acc = 0;
            for (c = 0; c < cin; c++)
                acc += (int64_t)in[f * cin + c] * (int64_t)w[m * cin + c];
            acc = acc / (int64_t)scale;
            if (acc > 32767)
                acc = 32767;
            if (acc < -32768)
                acc = -32768;
            out[f * cout + m] = (int16_t)acc;
        }
    }
    if (min)
        oracle_meter_accumulate(min, in, frames, cin);
    if (mout)
        oracle_meter_accumulate(mout, out, frames, cout);
}

typedef struct batch_arg {
    int16_t *pcm;
    size_t begin, end, stride;
    const uint32_t *frames;
    unsigned channels;
    const uint16_t *scale, *gain;
    oracle_meter_t *meters;
} batch_arg_t;

static void *batch_thread(void *p)
{
    batch_arg_t *a = p;
    oracle_batch(a->pcm + a->begin * a->stride, a->end - a->begin, a->stride, a->frames + a->begin,
                 a->channels, a->scale + a->begin, a->gain + a->begin * a->channels,
                 a->meters ? a->meters + a->begin : NULL);
    return NULL;
}

double oracle_batch_threads(int16_t *pcm, size_t n_streams, size_t stride_samples, const uint32_t *frames,
                            unsigned channels, const uint16_t *scale, const uint16_t *gain,
                            oracle_meter_t *meters, unsigned n_threads)
{
    pthread_t *tid;
    batch_arg_t *args;
    struct timespec t0, t1;
    unsigned t;

    if (!n_threads)
        n_threads = 1;
    if (n_threads > n_streams)
        n_threads = (unsigned)(n_streams ? n_streams : 1);
    tid = calloc(n_threads, sizeof(*tid));
    args = calloc(n_threads, sizeof(*args));
    if (!tid || !args)
        return -1.0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (t = 0; t < n_threads; t++) {
        batch_arg_t *a = &args[t];
        a->pcm = pcm; a->stride = stride_samples; a->frames = frames; a->channels = channels;
        a->scale = scale; a->gain = gain; a->meters = meters;
        a->begin = n_streams * t / n_threads;
        a->end = n_streams * (t + 1) / n_threads;
        pthread_create(&tid[t], NULL, batch_thread, a);
    }
    for (t = 0; t < n_threads; t++)
        pthread_join(tid[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(tid);
    free(args);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

uint64_t oracle_fnv1a64(const void *data, size_t len)
{
    const unsigned char *p = data;
    uint64_t h = 1469598103934665603ull;
    size_t i;
    for (i = 0; i < len; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}
