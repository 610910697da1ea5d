/* tests/shim_harness.c -- drives the product's coolmic_* object shim (include/coolmic_b200_shim.h)
 * exactly the way oracle/ref_harness.c drives the reference's objects, so that the same Python
 * test code can run both and compare. Test code only. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#include "../include/coolmic_b200_shim.h"

typedef struct shimh_result {
    int32_t  rc;
    uint32_t rate;
    uint32_t channels;
    int32_t  global_peak;
    uint64_t frames;
    double   global_power;
    int32_t  channel_peak[16];
    double   channel_power[16];
} shimh_result_t;

static void flatten(shimh_result_t *dst, int rc, const coolmic_vumeter_result_t *src)
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

typedef struct memsrc {
    const char *data;
    size_t len, pos, chunk;
    int fail_at_end;        /* report -1 instead of 0 once exhausted */
} memsrc_t;

static ssize_t memsrc_read(void *userdata, void *buffer, size_t len)
{
    memsrc_t *m = userdata;
    size_t n = m->len - m->pos;
    if (!n && m->fail_at_end)
        return -1;
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

unsigned shimh_sizeof_result(void) { return (unsigned)sizeof(coolmic_vumeter_result_t); }

long shimh_transform(const void *in, size_t in_bytes, unsigned rate, unsigned channels,
                     int set_gain, unsigned gain_n, unsigned scale, const uint16_t *gain,
                     size_t src_chunk, size_t pull, void *out, size_t out_cap, int *gain_rc)
{
    memsrc_t mem = { in, in_bytes, 0, src_chunk, 0 };
    coolmic_iohandle_t *src, *h;
    coolmic_transform_t *tr;
    size_t done = 0;

    tr = coolmic_transform_new("tr", NULL, rate, channels);
    if (!tr)
        return -1;
    if (set_gain) {
        int rc = coolmic_transform_set_master_gain(tr, gain_n, (uint16_t)scale, gain);
        if (gain_rc)
            *gain_rc = rc;
    }
    src = coolmic_iohandle_new("memsrc", NULL, &mem, NULL, memsrc_read, memsrc_eof);
    coolmic_transform_attach_iohandle(tr, src);
    coolmic_b200_unref(src);
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
    coolmic_b200_unref(h);
    coolmic_b200_unref(tr);
    return (long)done;
}

long shimh_vumeter(const void *in, size_t in_bytes, unsigned rate, unsigned channels,
                   size_t src_chunk, long maxlen, unsigned result_every,
                   shimh_result_t *results, size_t results_cap)
{
    memsrc_t mem = { in, in_bytes, 0, src_chunk, 0 };
    coolmic_iohandle_t *src;
    coolmic_vumeter_t *vu;
    coolmic_vumeter_result_t res;
    size_t n = 0;
    unsigned good = 0;

    vu = coolmic_vumeter_new("vu", NULL, rate, channels);
    if (!vu)
        return -1;
    src = coolmic_iohandle_new("memsrc", NULL, &mem, NULL, memsrc_read, memsrc_eof);
    coolmic_vumeter_attach_iohandle(vu, src);
    coolmic_b200_unref(src);
    for (;;) {
        ssize_t r = coolmic_vumeter_read(vu, maxlen);
        if (r <= 0)
            break;
        if (result_every && ++good == result_every) {
            int rc = coolmic_vumeter_result(vu, &res);
            good = 0;
            if (n < results_cap)
                flatten(&results[n++], rc, &res);
        }
    }
    {
        int rc = coolmic_vumeter_result(vu, &res);
        if (n < results_cap)
            flatten(&results[n++], rc, &res);
    }
    coolmic_b200_unref(vu);
    return (long)n;
}

/* mem -> transform -> vumeter, the meter pulling straight from the transform's handle */
long shimh_chain(const void *in, size_t in_bytes, unsigned rate, unsigned channels,
                 int set_gain, unsigned gain_n, unsigned scale, const uint16_t *gain,
                 size_t src_chunk, long maxlen, unsigned result_every,
                 shimh_result_t *results, size_t results_cap)
{
    memsrc_t mem = { in, in_bytes, 0, src_chunk, 0 };
    coolmic_iohandle_t *src, *h;
    coolmic_transform_t *tr;
    coolmic_vumeter_t *vu;
    coolmic_vumeter_result_t res;
    size_t n = 0;
    unsigned good = 0;

    tr = coolmic_transform_new("tr", NULL, rate, channels);
    vu = coolmic_vumeter_new("vu", NULL, rate, channels);
    if (!tr || !vu)
        return -1;
    if (set_gain)
        coolmic_transform_set_master_gain(tr, gain_n, (uint16_t)scale, gain);
    src = coolmic_iohandle_new("memsrc", NULL, &mem, NULL, memsrc_read, memsrc_eof);
    coolmic_transform_attach_iohandle(tr, src);
    coolmic_b200_unref(src);
    h = coolmic_transform_get_iohandle(tr);
    coolmic_vumeter_attach_iohandle(vu, h);
    coolmic_b200_unref(h);
    for (;;) {
        ssize_t r = coolmic_vumeter_read(vu, maxlen);
        if (r <= 0)
            break;
        if (result_every && ++good == result_every) {
            int rc = coolmic_vumeter_result(vu, &res);
            good = 0;
            if (n < results_cap)
                flatten(&results[n++], rc, &res);
        }
    }
    {
        int rc = coolmic_vumeter_result(vu, &res);
        if (n < results_cap)
            flatten(&results[n++], rc, &res);
    }
    coolmic_b200_unref(vu);
    coolmic_b200_unref(tr);
    return (long)n;
}

/* argument checks that need no device */
int shimh_null_checks(void)
{
    coolmic_vumeter_result_t res;
    int bad = 0;
    bad |= coolmic_transform_new("t", NULL, 0, 2) != NULL;
    bad |= coolmic_transform_new("t", NULL, 48000, 0) != NULL;
    bad |= coolmic_transform_new("t", NULL, 48000, 17) != NULL;
    bad |= coolmic_vumeter_new("v", NULL, 0, 2) != NULL;
    bad |= coolmic_vumeter_new("v", NULL, 48000, 0) != NULL;
    bad |= coolmic_transform_attach_iohandle(NULL, NULL) != COOLMIC_ERROR_FAULT;
    bad |= coolmic_transform_set_master_gain(NULL, 1, 1, NULL) != COOLMIC_ERROR_FAULT;
    bad |= coolmic_vumeter_attach_iohandle(NULL, NULL) != COOLMIC_ERROR_FAULT;
    bad |= coolmic_vumeter_reset(NULL) != COOLMIC_ERROR_FAULT;
    bad |= coolmic_vumeter_read(NULL, -1) != -1;
    bad |= coolmic_vumeter_result(NULL, &res) != COOLMIC_ERROR_FAULT;
    bad |= coolmic_iohandle_new("h", NULL, NULL, NULL, NULL, NULL) != NULL;
    bad |= coolmic_iohandle_read(NULL, &res, 4) != COOLMIC_ERROR_FAULT;
    bad |= coolmic_iohandle_eof(NULL) != COOLMIC_ERROR_FAULT;
    {
        /* lifecycle without ever touching the device: new -> set gain -> INVAL case -> unref */
        coolmic_transform_t *t = coolmic_transform_new("t", NULL, 48000, 3);
        uint16_t g[3] = { 1, 2, 3 };
        coolmic_vumeter_t *v = coolmic_vumeter_new("v", NULL, 48000, 2);
        bad |= !t || !v;
        bad |= coolmic_transform_set_master_gain(t, 3, 2, g) != COOLMIC_ERROR_NONE;
        bad |= coolmic_transform_set_master_gain(t, 2, 2, g) != COOLMIC_ERROR_INVAL;
        bad |= coolmic_transform_set_master_gain(t, 1, 2, g) != COOLMIC_ERROR_NONE;
        bad |= coolmic_transform_set_master_gain(t, 0, 0, NULL) != COOLMIC_ERROR_NONE;
        bad |= coolmic_vumeter_result(v, &res) != COOLMIC_ERROR_INVAL;     /* no frames yet */
        bad |= coolmic_vumeter_reset(v) != COOLMIC_ERROR_NONE;
        coolmic_b200_unref(t);
        coolmic_b200_unref(v);
    }
    return bad;
}

/* ---- batch mode: n_streams member transforms with fused meters on one engine --------------- */
long shimh_batch(const void *in, size_t n_streams, size_t bytes_per_stream, unsigned rate, unsigned channels,
                 const uint16_t *scale, const uint16_t *gain, size_t src_chunk, unsigned block_frames,
                 unsigned result_every_ticks, size_t pull, void *out, size_t *out_bytes,
                 shimh_result_t *results, size_t cap, size_t *n_results, int *fanout_mismatch)
{
    coolmic_b200_batch_t *batch = coolmic_b200_batch_new(-1, channels, (unsigned)n_streams, block_frames);
    memsrc_t *mem = calloc(n_streams, sizeof(*mem));
    coolmic_transform_t **tr = calloc(n_streams, sizeof(*tr));
    coolmic_vumeter_t **vu = calloc(n_streams, sizeof(*vu));
    coolmic_iohandle_t **rd0 = calloc(n_streams, sizeof(*rd0)), **rd1 = calloc(n_streams, sizeof(*rd1));
    char *scratch = malloc(pull ? pull : 1024), *scratch2 = malloc(pull ? pull : 1024);
    long ticks = 0;
    size_t s;

    if (!batch || !mem || !tr || !vu || !rd0 || !rd1)
        return -1;
    if (!pull)
        pull = 1024;
    *fanout_mismatch = 0;
    for (s = 0; s < n_streams; s++) {
        coolmic_iohandle_t *src;
        mem[s].data = (const char *)in + s * bytes_per_stream;
        mem[s].len = bytes_per_stream;
        mem[s].chunk = src_chunk;
        tr[s] = coolmic_b200_batch_transform_new(batch, "tr", NULL, rate);
        if (!tr[s])
            return -2;
        coolmic_transform_set_master_gain(tr[s], channels, scale[s], gain + s * channels);
        src = coolmic_iohandle_new("memsrc", NULL, &mem[s], NULL, memsrc_read, memsrc_eof);
        coolmic_transform_attach_iohandle(tr[s], src);
        coolmic_b200_unref(src);
        rd0[s] = coolmic_transform_get_iohandle(tr[s]);       /* "the encoder" */
        rd1[s] = coolmic_transform_get_iohandle(tr[s]);       /* a second consumer of the same stream */
        vu[s] = coolmic_b200_batch_vumeter_new(batch, tr[s], "vu", NULL);
        out_bytes[s] = 0;
        n_results[s] = 0;
    }
    if (coolmic_b200_batch_new(-1, 0, 1, 1) != NULL)
        return -3;
    for (;;) {
        int frames = coolmic_b200_batch_tick(batch);
        if (frames < 0)
            return -10 + frames;
        if (frames == 0)
            break;
        ticks++;
        /* a second tick before the readers caught up must be refused */
        if (coolmic_b200_batch_tick(batch) != -12)
            return -4;
        for (s = 0; s < n_streams; s++) {
            coolmic_vumeter_result_t res;
            for (;;) {
                ssize_t r = coolmic_iohandle_read(rd0[s], scratch, pull);
                ssize_t r2;
                if (r <= 0)
                    break;
                r2 = coolmic_iohandle_read(rd1[s], scratch2, (size_t)r);
                if (r2 != r || memcmp(scratch, scratch2, (size_t)r) != 0)
                    *fanout_mismatch = 1;
                memcpy((char *)out + s * bytes_per_stream + out_bytes[s], scratch, (size_t)r);
                out_bytes[s] += (size_t)r;
            }
            coolmic_vumeter_read(vu[s], -1);
            if (result_every_ticks && ticks % result_every_ticks == 0) {
                int rc = coolmic_vumeter_result(vu[s], &res);
                if (n_results[s] < cap)
                    flatten(&results[s * cap + n_results[s]], rc, &res);
                n_results[s]++;
            }
        }
    }
    for (s = 0; s < n_streams; s++) {
        coolmic_vumeter_result_t res;
        int rc = coolmic_vumeter_result(vu[s], &res);
        if (n_results[s] < cap)
            flatten(&results[s * cap + n_results[s]], rc, &res);
        n_results[s]++;
        if (coolmic_iohandle_eof(rd0[s]) != 1)
            *fanout_mismatch |= 2;
        coolmic_b200_unref(rd0[s]);
        coolmic_b200_unref(rd1[s]);
        coolmic_b200_unref(vu[s]);
        coolmic_b200_unref(tr[s]);
    }
    coolmic_b200_unref(batch);
    free(mem); free(tr); free(vu); free(rd0); free(rd1); free(scratch); free(scratch2);
    return ticks;
}

/* ---- ring batch: the producer runs ahead of the readers by up to `slots` ticks ---------------- */
long shimh_batch_ring(const void *in, size_t n_streams, size_t bytes_per_stream, unsigned rate, unsigned channels,
                      const uint16_t *scale, const uint16_t *gain, size_t src_chunk, unsigned block_frames,
                      unsigned slots, unsigned threads, size_t pull, void *out, size_t *out_bytes,
                      shimh_result_t *results, int *flags)
{
    coolmic_b200_batch_t *batch = coolmic_b200_batch_new_ring(-1, channels, (unsigned)n_streams, block_frames, slots, threads);
    memsrc_t *mem = calloc(n_streams, sizeof(*mem));
    coolmic_transform_t **tr = calloc(n_streams, sizeof(*tr));
    coolmic_vumeter_t **vu = calloc(n_streams, sizeof(*vu));
    coolmic_iohandle_t **rd = calloc(n_streams, sizeof(*rd));
    coolmic_vumeter_result_t *res = calloc(n_streams, sizeof(*res));
    int *rcs = calloc(n_streams, sizeof(*rcs));
    char *scratch = malloc(pull ? pull : 1024);
    long ticks = 0;
    size_t s;

    if (!batch || !mem || !tr || !vu || !rd || !res || !rcs || !scratch)
        return -1;
    if (!pull)
        pull = 1024;
    *flags = 0;
    for (s = 0; s < n_streams; s++) {
        coolmic_iohandle_t *src;
        mem[s].data = (const char *)in + s * bytes_per_stream;
        mem[s].len = bytes_per_stream;
        mem[s].chunk = src_chunk;
        tr[s] = coolmic_b200_batch_transform_new(batch, "tr", NULL, rate);
        if (!tr[s])
            return -2;
        coolmic_transform_set_master_gain(tr[s], channels, scale[s], gain + s * channels);
        src = coolmic_iohandle_new("memsrc", NULL, &mem[s], NULL, memsrc_read, memsrc_eof);
        coolmic_transform_attach_iohandle(tr[s], src);
        coolmic_b200_unref(src);
        rd[s] = coolmic_transform_get_iohandle(tr[s]);
        vu[s] = coolmic_b200_batch_vumeter_new(batch, tr[s], "vu", NULL);
        out_bytes[s] = 0;
    }
    for (;;) {
        int fr, any = 0;
        unsigned issued = 0;
        /* producer: as many ticks as the ring takes; it must refuse once `slots` ticks are unread */
        while ((fr = coolmic_b200_batch_tick(batch)) > 0) {
            ticks++;
            issued++;
        }
        if (fr != 0 && fr != -12)
            return -10 + fr;
        if (issued > slots)
            *flags |= 4;
        if (fr == -12 && coolmic_b200_batch_pending(batch) == 0)
            *flags |= 8;                /* BUSY without anything pending */
        /* consumers: everything that is there, in stream order (each read may wait for its slot) */
        for (s = 0; s < n_streams; s++) {
            for (;;) {
                ssize_t r = coolmic_iohandle_read(rd[s], scratch, pull);
                if (r < 0)
                    return -5;
                if (r == 0)
                    break;
                any = 1;
                memcpy((char *)out + s * bytes_per_stream + out_bytes[s], scratch, (size_t)r);
                out_bytes[s] += (size_t)r;
            }
        }
        if (fr == 0 && !any)
            break;
    }
    if (coolmic_b200_batch_pending(batch) != 0)
        *flags |= 16;
    if (coolmic_b200_batch_results(batch, res, rcs) != 0)
        return -6;
    for (s = 0; s < n_streams; s++) {
        flatten(&results[s], rcs[s], &res[s]);
        if (coolmic_iohandle_eof(rd[s]) != 1)
            *flags |= 2;
        coolmic_b200_unref(rd[s]);
        coolmic_b200_unref(vu[s]);
        coolmic_b200_unref(tr[s]);
    }
    coolmic_b200_unref(batch);
    free(mem); free(tr); free(vu); free(rd); free(res); free(rcs); free(scratch);
    return ticks;
}
