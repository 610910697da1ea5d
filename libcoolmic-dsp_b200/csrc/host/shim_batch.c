/* host/shim_batch.c -- many coolmic_transform_t / coolmic_vumeter_t objects on ONE batch engine.
 *
 * The reference runs one pull chain per stream, 1 KiB at a time, on one thread each
 * (src/simple.c:445-505). Here the objects keep their API but share a schedule: a driver calls
 * coolmic_b200_batch_tick(), which pulls up to block_frames from every member transform's input
 * handle straight into the pinned ring slot (whole frames only, the unfinished frame is carried
 * like transform.c:141-160), runs ONE fused transform+vumeter tick for all of them, and brings the
 * transformed PCM back. After that
 *   - every handle obtained from a member transform reads the stream's output with its own cursor
 *     (the fan-out tee.c:167-206 does with per-reader offsets, without the tee's copy), returning
 *     0 ("nothing now", iohandle.h:44-47) once it has caught up;
 *   - a fused vumeter needs no handle at all: the device metered the block in the same pass;
 *     coolmic_vumeter_read() reports the bytes metered since its last call and
 *     coolmic_vumeter_result() finalises the stream's device state (vumeter.c:189-218).
 * The meter window is therefore quantised to ticks: read(-1) ("an unspecified internal default",
 * vumeter.h:97) means "everything the last ticks metered".
 */
#include "shim_internal.h"

#include <stdlib.h>
#include <string.h>

typedef struct batch_cursor {
    unsigned stream;
    uint64_t epoch;            /* tick whose output the cursor is reading */
    size_t offset;
    struct batch_cursor *next;
} batch_cursor_t;

struct coolmic_b200_batch {
    shim_base_t base;
    cmgpu_ctx_t *ctx;
    unsigned channels, max_streams, block_frames;
    size_t stride, framesize;
    uint64_t epoch;                    /* number of ticks done */
    coolmic_transform_t **member;      /* [max_streams], not owned (members own the batch) */
    size_t *out_bytes;                 /* [max_streams] valid transformed bytes of the last tick */
    uint64_t *metered;                 /* [max_streams] bytes metered since the stream joined */
    uint32_t *frames;                  /* [max_streams] scratch for cmgpu_slot_set_frames */
    batch_cursor_t *cursors;
};

static void batch_destroy(shim_self_t self)
{
    coolmic_b200_batch_t *b = SHIM_SELF(self, coolmic_b200_batch_t);
    while (b->cursors) {
        batch_cursor_t *c = b->cursors;
        b->cursors = c->next;
        free(c);
    }
    if (b->ctx)
        cmgpu_ctx_destroy(b->ctx);
    free(b->member);
    free(b->out_bytes);
    free(b->metered);
    free(b->frames);
}

SHIM_TYPE(coolmic_b200_batch_t, batch_destroy);

coolmic_b200_batch_t *coolmic_b200_batch_new(int device, unsigned int channels, unsigned int max_streams,
                                             unsigned int block_frames)
{
    coolmic_b200_batch_t *b;
    if (!channels || channels > COOLMIC_B200_MAX_CHANNELS || !max_streams || !block_frames)
        return NULL;
    b = SHIM_NEW(coolmic_b200_batch_t, batch_destroy, "batch", SHIM_RO_NULL);
    if (!b)
        return NULL;
    b->channels = channels;
    b->max_streams = max_streams;
    b->block_frames = block_frames;
    b->framesize = 2u * channels;
    b->member = calloc(max_streams, sizeof(*b->member));
    b->out_bytes = calloc(max_streams, sizeof(*b->out_bytes));
    b->metered = calloc(max_streams, sizeof(*b->metered));
    b->frames = calloc(max_streams, sizeof(*b->frames));
    b->ctx = cmgpu_ctx_create(device < 0 ? shim_device() : device, channels, max_streams, 1, block_frames, 0);
    if (!b->member || !b->out_bytes || !b->metered || !b->frames || !b->ctx) {
        shim_unref(b);
        return NULL;
    }
    b->stride = cmgpu_block_stride(b->ctx);
    return b;
}

coolmic_transform_t *coolmic_b200_batch_transform_new(coolmic_b200_batch_t *b, const char *name,
                                                      coolmic_b200_ro_t associated, uint_least32_t rate)
{
    coolmic_transform_t *t;
    unsigned s;
    if (!b || !rate)
        return NULL;
    for (s = 0; s < b->max_streams && b->member[s]; s++)
        ;
    if (s == b->max_streams)
        return NULL;                                   /* batch is full */
    t = coolmic_transform_new(name, associated, rate, b->channels);
    if (!t)
        return NULL;
    shim_ref(b);
    t->batch = b;
    t->stream = s;
    b->member[s] = t;
    b->out_bytes[s] = 0;
    b->metered[s] = 0;
    cmgpu_stream_set_gain(b->ctx, s, 0, 0, NULL);      /* a fresh transform has no gain (transform.c:65-81) */
    cmgpu_meter_reset(b->ctx, s, 1);
    return t;
}

coolmic_vumeter_t *coolmic_b200_batch_vumeter_new(coolmic_b200_batch_t *b, coolmic_transform_t *of, const char *name,
                                                  coolmic_b200_ro_t associated)
{
    coolmic_vumeter_t *v;
    if (!b || !of || of->batch != b)
        return NULL;
    v = coolmic_vumeter_new(name, associated, of->rate, of->channels);
    if (!v)
        return NULL;
    shim_ref(b);
    v->batch = b;
    v->stream = of->stream;
    v->seen_bytes = b->metered[of->stream];
    return v;
}

size_t coolmic_b200_batch_pending(coolmic_b200_batch_t *b)
{
    size_t worst = 0;
    batch_cursor_t *c;
    if (!b)
        return 0;
    for (c = b->cursors; c; c = c->next) {
        const size_t have = b->out_bytes[c->stream];
        const size_t off = c->epoch == b->epoch ? c->offset : 0;
        if (have - off > worst)
            worst = have - off;
    }
    return worst;
}

int coolmic_b200_batch_tick(coolmic_b200_batch_t *b)
{
    unsigned char *slot;
    uint64_t before;
    long total = 0;
    unsigned s;

    if (!b)
        return COOLMIC_ERROR_FAULT;
    if (coolmic_b200_batch_pending(b))
        return COOLMIC_ERROR_BUSY;                     /* a reader has not caught up */
    slot = cmgpu_host_slot(b->ctx, 0);
    for (s = 0; s < b->max_streams; s++) {
        coolmic_transform_t *t = b->member[s];
        unsigned char *dst = slot + (size_t)s * b->stride;
        const size_t want = (size_t)b->block_frames * b->framesize;
        size_t have = 0, rest;
        b->frames[s] = 0;
        b->out_bytes[s] = 0;
        if (!t)
            continue;
        if (t->carry_fill) {
            memcpy(dst, t->carry, t->carry_fill);
            have = t->carry_fill;
            t->carry_fill = 0;
        }
        if (t->io) {
            ssize_t r = coolmic_iohandle_read(t->io, dst + have, want - have);
            if (r > 0)
                have += (size_t)r;
        }
        rest = have % b->framesize;
        if (rest) {
            memcpy(t->carry, dst + have - rest, rest);
            t->carry_fill = rest;
            have -= rest;
        }
        b->frames[s] = (uint32_t)(have / b->framesize);
        b->out_bytes[s] = have;
        b->metered[s] += have;
        total += (long)b->frames[s];
    }
    before = cmgpu_launch_count(b->ctx);
    if (cmgpu_slot_set_frames(b->ctx, 0, b->frames) != CMGPU_OK || cmgpu_submit(b->ctx, 0, NULL) != CMGPU_OK ||
        cmgpu_process(b->ctx, 0, CMGPU_FUSED) != CMGPU_OK || cmgpu_fetch(b->ctx, 0, NULL) != CMGPU_OK ||
        cmgpu_sync(b->ctx) != CMGPU_OK)
        return COOLMIC_ERROR_GENERIC;
    shim_count_launches(cmgpu_launch_count(b->ctx) - before);
    b->epoch++;
    return total > 0x7fffffffL ? 0x7fffffff : (int)total;
}

/* ---- hooks for the object code ------------------------------------------------------------ */
void *shim_batch_cursor_new(coolmic_b200_batch_t *b, unsigned stream)
{
    batch_cursor_t *c = calloc(1, sizeof(*c));
    if (!c)
        return NULL;
    c->stream = stream;
    c->epoch = b->epoch;
    c->offset = b->out_bytes[stream];                  /* a new reader starts at the next tick's output */
    c->next = b->cursors;
    b->cursors = c;
    return c;
}

void shim_batch_cursor_free(coolmic_b200_batch_t *b, void *cursor)
{
    batch_cursor_t **pp;
    for (pp = &b->cursors; *pp; pp = &(*pp)->next) {
        if (*pp == cursor) {
            *pp = (*pp)->next;
            free(cursor);
            return;
        }
    }
}

size_t shim_batch_cursor_unread(coolmic_b200_batch_t *b, void *cursor)
{
    batch_cursor_t *c = cursor;
    return b->out_bytes[c->stream] - (c->epoch == b->epoch ? c->offset : 0);
}

ssize_t shim_batch_read(coolmic_b200_batch_t *b, unsigned stream, void *cursor, void *buffer, size_t len)
{
    batch_cursor_t *c = cursor;
    size_t n;
    if (c->epoch != b->epoch) {
        c->epoch = b->epoch;
        c->offset = 0;
    }
    n = b->out_bytes[stream] - c->offset;
    if (n > len)
        n = len;
    if (!n)
        return 0;
    memcpy(buffer, (unsigned char *)cmgpu_host_slot(b->ctx, 0) + (size_t)stream * b->stride + c->offset, n);
    c->offset += n;
    return (ssize_t)n;
}

int shim_batch_set_gain(coolmic_b200_batch_t *b, unsigned stream, uint16_t scale, const uint16_t *gain)
{
    return cmgpu_stream_set_gain(b->ctx, stream, scale ? b->channels : 0, scale, gain) == CMGPU_OK ? 0 : -1;
}

void shim_batch_release(coolmic_b200_batch_t *b, unsigned stream)
{
    b->member[stream] = NULL;
    b->out_bytes[stream] = 0;
    shim_unref(b);
}

ssize_t shim_batch_vumeter_read(coolmic_b200_batch_t *b, coolmic_vumeter_t *v, ssize_t maxlen)
{
    uint64_t avail = b->metered[v->stream] - v->seen_bytes;
    if (maxlen >= 0 && avail > (uint64_t)maxlen)
        avail = (uint64_t)maxlen;
    v->seen_bytes += avail;
    return (ssize_t)avail;
}

int shim_batch_vumeter_result(coolmic_b200_batch_t *b, coolmic_vumeter_t *v, coolmic_vumeter_result_t *out)
{
    cmgpu_result_t res;
    unsigned c;
    int rc = cmgpu_meter_result(b->ctx, v->stream, (uint32_t)v->rate, &res);
    if (rc != CMGPU_OK)
        return rc == CMGPU_ERR_INVAL ? COOLMIC_ERROR_INVAL : COOLMIC_ERROR_GENERIC;
    memset(out, 0, sizeof(*out));
    out->rate = v->rate;
    out->channels = v->channels;
    out->frames = (size_t)res.frames;
    out->global_peak = res.global_peak;
    out->global_power = res.global_power;
    for (c = 0; c < v->channels; c++) {
        out->channel_peak[c] = res.channel_peak[c];
        out->channel_power[c] = res.channel_power[c];
    }
    return COOLMIC_ERROR_NONE;
}

int shim_batch_vumeter_reset(coolmic_b200_batch_t *b, coolmic_vumeter_t *v)
{
    return cmgpu_meter_reset(b->ctx, v->stream, 1) == CMGPU_OK ? COOLMIC_ERROR_NONE : COOLMIC_ERROR_GENERIC;
}
