/* host/shim_transform.c -- coolmic_transform_* on the GPU.
 *
 * Same contract as reference src/transform.c: a pull stage that hands out whole frames only,
 * carries the bytes of an unfinished frame to the next read (transform.c:126-165), applies the
 * per-channel master gain (transform.c:101-124) and adapts gain vectors of a different width
 * (transform.c:195-222). The gain itself runs as one tick of a private one-stream cmgpu context:
 * pulled bytes -> pinned slot -> H2D -> fused_tick (transform only) -> D2H -> caller's buffer.
 */
#include "shim_internal.h"

#include <stdlib.h>
#include <string.h>

static void transform_destroy(shim_self_t self)
{
    coolmic_transform_t *t = SHIM_SELF(self, coolmic_transform_t);
    shim_unref(t->io);
    if (t->ctx)
        cmgpu_ctx_destroy(t->ctx);
    if (t->batch)
        shim_batch_release(t->batch, t->stream);
}

SHIM_TYPE(coolmic_transform_t, transform_destroy);

coolmic_transform_t *coolmic_transform_new(const char *name, coolmic_b200_ro_t associated,
                                           uint_least32_t rate, unsigned int channels)
{
    coolmic_transform_t *t;
    /* the reference indexes gain[16] with `channels` unchecked (transform.c:69-70); refuse instead */
    if (!rate || !channels || channels > COOLMIC_B200_MAX_CHANNELS)
        return NULL;
    t = SHIM_NEW(coolmic_transform_t, transform_destroy, name, associated);
    if (!t)
        return NULL;
    t->rate = rate;
    t->channels = channels;
    t->block_frames = SHIM_BLOCK_BYTES / (2u * channels);
    return t;
}

int coolmic_transform_attach_iohandle(coolmic_transform_t *self, coolmic_iohandle_t *handle)
{
    if (!self)
        return COOLMIC_ERROR_FAULT;
    shim_unref(self->io);               /* NULL-safe, like igloo_ro_unref */
    self->io = handle;
    shim_ref(handle);
    return COOLMIC_ERROR_NONE;
}

int coolmic_transform_set_master_gain(coolmic_transform_t *self, unsigned int channels, uint16_t scale,
                                      const uint16_t *gain)
{
    unsigned int c;
    if (!self)
        return COOLMIC_ERROR_FAULT;
    if (!channels || !scale || !gain) {
        self->gain_scale = 0;
    } else if (channels == self->channels) {
        memcpy(self->gain, gain, sizeof(*gain) * channels);
        self->gain_scale = scale;
    } else if (channels == 1) {
        for (c = 0; c < self->channels; c++)
            self->gain[c] = gain[0];
        self->gain_scale = scale;
    } else if (channels == 2 && self->channels == 1) {
        self->gain[0] = (uint16_t)(((uint32_t)gain[0] + (uint32_t)gain[1]) / 2u);
        self->gain_scale = scale;
    } else {
        return COOLMIC_ERROR_INVAL;     /* previous setting stays in force */
    }
    self->gain_dirty = 1;
    if (self->batch)
        return shim_batch_set_gain(self->batch, self->stream, self->gain_scale, self->gain) == 0
                   ? COOLMIC_ERROR_NONE : COOLMIC_ERROR_GENERIC;
    return COOLMIC_ERROR_NONE;
}

static int transform_device_ready(coolmic_transform_t *t)
{
    if (!t->ctx) {
        t->ctx = cmgpu_ctx_create(shim_device(), t->channels, 1, 1, t->block_frames, 0);
        if (!t->ctx)
            return -1;
        t->gain_dirty = 1;
    }
    if (t->gain_dirty) {
        if (cmgpu_stream_set_gain(t->ctx, 0, t->gain_scale ? t->channels : 0, t->gain_scale, t->gain) != CMGPU_OK)
            return -1;
        t->gain_dirty = 0;
    }
    return 0;
}

/* whole frames in `buffer`, in place, on the GPU */
static int transform_process(coolmic_transform_t *t, void *buffer, size_t frames)
{
    const size_t framesize = 2u * t->channels;
    if (!frames || !t->gain_scale)      /* transform.c:107-108: no gain set, nothing to do */
        return 0;
    if (transform_device_ready(t) != 0)
        return -1;
    while (frames) {
        uint32_t n = frames > t->block_frames ? t->block_frames : (uint32_t)frames;
        uint64_t before = cmgpu_launch_count(t->ctx);
        void *slot = cmgpu_host_slot(t->ctx, 0);
        memcpy(slot, buffer, n * framesize);
        if (cmgpu_slot_set_frames(t->ctx, 0, &n) != CMGPU_OK || cmgpu_submit(t->ctx, 0, NULL) != CMGPU_OK ||
            cmgpu_process(t->ctx, 0, CMGPU_TRANSFORM) != CMGPU_OK || cmgpu_fetch(t->ctx, 0, NULL) != CMGPU_OK ||
            cmgpu_sync(t->ctx) != CMGPU_OK)
            return -1;
        memcpy(buffer, slot, n * framesize);
        shim_count_launches(cmgpu_launch_count(t->ctx) - before);
        buffer = (char *)buffer + n * framesize;
        frames -= n;
    }
    return 0;
}

/* One handle = one reader. In a batch every reader has its own cursor into the stream's output,
 * which is what makes the handles tee-compatible (tee.c:167-206 keeps one offset per reader). */
typedef struct transform_reader {
    coolmic_transform_t *t;
    void *cursor;
} transform_reader_t;

static ssize_t transform_read(void *userdata, void *buffer, size_t len)
{
    transform_reader_t *rd = userdata;
    coolmic_transform_t *t = rd->t;
    const size_t framesize = 2u * t->channels;
    size_t have = 0, rest;
    ssize_t r;

    len -= len % framesize;
    if (!len)
        return 0;
    if (t->batch)                       /* whole frames the last ticks left for this reader */
        return shim_batch_read(t->batch, t->stream, rd->cursor, buffer, len);
    if (t->carry_fill) {                /* an unfinished frame from last time goes first */
        memcpy(buffer, t->carry, t->carry_fill);
        have = t->carry_fill;
        t->carry_fill = 0;
    }
    r = coolmic_iohandle_read(t->io, (char *)buffer + have, len - have);
    if (r > 0)
        have += (size_t)r;
    rest = have % framesize;
    if (rest) {
        memcpy(t->carry, (char *)buffer + have - rest, rest);
        t->carry_fill = rest;
        have -= rest;
    }
    if (transform_process(t, buffer, have / framesize) != 0)
        return -1;                      /* any CUDA failure is a read error; no CPU fallback */
    return (ssize_t)have;
}

static int transform_eof(void *userdata)
{
    transform_reader_t *rd = userdata;
    coolmic_transform_t *t = rd->t;
    /* like tee.c:208-217: data still waiting for this reader is not EOF */
    if (t->batch && shim_batch_cursor_unread(t->batch, rd->cursor))
        return 0;
    if (!t->io)
        return 1;
    return coolmic_iohandle_eof(t->io);
}

static int transform_release(void *userdata)
{
    transform_reader_t *rd = userdata;
    coolmic_transform_t *t = rd->t;
    if (t->batch && rd->cursor)
        shim_batch_cursor_free(t->batch, rd->cursor);
    free(rd);
    return shim_unref(t);
}

coolmic_iohandle_t *coolmic_transform_get_iohandle(coolmic_transform_t *self)
{
    coolmic_iohandle_t *h;
    transform_reader_t *rd;
    if (shim_ref(self) != COOLMIC_ERROR_NONE)
        return NULL;
    rd = calloc(1, sizeof(*rd));
    if (rd) {
        rd->t = self;
        rd->cursor = self->batch ? shim_batch_cursor_new(self->batch, self->stream) : NULL;
    }
    h = rd && (!self->batch || rd->cursor)
            ? coolmic_iohandle_new(NULL, SHIM_RO_NULL, rd, transform_release, transform_read, transform_eof) : NULL;
    if (!h) {
        if (rd && rd->cursor)
            shim_batch_cursor_free(self->batch, rd->cursor);
        free(rd);
        shim_unref(self);
    }
    return h;
}
