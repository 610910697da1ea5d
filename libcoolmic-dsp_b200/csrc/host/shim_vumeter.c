/* host/shim_vumeter.c -- coolmic_vumeter_* on the GPU.
 *
 * Same contract as reference src/vumeter.c: read() pulls at most (1024 - pending) bytes, or
 * `maxlen` if that is smaller and non-negative (vumeter.c:112-136), meters every whole frame it
 * then holds and keeps the rest for next time (vumeter.c:159-184); result() needs at least one
 * frame, converts to dB, copies the 192-byte result and resets (vumeter.c:189-218).
 * Metering = one meter-only tick of a private one-stream cmgpu context per read; the integer
 * state lives on the device between reads and is finalised on the host with the reference's
 * expression (cmgpu_finalise), so peaks, frame counts and dB values are bit-identical.
 */
#include "shim_internal.h"

#include <string.h>

static void vumeter_destroy(shim_self_t self)
{
    coolmic_vumeter_t *v = SHIM_SELF(self, coolmic_vumeter_t);
    shim_unref(v->in);
    if (v->ctx)
        cmgpu_ctx_destroy(v->ctx);
    if (v->batch)
        shim_unref(v->batch);
}

SHIM_TYPE(coolmic_vumeter_t, vumeter_destroy);

coolmic_vumeter_t *coolmic_vumeter_new(const char *name, coolmic_b200_ro_t associated,
                                       uint_least32_t rate, unsigned int channels)
{
    coolmic_vumeter_t *v;
    if (!rate || !channels || channels > COOLMIC_B200_MAX_CHANNELS)
        return NULL;
    v = SHIM_NEW(coolmic_vumeter_t, vumeter_destroy, name, associated);
    if (!v)
        return NULL;
    v->rate = rate;
    v->channels = channels;
    return v;
}

int coolmic_vumeter_reset(coolmic_vumeter_t *self)
{
    if (!self)
        return COOLMIC_ERROR_FAULT;
    if (self->batch)
        return shim_batch_vumeter_reset(self->batch, self);
    if (self->ctx && cmgpu_meter_reset(self->ctx, 0, 1) != CMGPU_OK)
        return COOLMIC_ERROR_GENERIC;
    return COOLMIC_ERROR_NONE;
}

int coolmic_vumeter_attach_iohandle(coolmic_vumeter_t *self, coolmic_iohandle_t *handle)
{
    if (!self)
        return COOLMIC_ERROR_FAULT;
    shim_unref(self->in);
    self->in = handle;
    shim_ref(handle);
    return COOLMIC_ERROR_NONE;
}

static int vumeter_meter_frames(coolmic_vumeter_t *v, uint32_t frames)
{
    uint64_t before;
    if (!frames)
        return 0;
    if (!v->ctx) {
        v->ctx = cmgpu_ctx_create(shim_device(), v->channels, 1, 1, SHIM_VU_BUFFER / (2u * v->channels), 0);
        if (!v->ctx)
            return -1;
    }
    before = cmgpu_launch_count(v->ctx);
    memcpy(cmgpu_host_slot(v->ctx, 0), v->buffer, (size_t)frames * 2u * v->channels);
    if (cmgpu_slot_set_frames(v->ctx, 0, &frames) != CMGPU_OK || cmgpu_submit(v->ctx, 0, NULL) != CMGPU_OK ||
        cmgpu_process(v->ctx, 0, CMGPU_METER) != CMGPU_OK || cmgpu_sync(v->ctx) != CMGPU_OK)
        return -1;
    shim_count_launches(cmgpu_launch_count(v->ctx) - before);
    return 0;
}

ssize_t coolmic_vumeter_read(coolmic_vumeter_t *self, ssize_t maxlen)
{
    size_t want, framesize, frames, used;
    ssize_t r;

    if (!self)
        return -1;
    if (self->batch)
        return shim_batch_vumeter_read(self->batch, self, maxlen);
    want = sizeof(self->buffer) - self->fill;
    if (maxlen >= 0 && want > (size_t)maxlen)
        want = (size_t)maxlen;

    r = coolmic_iohandle_read(self->in, self->buffer + self->fill, want);
    if (r < 0) {
        /* vumeter.c:127-131: an error with nothing pending is the caller's error; with a
         * partial frame pending it is "nothing now". The reference only tests for -1 and would
         * add other negatives (e.g. -9 for a missing handle) to its fill level; we do not. */
        if (!self->fill)
            return -1;
        r = 0;
    } else {
        self->fill += (size_t)r;
    }

    framesize = 2u * self->channels;
    frames = self->fill / framesize;
    used = frames * framesize;
    if (vumeter_meter_frames(self, (uint32_t)frames) != 0)
        return -1;
    if (used < self->fill)
        memmove(self->buffer, self->buffer + used, self->fill - used);
    self->fill -= used;
    return r;
}

int coolmic_vumeter_result(coolmic_vumeter_t *self, coolmic_vumeter_result_t *result)
{
    cmgpu_result_t res;
    unsigned int c;
    int rc;

    if (!self || !result)
        return COOLMIC_ERROR_FAULT;
    if (self->batch)
        return shim_batch_vumeter_result(self->batch, self, result);
    if (!self->ctx)
        return COOLMIC_ERROR_INVAL;     /* nothing metered yet: frames == 0 (vumeter.c:198-199) */
    rc = cmgpu_meter_result(self->ctx, 0, (uint32_t)self->rate, &res);
    if (rc != CMGPU_OK)
        return rc == CMGPU_ERR_INVAL ? COOLMIC_ERROR_INVAL : COOLMIC_ERROR_GENERIC;
    memset(result, 0, sizeof(*result));
    result->rate = self->rate;
    result->channels = self->channels;
    result->frames = (size_t)res.frames;
    result->global_peak = res.global_peak;
    result->global_power = res.global_power;
    for (c = 0; c < self->channels; c++) {
        result->channel_peak[c] = res.channel_peak[c];
        result->channel_power[c] = res.channel_power[c];
    }
    return COOLMIC_ERROR_NONE;
}
