/* tests/stub/cmgpu_stub.c -- TEST INFRASTRUCTURE: a CPU stand-in for the 16 cmgpu_* entry points the
 * host shim (libcoolmic-dsp_b200/csrc/host/*.c) calls, so that the shim's HOST LOGIC -- framing and
 * carry, reader cursors, the ring batch's slot reuse and back-pressure, its threads, reference
 * counting -- can run here without a GPU, and under ASan / UBSan / TSan.
 * The arithmetic comes from the oracle port (oracle/coolmic_oracle.c). The product never links this.
 *
 * Asynchrony is imitated where it matters for catching ordering bugs: cmgpu_fetch only marks a
 * download as pending; the transformed PCM reaches the pinned slot when cmgpu_slot_wait / cmgpu_sync
 * "wait" for it, and until then the slot still shows the untransformed input (poisoned when the gain
 * would not change it), so a reader that does not wait is caught.
 */
#include "../../include/cmgpu.h"
#include "../../oracle/coolmic_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct cmgpu_ctx {
    unsigned channels, max_streams, slots, block_frames;
    size_t stride, slot_bytes;
    unsigned char *host, *dev;
    uint32_t *frames;             /* [slots][max_streams] */
    unsigned char *has_frames, *down_pending;
    uint16_t *scale, *gain;
    oracle_meter_t *meters;
    uint64_t launches;
    pthread_mutex_t mu;
};

static __thread char g_err[128];
const char *cmgpu_last_error(void) { return g_err; }
int cmgpu_device_count(void) { return 1; }

cmgpu_ctx_t *cmgpu_ctx_create(int device, unsigned channels, unsigned max_streams, unsigned ring_slots,
                              unsigned block_frames, unsigned flags)
{
    cmgpu_ctx_t *c;
    (void)device;
    (void)flags;
    if (!channels || channels > 16 || !max_streams || !ring_slots || !block_frames)
        return NULL;
    c = calloc(1, sizeof(*c));
    c->channels = channels;
    c->max_streams = max_streams;
    c->slots = ring_slots;
    c->block_frames = block_frames;
    c->stride = ((size_t)block_frames * channels * 2u + 15u) & ~(size_t)15u;
    c->slot_bytes = c->stride * max_streams;
    c->host = calloc(ring_slots, c->slot_bytes);
    c->dev = calloc(ring_slots, c->slot_bytes);
    c->frames = calloc((size_t)ring_slots * max_streams, sizeof(uint32_t));
    c->has_frames = calloc(ring_slots, 1);
    c->down_pending = calloc(ring_slots, 1);
    c->scale = calloc(max_streams, sizeof(uint16_t));
    c->gain = calloc((size_t)max_streams * channels, sizeof(uint16_t));
    c->meters = calloc(max_streams, sizeof(*c->meters));
    pthread_mutex_init(&c->mu, NULL);
    return c;
}

void cmgpu_ctx_destroy(cmgpu_ctx_t *c)
{
    if (!c)
        return;
    free(c->host); free(c->dev); free(c->frames); free(c->has_frames); free(c->down_pending);
    free(c->scale); free(c->gain); free(c->meters);
    pthread_mutex_destroy(&c->mu);
    free(c);
}

size_t cmgpu_block_stride(const cmgpu_ctx_t *c) { return c ? c->stride : 0; }
uint64_t cmgpu_launch_count(const cmgpu_ctx_t *c) { return c ? c->launches : 0; }
void *cmgpu_host_slot(cmgpu_ctx_t *c, unsigned slot) { return c && slot < c->slots ? c->host + slot * c->slot_bytes : NULL; }

int cmgpu_stream_set_gain(cmgpu_ctx_t *c, unsigned stream, unsigned n, uint16_t scale, const uint16_t *gain)
{
    int rc;
    if (!c)
        return CMGPU_ERR_FAULT;
    if (stream >= c->max_streams)
        return CMGPU_ERR_INVAL;
    pthread_mutex_lock(&c->mu);
    {
        uint16_t g16[16];
        memcpy(g16, c->gain + (size_t)stream * c->channels, sizeof(uint16_t) * c->channels);
        rc = oracle_gain_adapt(c->channels, n, scale, gain, &c->scale[stream], g16);
        memcpy(c->gain + (size_t)stream * c->channels, g16, sizeof(uint16_t) * c->channels);
    }
    pthread_mutex_unlock(&c->mu);
    return rc;
}

int cmgpu_slot_set_frames(cmgpu_ctx_t *c, unsigned slot, const uint32_t *frames)
{
    if (!c || slot >= c->slots)
        return c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT;
    pthread_mutex_lock(&c->mu);
    c->has_frames[slot] = frames != NULL;
    if (frames)
        memcpy(c->frames + (size_t)slot * c->max_streams, frames, sizeof(uint32_t) * c->max_streams);
    pthread_mutex_unlock(&c->mu);
    return CMGPU_OK;
}

static void land(cmgpu_ctx_t *c, unsigned slot)
{
    if (c->down_pending[slot]) {
        memcpy(c->host + slot * c->slot_bytes, c->dev + slot * c->slot_bytes, c->slot_bytes);
        c->down_pending[slot] = 0;
    }
}

int cmgpu_submit(cmgpu_ctx_t *c, unsigned slot, const void *host)
{
    if (!c || slot >= c->slots)
        return c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT;
    pthread_mutex_lock(&c->mu);
    land(c, slot);                  /* the real upload waits for the slot's download, too */
    memcpy(c->dev + slot * c->slot_bytes, host ? host : c->host + slot * c->slot_bytes, c->slot_bytes);
    pthread_mutex_unlock(&c->mu);
    return CMGPU_OK;
}

/* test hook: the next `n` ticks fail the way a CUDA error would (called from the ticking thread only) */
static int g_fail_process;
void cmgpu_stub_fail_next_process(int n) { g_fail_process = n; }

int cmgpu_process(cmgpu_ctx_t *c, unsigned slot, unsigned flags)
{
    unsigned s;
    if (!c || slot >= c->slots)
        return c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT;
    if (g_fail_process > 0) {
        g_fail_process--;
        return CMGPU_ERR_GENERIC;
    }
    pthread_mutex_lock(&c->mu);
    for (s = 0; s < c->max_streams; s++) {
        int16_t *p = (int16_t *)(c->dev + slot * c->slot_bytes + (size_t)s * c->stride);
        const uint32_t n = c->has_frames[slot] ? c->frames[(size_t)slot * c->max_streams + s] : c->block_frames;
        if (flags & CMGPU_TRANSFORM)
            oracle_gain_process(p, n, c->channels, c->scale[s], c->gain + (size_t)s * c->channels);
        if (flags & CMGPU_METER)
            oracle_meter_accumulate(&c->meters[s], p, n, c->channels);
    }
    c->launches++;
    pthread_mutex_unlock(&c->mu);
    return CMGPU_OK;
}

int cmgpu_fetch(cmgpu_ctx_t *c, unsigned slot, void *host)
{
    if (!c || slot >= c->slots)
        return c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT;
    pthread_mutex_lock(&c->mu);
    if (host) {
        memcpy(host, c->dev + slot * c->slot_bytes, c->slot_bytes);
    } else {
        /* not landed yet: whoever reads the pinned slot before waiting sees garbage */
        memset(c->host + slot * c->slot_bytes, 0x5a, c->slot_bytes);
        c->down_pending[slot] = 1;
    }
    pthread_mutex_unlock(&c->mu);
    return CMGPU_OK;
}

int cmgpu_slot_wait(cmgpu_ctx_t *c, unsigned slot)
{
    if (!c || slot >= c->slots)
        return c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT;
    pthread_mutex_lock(&c->mu);
    land(c, slot);
    pthread_mutex_unlock(&c->mu);
    return CMGPU_OK;
}

int cmgpu_sync(cmgpu_ctx_t *c)
{
    unsigned i;
    if (!c)
        return CMGPU_ERR_FAULT;
    pthread_mutex_lock(&c->mu);
    for (i = 0; i < c->slots; i++)
        land(c, i);
    pthread_mutex_unlock(&c->mu);
    return CMGPU_OK;
}

int cmgpu_meter_reset(cmgpu_ctx_t *c, unsigned first, unsigned count)
{
    if (!c)
        return CMGPU_ERR_FAULT;
    if ((uint64_t)first + count > c->max_streams)
        return CMGPU_ERR_INVAL;
    pthread_mutex_lock(&c->mu);
    memset(c->meters + first, 0, sizeof(*c->meters) * count);
    pthread_mutex_unlock(&c->mu);
    return CMGPU_OK;
}

static int one_result(cmgpu_ctx_t *c, unsigned s, uint32_t rate, int reset, cmgpu_result_t *out)
{
    oracle_result_t r;
    oracle_meter_t m = c->meters[s];
    unsigned ch;
    int rc = oracle_meter_finalise(&m, rate, c->channels, &r);
    if (rc != 0)
        return CMGPU_ERR_INVAL;
    if (reset)
        c->meters[s] = m;
    memset(out, 0, sizeof(*out));
    out->rate = r.rate;
    out->channels = r.channels;
    out->frames = r.frames;
    out->global_peak = (int16_t)r.global_peak;
    out->global_power = r.global_power;
    for (ch = 0; ch < c->channels; ch++) {
        out->channel_peak[ch] = (int16_t)r.channel_peak[ch];
        out->channel_power[ch] = r.channel_power[ch];
    }
    return CMGPU_OK;
}

int cmgpu_meter_results(cmgpu_ctx_t *c, unsigned first, unsigned count, uint32_t rate, int reset, unsigned flags,
                        cmgpu_result_t *results, cmgpu_meter_state_t *states, int *rcs)
{
    unsigned i;
    (void)flags;
    (void)states;
    if (!c)
        return CMGPU_ERR_FAULT;
    if ((uint64_t)first + count > c->max_streams)
        return CMGPU_ERR_INVAL;
    pthread_mutex_lock(&c->mu);
    for (i = 0; i < count; i++) {
        cmgpu_result_t tmp;
        int rc = one_result(c, first + i, rate, reset, &tmp);
        if (rc != CMGPU_OK)
            memset(&tmp, 0, sizeof(tmp));
        if (results)
            results[i] = tmp;
        if (rcs)
            rcs[i] = rc;
    }
    pthread_mutex_unlock(&c->mu);
    return CMGPU_OK;
}

int cmgpu_meter_result(cmgpu_ctx_t *c, unsigned stream, uint32_t rate, cmgpu_result_t *out)
{
    int rc;
    if (!c || !out)
        return CMGPU_ERR_FAULT;
    pthread_mutex_lock(&c->mu);
    rc = one_result(c, stream, rate, 1, out);
    pthread_mutex_unlock(&c->mu);
    return rc;
}
