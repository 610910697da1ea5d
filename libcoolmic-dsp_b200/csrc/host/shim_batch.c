/* host/shim_batch.c -- many coolmic_transform_t / coolmic_vumeter_t objects on ONE batch engine.
 *
 * The reference runs one pull chain per stream, 1 KiB at a time, on one thread each
 * (src/simple.c:445-505). Here the objects keep their API but share a schedule: a driver calls
 * coolmic_b200_batch_tick(), which pulls up to block_frames from every member transform's input
 * handle straight into a pinned ring slot (whole frames only, the unfinished frame is carried
 * like transform.c:141-160), queues the slot's upload, ONE fused transform+vumeter tick for all
 * members and the download of the transformed PCM -- and returns without waiting for any of it.
 * With a ring of n slots (coolmic_b200_batch_new_ring) tick t uploads while tick t-1 computes and
 * tick t-2 downloads; a 1-slot batch (coolmic_b200_batch_new) degenerates to tick, read, tick, read.
 *   - every handle obtained from a member transform reads the stream's output with its own cursor,
 *     tick after tick in order (the fan-out tee.c:167-206 does with per-reader offsets, without the
 *     tee's copy); the first read of a tick's output waits for THAT slot's download only
 *     (cmgpu_slot_wait), and a reader that has caught up gets 0 ("nothing now", iohandle.h:44-47);
 *   - tick() answers COOLMIC_ERROR_BUSY while some reader still has unread output in the slot the
 *     tick would reuse -- the back-pressure tee.c:145-151 applies to a lagging reader;
 *   - a fused vumeter needs no handle at all: the device metered the block in the same pass;
 *     coolmic_vumeter_read() reports the bytes metered since its last call and
 *     coolmic_vumeter_result() finalises the stream's device state (vumeter.c:189-218).
 * The meter window is therefore quantised to ticks: read(-1) ("an unspecified internal default",
 * vumeter.h:97) means "everything the last ticks metered".
 *
 * Pulling 1,024+ inputs is host memcpy work; a ring batch can spread it over a small pool of
 * threads (each member's input handle is only ever called from one thread at a time).
 */
#include "shim_internal.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct batch_cursor {
    unsigned stream;
    uint64_t epoch;            /* tick whose output the cursor is reading (ticks are numbered from 0) */
    size_t offset;             /* bytes of that output already handed out */
    struct batch_cursor *next;
} batch_cursor_t;

struct coolmic_b200_batch {
    shim_base_t base;
    cmgpu_ctx_t *ctx;
    unsigned channels, max_streams, block_frames, slots;
    size_t stride, framesize;
    uint64_t epoch;                    /* ticks issued so far; tick k lives in slot k % slots */
    coolmic_transform_t **member;      /* [max_streams], not owned (members own the batch) */
    size_t *out_bytes;                 /* [slots][max_streams] valid transformed bytes of the slot's tick */
    uint64_t *metered;                 /* [max_streams] bytes metered since the stream joined */
    uint32_t *frames;                  /* [max_streams] scratch for cmgpu_slot_set_frames */
    unsigned char *landed;             /* [slots] the slot's download is known to be complete */
    unsigned char *failed;             /* [slots] ... or to have failed */
    batch_cursor_t *cursors;
    pthread_mutex_t cursor_mu;         /* the cursor list and landed[] (readers may run on several threads) */
    /* pull pool */
    unsigned n_workers;
    pthread_t *workers;
    pthread_mutex_t mu;
    pthread_cond_t cv_work, cv_done;
    uint64_t job;                      /* generation of the current pull job (its own counter: a tick that
                                          fails after its pull does not advance `epoch`, the next one is still a new job) */
    unsigned job_left;
    unsigned char *job_slot;           /* pinned slot being filled ... */
    size_t *job_ob;                    /* ... and its out_bytes row */
    long job_total;                    /* frames pulled by the workers */
    int quit;
};

/* one member's pull: carry first, then the input handle, whole frames only (transform.c:141-160) */
static uint32_t pull_member(coolmic_b200_batch_t *b, unsigned s, unsigned char *slot, size_t *out_have)
{
    coolmic_transform_t *t = b->member[s];
    unsigned char *dst = slot + (size_t)s * b->stride;
    const size_t want = (size_t)b->block_frames * b->framesize;
    size_t have = 0, rest;
    *out_have = 0;
    if (!t)
        return 0;
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
    *out_have = have;
    return (uint32_t)(have / b->framesize);
}

static long pull_range(coolmic_b200_batch_t *b, unsigned char *slot, size_t *out_bytes, unsigned lo, unsigned hi)
{
    long total = 0;
    unsigned s;
    for (s = lo; s < hi; s++) {
        b->frames[s] = pull_member(b, s, slot, &out_bytes[s]);
        total += (long)b->frames[s];
    }
    return total;
}

typedef struct worker_arg {
    coolmic_b200_batch_t *b;
    unsigned index;
} worker_arg_t;

static void *pull_worker(void *p)
{
    worker_arg_t *a = p;
    coolmic_b200_batch_t *b = a->b;
    const unsigned parts = b->n_workers + 1;       /* the caller of tick() takes the last share itself */
    const unsigned lo = (unsigned)((uint64_t)b->max_streams * a->index / parts);
    const unsigned hi = (unsigned)((uint64_t)b->max_streams * (a->index + 1) / parts);
    uint64_t seen = 0;
    free(a);
    pthread_mutex_lock(&b->mu);
    for (;;) {
        while (!b->quit && b->job == seen)
            pthread_cond_wait(&b->cv_work, &b->mu);
        if (b->quit)
            break;
        seen = b->job;
        {
            unsigned char *slot = b->job_slot;
            size_t *ob = b->job_ob;
            long got;
            pthread_mutex_unlock(&b->mu);
            got = pull_range(b, slot, ob, lo, hi);
            pthread_mutex_lock(&b->mu);
            b->job_total += got;
        }
        if (--b->job_left == 0)
            pthread_cond_signal(&b->cv_done);
    }
    pthread_mutex_unlock(&b->mu);
    return NULL;
}

static void batch_destroy(shim_self_t self)
{
    coolmic_b200_batch_t *b = SHIM_SELF(self, coolmic_b200_batch_t);
    unsigned i;
    if (b->workers) {
        pthread_mutex_lock(&b->mu);
        b->quit = 1;
        pthread_cond_broadcast(&b->cv_work);
        pthread_mutex_unlock(&b->mu);
        for (i = 0; i < b->n_workers; i++)
            pthread_join(b->workers[i], NULL);
        free(b->workers);
    }
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
    free(b->landed);
    free(b->failed);
    pthread_mutex_destroy(&b->mu);
    pthread_mutex_destroy(&b->cursor_mu);
    pthread_cond_destroy(&b->cv_work);
    pthread_cond_destroy(&b->cv_done);
}

SHIM_TYPE(coolmic_b200_batch_t, batch_destroy);

coolmic_b200_batch_t *coolmic_b200_batch_new_ring(int device, unsigned int channels, unsigned int max_streams,
                                                  unsigned int block_frames, unsigned int ring_slots,
                                                  unsigned int pull_threads)
{
    coolmic_b200_batch_t *b;
    unsigned i;
    if (!channels || channels > COOLMIC_B200_MAX_CHANNELS || !max_streams || !block_frames || !ring_slots ||
        ring_slots > 64)
        return NULL;
    b = SHIM_NEW(coolmic_b200_batch_t, batch_destroy, "batch", SHIM_RO_NULL);
    if (!b)
        return NULL;
    pthread_mutex_init(&b->mu, NULL);
    pthread_mutex_init(&b->cursor_mu, NULL);
    pthread_cond_init(&b->cv_work, NULL);
    pthread_cond_init(&b->cv_done, NULL);
    b->channels = channels;
    b->max_streams = max_streams;
    b->block_frames = block_frames;
    b->slots = ring_slots;
    b->framesize = 2u * channels;
    b->member = calloc(max_streams, sizeof(*b->member));
    b->out_bytes = calloc((size_t)max_streams * ring_slots, sizeof(*b->out_bytes));
    b->metered = calloc(max_streams, sizeof(*b->metered));
    b->frames = calloc(max_streams, sizeof(*b->frames));
    b->landed = calloc(ring_slots, 1);
    b->failed = calloc(ring_slots, 1);
    b->ctx = cmgpu_ctx_create(device < 0 ? shim_device() : device, channels, max_streams, ring_slots, block_frames, 0);
    if (!b->member || !b->out_bytes || !b->metered || !b->frames || !b->landed || !b->failed || !b->ctx) {
        shim_unref(b);
        return NULL;
    }
    b->stride = cmgpu_block_stride(b->ctx);
    if (pull_threads > max_streams)
        pull_threads = max_streams;
    if (pull_threads > 1) {
        b->workers = calloc(pull_threads - 1, sizeof(*b->workers));
        if (!b->workers) {
            shim_unref(b);
            return NULL;
        }
        b->n_workers = pull_threads - 1;
        for (i = 0; i < pull_threads - 1; i++) {
            worker_arg_t *a = malloc(sizeof(*a));
            if (!a || (a->b = b, a->index = i, pthread_create(&b->workers[i], NULL, pull_worker, a)) != 0) {
                free(a);
                b->n_workers = i;           /* the ones started so far are joined by the destructor */
                shim_unref(b);
                return NULL;
            }
        }
    }
    return b;
}

coolmic_b200_batch_t *coolmic_b200_batch_new(int device, unsigned int channels, unsigned int max_streams,
                                             unsigned int block_frames)
{
    return coolmic_b200_batch_new_ring(device, channels, max_streams, block_frames, 1, 1);
}

coolmic_transform_t *coolmic_b200_batch_transform_new(coolmic_b200_batch_t *b, const char *name,
                                                      coolmic_b200_ro_t associated, uint_least32_t rate)
{
    coolmic_transform_t *t;
    unsigned s, k;
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
    for (k = 0; k < b->slots; k++)
        b->out_bytes[(size_t)k * b->max_streams + s] = 0;
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

/* a cursor that is through with a tick's output moves on to the next tick (cursor_mu held) */
static void cursor_settle(const coolmic_b200_batch_t *b, batch_cursor_t *c)
{
    while (c->epoch < b->epoch && c->offset >= b->out_bytes[(size_t)(c->epoch % b->slots) * b->max_streams + c->stream]) {
        c->epoch++;
        c->offset = 0;
    }
}

/* unread bytes of `c` over the ticks [c->epoch, b->epoch), all of which are still in the ring */
static size_t cursor_unread(const coolmic_b200_batch_t *b, const batch_cursor_t *c)
{
    size_t n = 0;
    uint64_t e;
    for (e = c->epoch; e < b->epoch; e++) {
        const size_t have = b->out_bytes[(size_t)(e % b->slots) * b->max_streams + c->stream];
        if (e != c->epoch)
            n += have;
        else if (have > c->offset)
            n += have - c->offset;
    }
    return n;
}

size_t coolmic_b200_batch_pending(coolmic_b200_batch_t *b)
{
    size_t worst = 0;
    batch_cursor_t *c;
    if (!b)
        return 0;
    pthread_mutex_lock(&b->cursor_mu);
    for (c = b->cursors; c; c = c->next) {
        const size_t n = cursor_unread(b, c);
        if (n > worst)
            worst = n;
    }
    pthread_mutex_unlock(&b->cursor_mu);
    return worst;
}

int coolmic_b200_batch_tick(coolmic_b200_batch_t *b)
{
    unsigned char *slot_mem;
    size_t *ob;
    uint64_t before;
    long total = 0;
    unsigned s, slot;
    batch_cursor_t *c;
    int busy = 0;

    if (!b)
        return COOLMIC_ERROR_FAULT;
    slot = (unsigned)(b->epoch % b->slots);
    ob = b->out_bytes + (size_t)slot * b->max_streams;
    /* the slot still holds the output of tick epoch - slots: every reader must be through with it */
    if (b->epoch >= b->slots) {
        const uint64_t old = b->epoch - b->slots;
        pthread_mutex_lock(&b->cursor_mu);
        for (c = b->cursors; c && !busy; c = c->next) {
            cursor_settle(b, c);
            busy = c->epoch <= old;
        }
        pthread_mutex_unlock(&b->cursor_mu);
        if (busy)
            return COOLMIC_ERROR_BUSY;                 /* a reader has not caught up */
        /* ... and its download must have left the pinned slot we are about to refill */
        if (!b->landed[slot] && cmgpu_slot_wait(b->ctx, slot) != CMGPU_OK)
            return COOLMIC_ERROR_GENERIC;
    }
    slot_mem = cmgpu_host_slot(b->ctx, slot);
    if (b->n_workers) {
        pthread_mutex_lock(&b->mu);
        b->job_slot = slot_mem;
        b->job_ob = ob;
        b->job_total = 0;
        b->job_left = b->n_workers;
        b->job++;
        pthread_cond_broadcast(&b->cv_work);
        pthread_mutex_unlock(&b->mu);
        total = pull_range(b, slot_mem, ob, (unsigned)((uint64_t)b->max_streams * b->n_workers / (b->n_workers + 1)),
                           b->max_streams);
        pthread_mutex_lock(&b->mu);
        while (b->job_left)
            pthread_cond_wait(&b->cv_done, &b->mu);
        total += b->job_total;
        pthread_mutex_unlock(&b->mu);
    } else {
        total = pull_range(b, slot_mem, ob, 0, b->max_streams);
    }
    before = cmgpu_launch_count(b->ctx);
    pthread_mutex_lock(&b->cursor_mu);
    b->landed[slot] = 0;
    b->failed[slot] = 0;
    pthread_mutex_unlock(&b->cursor_mu);
    if (cmgpu_slot_set_frames(b->ctx, slot, b->frames) != CMGPU_OK || cmgpu_submit(b->ctx, slot, NULL) != CMGPU_OK ||
        cmgpu_process(b->ctx, slot, CMGPU_FUSED) != CMGPU_OK || cmgpu_fetch(b->ctx, slot, NULL) != CMGPU_OK) {
        /* nothing of this tick is exposed: readers never see PCM that was not processed */
        for (s = 0; s < b->max_streams; s++)
            ob[s] = 0;
        b->failed[slot] = 1;
        return COOLMIC_ERROR_GENERIC;
    }
    shim_count_launches(cmgpu_launch_count(b->ctx) - before);
    /* committed only now: the tick is queued; its results become readable as the slot lands */
    for (s = 0; s < b->max_streams; s++)
        b->metered[s] += ob[s];
    pthread_mutex_lock(&b->cursor_mu);
    b->epoch++;
    pthread_mutex_unlock(&b->cursor_mu);
    return total > 0x7fffffffL ? 0x7fffffff : (int)total;
}

/* ---- hooks for the object code ------------------------------------------------------------ */
void *shim_batch_cursor_new(coolmic_b200_batch_t *b, unsigned stream)
{
    batch_cursor_t *c = calloc(1, sizeof(*c));
    if (!c)
        return NULL;
    c->stream = stream;
    c->offset = 0;
    pthread_mutex_lock(&b->cursor_mu);
    c->epoch = b->epoch;                               /* a new reader starts at the next tick's output */
    c->next = b->cursors;
    b->cursors = c;
    pthread_mutex_unlock(&b->cursor_mu);
    return c;
}

void shim_batch_cursor_free(coolmic_b200_batch_t *b, void *cursor)
{
    batch_cursor_t **pp;
    pthread_mutex_lock(&b->cursor_mu);
    for (pp = &b->cursors; *pp; pp = &(*pp)->next) {
        if (*pp == cursor) {
            *pp = (*pp)->next;
            free(cursor);
            break;
        }
    }
    pthread_mutex_unlock(&b->cursor_mu);
}

size_t shim_batch_cursor_unread(coolmic_b200_batch_t *b, void *cursor)
{
    size_t n;
    pthread_mutex_lock(&b->cursor_mu);
    n = cursor_unread(b, cursor);
    pthread_mutex_unlock(&b->cursor_mu);
    return n;
}

ssize_t shim_batch_read(coolmic_b200_batch_t *b, unsigned stream, void *cursor, void *buffer, size_t len)
{
    batch_cursor_t *c = cursor;
    unsigned slot;
    size_t have, off, n;
    int landed, failed;
    /* every change of cursor state happens under cursor_mu (the driver's tick() settles cursors too);
     * only the wait for the slot and the copy itself run unlocked */
    pthread_mutex_lock(&b->cursor_mu);
    cursor_settle(b, c);                               /* through with a tick: on to the next one */
    if (c->epoch >= b->epoch) {
        pthread_mutex_unlock(&b->cursor_mu);
        return 0;                                      /* caught up: nothing now (iohandle.h:44-47) */
    }
    slot = (unsigned)(c->epoch % b->slots);
    have = b->out_bytes[(size_t)slot * b->max_streams + stream];
    off = c->offset;
    landed = b->landed[slot];
    failed = b->failed[slot];
    pthread_mutex_unlock(&b->cursor_mu);
    if (!landed) {
        /* the first reader of a tick's output waits for that slot's download, nothing else */
        const int rc = cmgpu_slot_wait(b->ctx, slot);
        pthread_mutex_lock(&b->cursor_mu);
        if (rc != CMGPU_OK)
            b->failed[slot] = 1;
        b->landed[slot] = 1;
        failed = b->failed[slot];
        pthread_mutex_unlock(&b->cursor_mu);
    }
    if (failed)
        return -1;                                     /* a CUDA failure is a read error; no CPU fallback */
    n = have - off;
    if (n > len)
        n = len;
    memcpy(buffer, (unsigned char *)cmgpu_host_slot(b->ctx, slot) + (size_t)stream * b->stride + off, n);
    pthread_mutex_lock(&b->cursor_mu);
    c->offset = off + n;
    pthread_mutex_unlock(&b->cursor_mu);
    return (ssize_t)n;
}

int shim_batch_set_gain(coolmic_b200_batch_t *b, unsigned stream, uint16_t scale, const uint16_t *gain)
{
    return cmgpu_stream_set_gain(b->ctx, stream, scale ? b->channels : 0, scale, gain) == CMGPU_OK ? 0 : -1;
}

void shim_batch_release(coolmic_b200_batch_t *b, unsigned stream)
{
    unsigned k;
    b->member[stream] = NULL;
    for (k = 0; k < b->slots; k++)
        b->out_bytes[(size_t)k * b->max_streams + stream] = 0;
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

/* All members' results with ONE device round trip (cmgpu_meter_results): results[s] / rcs[s] for the
 * member streams s = 0 .. max_streams-1; rcs[s] = COOLMIC_ERROR_INVAL where nothing was metered. */
int coolmic_b200_batch_results(coolmic_b200_batch_t *b, coolmic_vumeter_result_t *results, int *rcs)
{
    cmgpu_result_t *tmp;
    unsigned s, c;
    int rc;
    if (!b || !results || !rcs)
        return COOLMIC_ERROR_FAULT;
    tmp = calloc(b->max_streams, sizeof(*tmp));
    if (!tmp)
        return COOLMIC_ERROR_NOMEM;
    rc = cmgpu_meter_results(b->ctx, 0, b->max_streams, 0, 1, 0, tmp, NULL, rcs);
    if (rc != CMGPU_OK) {
        free(tmp);
        return COOLMIC_ERROR_GENERIC;
    }
    for (s = 0; s < b->max_streams; s++) {
        coolmic_vumeter_result_t *out = &results[s];
        coolmic_transform_t *t = b->member[s];
        memset(out, 0, sizeof(*out));
        if (rcs[s] != CMGPU_OK || !t) {
            rcs[s] = COOLMIC_ERROR_INVAL;
            continue;
        }
        out->rate = t->rate;
        out->channels = t->channels;
        out->frames = (size_t)tmp[s].frames;
        out->global_peak = tmp[s].global_peak;
        out->global_power = tmp[s].global_power;
        for (c = 0; c < t->channels; c++) {
            out->channel_peak[c] = tmp[s].channel_peak[c];
            out->channel_power[c] = tmp[s].channel_power[c];
        }
        rcs[s] = COOLMIC_ERROR_NONE;
    }
    free(tmp);
    return COOLMIC_ERROR_NONE;
}
