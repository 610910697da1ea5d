/* include/coolmic_b200_shim.h -- the reference's public transform / vumeter / iohandle entry
 * points, exported by libcoolmic_b200.so on top of the cmgpu_* batch engine (include/cmgpu.h).
 *
 * Same names, argument meaning, ownership and error behaviour as libcoolmic-dsp, so that the
 * objects drop into the existing pull chain  snddev -> transform -> tee -> {enc, vumeter}:
 *
 *   coolmic_iohandle_new/read/eof        reference include/coolmic-dsp/iohandle.h:54-66, src/iohandle.c:54-113
 *   coolmic_transform_new/attach_iohandle/get_iohandle/set_master_gain
 *                                        reference include/coolmic-dsp/transform.h:41-53, src/transform.c:65-222
 *   coolmic_vumeter_new/reset/attach_iohandle/read/result
 *                                        reference include/coolmic-dsp/vumeter.h:86-107, src/vumeter.c:69-218
 *
 * Objects are libigloo-style reference-counted objects: the first member is a base carrying the
 * reference count; igloo_ro_ref()/igloo_ro_unref() semantics are provided by coolmic_b200_ref()
 * / coolmic_b200_unref() when the library is built stand-alone (libigloo is an external,
 * un-vendored dependency of the reference). When built into libcoolmic-dsp proper, compile the
 * csrc/host sources with -DCOOLMIC_B200_WITH_IGLOO: they then include the reference's own headers,
 * are libigloo objects, and use the reference's iohandle.c (see below and INTEGRATION.md).
 *
 * The arithmetic of every read runs on the GPU (one tick of a private one-stream cmgpu context
 * per object); there is no CPU implementation behind these calls. Thousands of streams should
 * use the batch engine in cmgpu.h directly -- see INTEGRATION.md.
 */
#ifndef COOLMIC_B200_SHIM_H
#define COOLMIC_B200_SHIM_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef COOLMIC_B200_WITH_IGLOO
/* Integration build (the sources under csrc/host compiled INTO libcoolmic-dsp in place of
 * src/transform.c and src/vumeter.c): no prototype of the reference's API is restated here. The
 * reference's own headers declare it, so the compiler checks every definition in csrc/host against
 * include/coolmic-dsp/{iohandle,transform,vumeter}.h, and objects are libigloo objects
 * (igloo_ro_new_raw / igloo_ro_ref / igloo_ro_unref, like src/transform.c:60-81). iohandle.c stays
 * the reference's. oracle/Makefile builds exactly this configuration into
 * oracle/_ref/libcoolmic_dropin.so for tests/test_gpu_dropin.py. */
#include "types_private.h"                /* reference src/types_private.h: igloo types + ro.h */
#include <coolmic-dsp/coolmic-dsp.h>
#include <coolmic-dsp/iohandle.h>
#include <coolmic-dsp/transform.h>
#include <coolmic-dsp/vumeter.h>
#include <coolmic-dsp/util.h>
typedef igloo_ro_t coolmic_b200_ro_t;
#define COOLMIC_B200_MAX_CHANNELS COOLMIC_DSP_TRANSFORM_MAX_CHANNELS
#else
#ifndef COOLMIC_ERROR_NONE
#define COOLMIC_ERROR_NONE      (0)
#define COOLMIC_ERROR_GENERIC   (-1)
#define COOLMIC_ERROR_NOSYS     (-8)
#define COOLMIC_ERROR_FAULT     (-9)
#define COOLMIC_ERROR_INVAL     (-10)
#define COOLMIC_ERROR_NOMEM     (-11)
#define COOLMIC_ERROR_BUSY      (-12)
#endif

#define COOLMIC_B200_MAX_CHANNELS 16

typedef void *coolmic_b200_ro_t;                 /* what igloo_ro_t is to callers: any object */
int coolmic_b200_ref(coolmic_b200_ro_t object);  /* igloo_ro_ref:   0, or non-zero for NULL */
int coolmic_b200_unref(coolmic_b200_ro_t object);/* igloo_ro_unref: frees at zero            */

typedef struct coolmic_iohandle  coolmic_iohandle_t;
typedef struct coolmic_transform coolmic_transform_t;
typedef struct coolmic_vumeter   coolmic_vumeter_t;

/* Field-for-field the reference's result type (vumeter.h:48-83); 192 bytes on LP64. */
typedef struct {
    uint_least32_t rate;
    unsigned int   channels;
    size_t         frames;
    int16_t        global_peak;
    double         global_power;
    int16_t        channel_peak[COOLMIC_B200_MAX_CHANNELS];
    double         channel_power[COOLMIC_B200_MAX_CHANNELS];
} coolmic_vumeter_result_t;

coolmic_iohandle_t *coolmic_iohandle_new(const char *name, coolmic_b200_ro_t associated, void *userdata,
                                         int (*free)(void *), ssize_t (*read)(void *, void *, size_t),
                                         int (*eof)(void *));
ssize_t             coolmic_iohandle_read(coolmic_iohandle_t *self, void *buffer, size_t len);
int                 coolmic_iohandle_eof(coolmic_iohandle_t *self);

coolmic_transform_t *coolmic_transform_new(const char *name, coolmic_b200_ro_t associated,
                                           uint_least32_t rate, unsigned int channels);
int                  coolmic_transform_attach_iohandle(coolmic_transform_t *self, coolmic_iohandle_t *handle);
coolmic_iohandle_t  *coolmic_transform_get_iohandle(coolmic_transform_t *self);
int                  coolmic_transform_set_master_gain(coolmic_transform_t *self, unsigned int channels,
                                                       uint16_t scale, const uint16_t *gain);

coolmic_vumeter_t *coolmic_vumeter_new(const char *name, coolmic_b200_ro_t associated,
                                       uint_least32_t rate, unsigned int channels);
int                coolmic_vumeter_reset(coolmic_vumeter_t *self);
int                coolmic_vumeter_attach_iohandle(coolmic_vumeter_t *self, coolmic_iohandle_t *handle);
ssize_t            coolmic_vumeter_read(coolmic_vumeter_t *self, ssize_t maxlen);
int                coolmic_vumeter_result(coolmic_vumeter_t *self, coolmic_vumeter_result_t *result);

/* Meter results -> colours (reference include/coolmic-dsp/util.h:40-47, src/util.c:59-139). Host-side
 * double arithmetic with the reference's expressions and libm: identical doubles and ARGB words. */
#define COOLMIC_UTIL_PROFILE_DEFAULT "default"
typedef uint32_t coolmic_argb_t;
coolmic_argb_t coolmic_util_ahsv2argb(double alpha, double hue, double saturation, double value);
double         coolmic_util_power2hue(double power, const char *profile);
double         coolmic_util_peak2hue(int16_t peak, const char *profile);
#endif /* COOLMIC_B200_WITH_IGLOO */

/* ---- batch mode: the same objects, many streams per GPU tick ---------------
 * A batch owns one cmgpu context for `channels`-channel streams. Its member transforms are
 * ordinary coolmic_transform_t objects (attach_iohandle / get_iohandle / set_master_gain work as
 * usual); coolmic_b200_batch_tick() pulls up to block_frames from every member's input, runs ONE
 * fused transform+vumeter launch for all members and makes the transformed PCM readable through
 * the members' handles -- every handle has its own read position, so several consumers can read
 * one transform without a tee. A batch vumeter is fused with its transform: no handle, no copy.
 * tick() returns the frames processed (>= 0), COOLMIC_ERROR_BUSY (-12) while some reader has not
 * consumed the previous tick's output, or another negative COOLMIC_ERROR_* code. */
typedef struct coolmic_b200_batch coolmic_b200_batch_t;
coolmic_b200_batch_t *coolmic_b200_batch_new(int device, unsigned int channels, unsigned int max_streams,
                                             unsigned int block_frames);
coolmic_transform_t  *coolmic_b200_batch_transform_new(coolmic_b200_batch_t *batch, const char *name,
                                                       coolmic_b200_ro_t associated, uint_least32_t rate);
coolmic_vumeter_t    *coolmic_b200_batch_vumeter_new(coolmic_b200_batch_t *batch, coolmic_transform_t *of,
                                                     const char *name, coolmic_b200_ro_t associated);
int                   coolmic_b200_batch_tick(coolmic_b200_batch_t *batch);
size_t                coolmic_b200_batch_pending(coolmic_b200_batch_t *batch);
/* The same batch on a ring of `ring_slots` (1..64) pinned/device slots: tick() only QUEUES the slot's
 * upload, fused launch and download and returns, so tick t uploads while t-1 computes and t-2
 * downloads (simple.c:445-505's loop, `ring_slots` blocks deep). Readers follow tick by tick; the
 * first read of a tick's output waits for that slot's download alone. tick() is refused with
 * COOLMIC_ERROR_BUSY while a reader still has unread output in the slot it would reuse, i.e. once
 * the slowest reader is `ring_slots` ticks behind. `pull_threads` (>= 1) host threads share the
 * members' input pulls of a tick (each input handle is called from one thread at a time).
 * coolmic_b200_batch_new() is the ring_slots = 1, pull_threads = 1 case. Readers of DIFFERENT handles
 * may run on different threads, also while the driver is inside tick() (which never touches a slot
 * with unread output); one handle, like any coolmic object, belongs to one thread at a time. */
coolmic_b200_batch_t *coolmic_b200_batch_new_ring(int device, unsigned int channels, unsigned int max_streams,
                                                  unsigned int block_frames, unsigned int ring_slots,
                                                  unsigned int pull_threads);
/* coolmic_vumeter_result() for every member at once, with ONE device round trip: results[s], rcs[s]
 * for s = 0 .. max_streams-1 (rcs[s] = COOLMIC_ERROR_INVAL where nothing was metered; vumeter.c:198). */
int                   coolmic_b200_batch_results(coolmic_b200_batch_t *batch, coolmic_vumeter_result_t *results,
                                                 int *rcs);

/* Which CUDA device the objects created from now on use (default 0, or $COOLMIC_B200_DEVICE). */
int coolmic_b200_set_device(int device);
/* Kernel launches issued on behalf of shim objects so far (evidence that reads run on the GPU). */
uint64_t coolmic_b200_shim_launches(void);
/* Measurement: the loop a host application runs (simple.c:445-505), through the objects above only.
 * `streams` member transforms of a ring batch, each fed by a memory iohandle that cycles over its row
 * of `pcm` ([streams][bytes_per_stream]), gains as in bench.py, fused vumeters; `n_ticks` ticks of
 * `block_frames`; every transform's output is read back through its own iohandle by `threads`
 * consumer threads, `ring_slots - 1` ticks behind the producer; one coolmic_b200_batch_results() at
 * the end. *seconds = wall time of the loop, *frames_metered = sum of result.frames. */
int coolmic_b200_bench_objects(int device, unsigned int channels, unsigned int streams, unsigned int block_frames,
                               unsigned int n_ticks, unsigned int ring_slots, unsigned int threads,
                               unsigned int bytes_per_stream, const void *pcm, double *seconds,
                               uint64_t *frames_metered);

#ifdef __cplusplus
}
#endif
#endif
