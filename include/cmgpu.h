/* include/cmgpu.h -- the thin C ABI beneath the coolmic_transform_* / coolmic_vumeter_* API.
 *
 * B200-native batch engine for the one data-parallel hot path of libcoolmic-dsp:
 *   coolmic_transform  (per-channel integer master gain on interleaved S16,
 *                       reference src/transform.c:101-124)
 *   coolmic_vumeter    (per-channel + global first-occurrence signed peak and exact
 *                       int64 sum of squares, reference src/vumeter.c:161-177, and the dB
 *                       finaliser, reference src/vumeter.c:189-218)
 * fused into one pass over a device-resident ring of interleaved S16 stream-blocks.
 *
 * Plain C types only: pointers and sizes, no CUDA or torch types in any signature. Every
 * int-returning entry point returns a COOLMIC_ERROR_*-compatible code (reference
 * include/coolmic-dsp/coolmic-dsp.h:35-49): 0 ok, -1 generic (any CUDA error), -9 NULL
 * argument, -10 invalid argument, -11 out of memory, -12 busy. There is NO CPU fallback:
 * without a usable CUDA device cmgpu_ctx_create() fails and everything else refuses.
 *
 * Vocabulary
 *   stream        one independent PCM stream (= one coolmic_transform_t + coolmic_vumeter_t pair)
 *   stream-block  `block_frames` frames of one stream: block_frames * channels S16 samples,
 *                 stored contiguously; its byte stride in a slot is rounded up to 16
 *   slot          one tick of the ring: [max_streams] stream-blocks, device resident
 *   tick          one cmgpu_process() over a slot: ONE fused kernel launch over all streams
 *
 * Which reference interface each entry point replaces is noted beside it.
 */
#ifndef CMGPU_H
#define CMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMGPU_MAX_CHANNELS 16      /* COOLMIC_DSP_TRANSFORM_MAX_CHANNELS, transform.h:35; vumeter.h:42 */

#define CMGPU_OK            0      /* COOLMIC_ERROR_NONE    */
#define CMGPU_ERR_GENERIC  (-1)    /* COOLMIC_ERROR_GENERIC */
#define CMGPU_ERR_NOSYS    (-8)    /* COOLMIC_ERROR_NOSYS   */
#define CMGPU_ERR_FAULT    (-9)    /* COOLMIC_ERROR_FAULT   */
#define CMGPU_ERR_INVAL   (-10)    /* COOLMIC_ERROR_INVAL   */
#define CMGPU_ERR_NOMEM   (-11)    /* COOLMIC_ERROR_NOMEM   */
#define CMGPU_ERR_BUSY    (-12)    /* COOLMIC_ERROR_BUSY    */

/* cmgpu_ctx_create flags */
#define CMGPU_SEPARATE_OUT   0x1u  /* transformed PCM goes to a second ring instead of in place
                                      (the reference works in place, transform.c:120) */
#define CMGPU_NO_PINNED      0x2u  /* do not allocate the pinned host staging rings */
#define CMGPU_FORCE_GENERIC  0x4u  /* always use the any-channel-count kernel (test hook) */
#define CMGPU_PLANAR_F32     0x8u  /* also keep a ring of de-interleaved float planes (see CMGPU_PLANAR) */
#define CMGPU_MIX_OUTPUT_METER_ONLY 0x10u /* downmix contexts (extension): meter the outputs only, not the inputs */

/* cmgpu_process flags */
#define CMGPU_TRANSFORM      0x1u  /* apply the gain tables (else pass PCM through untouched) */
#define CMGPU_METER          0x2u  /* accumulate the meters over what the slot holds afterwards */
#define CMGPU_FUSED          (CMGPU_TRANSFORM | CMGPU_METER)
#define CMGPU_PLANAR         0x4u  /* second output of the same pass: per channel a plane of
                                      sample / 32768.f floats, what enc_vorbis.c:108-117 builds for
                                      vorbis_analysis_buffer(); needs a CMGPU_PLANAR_F32 context */

typedef struct cmgpu_ctx cmgpu_ctx_t;

/* Integer meter state of one stream between reset and result: the exact content of
 * struct coolmic_vumeter's accumulators (vumeter.c:48-56) in decoded form. */
typedef struct cmgpu_meter_state {
    uint64_t frames;                              /* result.frames                     */
    int64_t  power[CMGPU_MAX_CHANNELS];           /* power[c] = sum of squares         */
    int16_t  channel_peak[CMGPU_MAX_CHANNELS];    /* result.channel_peak[c]            */
    int16_t  global_peak;                         /* result.global_peak                */
    int16_t  reserved[3];
} cmgpu_meter_state_t;

/* Same members as coolmic_vumeter_result_t (vumeter.h:48-83) with fixed-width types. */
typedef struct cmgpu_result {
    uint32_t rate;
    uint32_t channels;
    uint64_t frames;
    int16_t  global_peak;
    double   global_power;
    int16_t  channel_peak[CMGPU_MAX_CHANNELS];
    double   channel_power[CMGPU_MAX_CHANNELS];
} cmgpu_result_t;

/* ---- library ------------------------------------------------------------------ */
const char *cmgpu_version(void);
int         cmgpu_device_count(void);                 /* 0 when no CUDA device is usable   */
const char *cmgpu_last_error(void);                   /* thread-local text of the last failure */

/* Page-locked host memory for callers that stream whole slots from their own buffers
 * (cmgpu_submit / cmgpu_fetch are asynchronous only for page-locked memory). */
void *cmgpu_host_alloc(size_t bytes);
/* The same, write-combined: for buffers the host only ever WRITES (capture -> upload); reads by the
 * host are very slow. */
void *cmgpu_host_alloc_wc(size_t bytes);
void  cmgpu_host_free(void *p);

/* ---- context: one per (GPU, channel count) -------------------------------------- */
/* All streams of a context share `channels` (1..16) and `block_frames` (>= 1). `ring_slots`
 * (>= 1) stream-block sets are resident on the device; with pinned staging the same number
 * exist in page-locked host memory. Returns NULL on error (see cmgpu_last_error()). */
cmgpu_ctx_t *cmgpu_ctx_create(int device, unsigned channels, unsigned max_streams,
                              unsigned ring_slots, unsigned block_frames, unsigned flags);
void         cmgpu_ctx_destroy(cmgpu_ctx_t *ctx);

unsigned cmgpu_channels(const cmgpu_ctx_t *ctx);
unsigned cmgpu_max_streams(const cmgpu_ctx_t *ctx);
unsigned cmgpu_ring_slots(const cmgpu_ctx_t *ctx);
unsigned cmgpu_block_frames(const cmgpu_ctx_t *ctx);
size_t   cmgpu_block_stride(const cmgpu_ctx_t *ctx);  /* bytes between stream-blocks in a slot */
size_t   cmgpu_slot_bytes(const cmgpu_ctx_t *ctx);    /* max_streams * block_stride          */
/* How many streams a tick covers (default max_streams). Lets one context serve fewer. */
int      cmgpu_set_active_streams(cmgpu_ctx_t *ctx, unsigned n);

/* ---- gain tables (replaces coolmic_transform_set_master_gain, transform.c:195-222) -- */
/* Reference adaptation rules for one stream: n==0 || scale==0 || gain==NULL disables;
 * n==channels copies; n==1 broadcasts; n==2 on a mono stream averages; anything else is
 * CMGPU_ERR_INVAL and the previous setting is kept. Takes effect at the next tick. */
int cmgpu_stream_set_gain(cmgpu_ctx_t *ctx, unsigned stream, unsigned n, uint16_t scale,
                          const uint16_t *gain);
/* Bulk form: scale[count], gain[count][channels], already per-channel. */
int cmgpu_set_gain_table(cmgpu_ctx_t *ctx, unsigned first, unsigned count,
                         const uint16_t *scale, const uint16_t *gain);
/* Reads back what the device will use (after adaptation). */
int cmgpu_stream_get_gain(const cmgpu_ctx_t *ctx, unsigned stream, uint16_t *scale,
                          uint16_t gain[CMGPU_MAX_CHANNELS]);

/* ---- ring I/O (replaces the byte shuffling of transform.c:126-165 and tee.c:137-206) --- */
void *cmgpu_host_slot(cmgpu_ctx_t *ctx, unsigned slot);      /* pinned staging, in/out     */
void *cmgpu_device_slot(cmgpu_ctx_t *ctx, unsigned slot);    /* device input ring          */
void *cmgpu_device_out_slot(cmgpu_ctx_t *ctx, unsigned slot);/* == input ring when in place */
/* Valid frames per stream for a slot: frames[active_streams] (each <= block_frames) or
 * NULL for "every stream-block is full". Copied; applies to subsequent ticks of the slot. */
int cmgpu_slot_set_frames(cmgpu_ctx_t *ctx, unsigned slot, const uint32_t *frames);
/* Host -> device copy of a slot on the upload stream. `host` NULL = the pinned staging slot.
 * Asynchronous when the source is page-locked. */
int cmgpu_submit(cmgpu_ctx_t *ctx, unsigned slot, const void *host);
/* One tick: ONE fused kernel launch over all active streams of the slot, on the compute
 * stream, ordered after the slot's last submit. Asynchronous. Ticks on data that is already on
 * the device (no submit / fetch of the slot since its last tick) are bare kernel launches, and
 * consecutive ones overlap at their edges when they cannot conflict (different slots, or a
 * CMGPU_SEPARATE_OUT context): the meters do not depend on completion order. */
int cmgpu_process(cmgpu_ctx_t *ctx, unsigned slot, unsigned flags);
/* Small-buffer regime: `n_slots` consecutive ticks (slots first_slot ..) issued as ONE launch that
 * walks all of them (the ring is one allocation), so that per-tick launch latency does not
 * dominate 20 ms blocks -- with enough streams to fill the GPU, one 8-lane group per STREAM that keeps
 * its meter partials in registers across the ticks and publishes them once; contexts whose kernels
 * cannot do that (channel counts that do not tile 16
 * bytes, downmix, float planes, frame counts set for only some of the slots) get ONE cached CUDA
 * graph of per-tick launches instead, rebuilt when gains, frames or the active stream count
 * change. Either way the ticks count as slots first_slot, first_slot + 1, ... in time order.
 * Ordered after the slots' last submits. Asynchronous. */
int cmgpu_process_cycle(cmgpu_ctx_t *ctx, unsigned first_slot, unsigned n_slots, unsigned flags);
/* Device -> host copy of the slot's (transformed) PCM on the download stream, ordered after
 * the slot's last tick. `host` NULL = the pinned staging slot. Asynchronous if page-locked.
 * The download is skipped when it could only copy the input onto itself: in-place context, `host`
 * NULL after a submit with `host` NULL, and no tick has written the slot's PCM since -- pass-through
 * streams, the reference's default state (transform.c:107-108), are metered without a byte coming
 * back. cmgpu_transfer_bytes reports what submit / fetch have really copied. */
int cmgpu_fetch(cmgpu_ctx_t *ctx, unsigned slot, void *host);
int cmgpu_transfer_bytes(const cmgpu_ctx_t *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);
/* Float planes of a slot (CMGPU_PLANAR): [stream][channel][cmgpu_plane_stride()] float, valid for
 * the stream's frames of the tick. Device pointer, and a download like cmgpu_fetch (`host` must be
 * given: max_streams * channels * plane_stride floats). */
void  *cmgpu_device_planar_slot(cmgpu_ctx_t *ctx, unsigned slot);
size_t cmgpu_plane_stride(const cmgpu_ctx_t *ctx);           /* in floats */
int    cmgpu_fetch_planar(cmgpu_ctx_t *ctx, unsigned slot, float *host);
/* Wait for everything queued on the context. Polls for 60 us before it blocks. Streams nothing was queued
 * on since the last wait are not asked; when the last thing queued on the compute stream is a tick launch
 * issued after a previous cmgpu_sync, its end is seen through a word of mapped host memory the launch's last
 * CTA writes (no driver call: 14 instead of 18 us launch-to-complete for a 20 ms-block tick). Environment
 * CMGPU_NO_DONE_WORD=1 at cmgpu_ctx_create switches the word off. */
int cmgpu_sync(cmgpu_ctx_t *ctx);
/* Wait until the slot's last fetch (or tick, if none) has completed. */
int cmgpu_slot_wait(cmgpu_ctx_t *ctx, unsigned slot);

/* ---- meters (replaces coolmic_vumeter_read/result/reset, vumeter.c:93-99,138-218) ------
 * "First occurrence" of a peak is kept across ticks by a position key that holds 46 - ceil(log2
 * block_frames) bits of tick number: one meter window (reset to result) may span that many ticks --
 * 2^33 at 4,800-frame ticks. A reset / result of ALL streams starts the count again. */
/* Integer state of streams [first, first+count), waiting for queued ticks first. With
 * `reset` the device state is cleared in the same stream-ordered step. */
int cmgpu_meter_snapshot(cmgpu_ctx_t *ctx, unsigned first, unsigned count,
                         cmgpu_meter_state_t *out, int reset);
int cmgpu_meter_reset(cmgpu_ctx_t *ctx, unsigned first, unsigned count);
/* coolmic_vumeter_result for one stream: CMGPU_ERR_INVAL (and no reset) if no frames were
 * metered; otherwise fills *out (dB computed on the host with the reference's expression and
 * the same libm) and resets the stream's meter. */
int cmgpu_meter_result(cmgpu_ctx_t *ctx, unsigned stream, uint32_t rate, cmgpu_result_t *out);
/* The finaliser alone (vumeter.c:198-212) on a snapshot: pure host arithmetic. */
int cmgpu_finalise(const cmgpu_meter_state_t *state, uint32_t rate, unsigned channels,
                   cmgpu_result_t *out);
/* Device-side raw meter table for collectives: [max_streams][cmgpu_meter_row_u64()] uint64,
 * row = { peak_key[channels], power[channels], frames, 0 }. cmgpu_meter_decode() turns rows
 * gathered from other ranks into states on the host. */
void    *cmgpu_device_meters(cmgpu_ctx_t *ctx);
unsigned cmgpu_meter_row_u64(const cmgpu_ctx_t *ctx);
int      cmgpu_meter_decode(const uint64_t *rows, unsigned count, unsigned channels,
                            cmgpu_meter_state_t *out);

/* coolmic_vumeter_result for MANY streams with one device round trip (a caller that reports every
 * stream each interval, like simple.c:486-491 does for its one stream, would otherwise pay one
 * blocking copy per stream): the rows of [first, first+count) are taken -- copied and, with `reset`,
 * cleared where frames != 0, which is what vumeter.c:198-199,214-215 does per object -- in ONE
 * stream-ordered device step, copied to the host once, and finalised there with the reference's
 * expression and libm (bit-identical doubles). rcs[i] (optional) = CMGPU_OK or CMGPU_ERR_INVAL (no
 * frames metered: results[i] zeroed, row untouched). `states` (optional) receives the integer state.
 * CMGPU_RESULTS_DEVICE_DB: compute the dB values on the device instead (fp64 sqrt/log10 in the take
 * kernel; agrees with the host finaliser to <= 1e-12 relative, NOT guaranteed bit-identical). */
#define CMGPU_RESULTS_DEVICE_DB 0x1u
int cmgpu_meter_results(cmgpu_ctx_t *ctx, unsigned first, unsigned count, uint32_t rate, int reset, unsigned flags,
                        cmgpu_result_t *results, cmgpu_meter_state_t *states, int *rcs);

/* ---- multi-GPU: the one collective of the path ---------------------------------------
 * Streams shard by contiguous range across GPUs, one process (or thread) and one context per GPU, no
 * data-path exchange. Once per reporting interval the per-stream meter rows travel to one rank over
 * NCCL (NVLink / NVSwitch): grouped ncclSend / ncclRecv of the raw integer rows on the compute stream,
 * decoded and finalised on the root with the same host code as cmgpu_meter_results. What arrives at
 * the root is, per stream, exactly what coolmic_vumeter_result() fills (vumeter.h:48-83).
 *
 * A communicator is created from a 128-byte NCCL unique id that rank 0 makes and hands to the other
 * ranks by any means (cmgpu_comm_create), or through a file on a shared file system
 * (cmgpu_comm_create_file: rank 0 writes `path` atomically and removes it once every rank has joined;
 * the others wait up to timeout_ms for it; `path` must be unique to the job -- a file left behind by
 * a job that died before joining would be taken for this job's id, so name it after the launcher, as
 * bench.py does with the launcher's pid, start time and port) -- no PyTorch, no MPI. An existing
 * ncclComm_t can be adopted instead (cmgpu_comm_adopt; not destroyed by cmgpu_comm_destroy). */
typedef struct cmgpu_comm cmgpu_comm_t;
#define CMGPU_COMM_ID_BYTES 128
int           cmgpu_comm_unique_id(unsigned char id[CMGPU_COMM_ID_BYTES]);
cmgpu_comm_t *cmgpu_comm_create(int device, int rank, int nranks, const unsigned char id[CMGPU_COMM_ID_BYTES]);
cmgpu_comm_t *cmgpu_comm_create_file(int device, int rank, int nranks, const char *path, int timeout_ms);
cmgpu_comm_t *cmgpu_comm_adopt(void *nccl_comm, int device);
void          cmgpu_comm_destroy(cmgpu_comm_t *comm);
int           cmgpu_comm_rank(const cmgpu_comm_t *comm);
int           cmgpu_comm_size(const cmgpu_comm_t *comm);
int           cmgpu_comm_nccl_version(void);                       /* e.g. 22703 */
/* Plumbing for multi-rank measurements: a barrier, and an in-place max / sum over ranks. */
int           cmgpu_comm_barrier(cmgpu_comm_t *comm);
int           cmgpu_comm_max(cmgpu_comm_t *comm, double *values, unsigned n);
int           cmgpu_comm_sum(cmgpu_comm_t *comm, double *values, unsigned n);
/* Every rank calls it with its own context (collective). Each rank contributes the rows of its
 * active streams; `reset` as in cmgpu_meter_results. On `root` the outputs (each optional) receive,
 * rank after rank in stream order: results[total], states[total], rcs[total], and counts[nranks] =
 * streams per rank; elsewhere they are ignored. The root returns once it has everything; the other
 * ranks return as soon as their part is queued on their compute stream (they never wait on the host,
 * so a rank behind a slower host link does not hold the others up). */
int cmgpu_gather_results(cmgpu_ctx_t *ctx, cmgpu_comm_t *comm, int root, uint32_t rate, int reset,
                         cmgpu_result_t *results, cmgpu_meter_state_t *states, int *rcs, unsigned *counts);
/* Device milliseconds the last cmgpu_gather_results spent on the ROOT between its first and last
 * stream operation (take + NCCL + copy to the host); 0 on the other ranks. */
float cmgpu_comm_last_gather_ms(const cmgpu_comm_t *comm);

/* ---- adjacent consumers / producers on the device ------------------------------
 * Meter colours: what the app derives from each result with coolmic_util_power2hue / peak2hue /
 * ahsv2argb (util.h:40-47, util.c:59-139), for every stream of [first, first+count) in one kernel over
 * the CURRENT meter rows (no reset): dB on the device, "default" profile hues, packed 0xAARRGGBB.
 * Host doubles are the reference (csrc/host/shim_util.c is bit-exact with util.c); the device's sin()
 * and log10() may differ in the last places, so hues agree to ~1e-13 relative and a colour byte can
 * differ by one only when x*255 lands within that of an integer. Streams with no frames metered get all-zero colours. */
typedef struct cmgpu_colors {
    uint32_t global_power_argb, global_peak_argb;
    uint32_t channel_power_argb[CMGPU_MAX_CHANNELS];
    uint32_t channel_peak_argb[CMGPU_MAX_CHANNELS];
    double   global_power_hue;                      /* for checks: the hue the global power mapped to */
    double   channel_power_hue[CMGPU_MAX_CHANNELS];
} cmgpu_colors_t;
int cmgpu_meter_colors(cmgpu_ctx_t *ctx, unsigned first, unsigned count, double alpha, double saturation,
                       double value, cmgpu_colors_t *out);
/* Tone source: fills the valid frames of a slot ON THE DEVICE the way a snddev_sine-fed capture would
 * (snddev_sine.c:118-150: a cyclic copy of a one-period table), without embedding the driver's data:
 * the table is given at run time (e.g. one period read from the real driver).
 *   sample(stream s, frame f, channel c) = period[(first_frame + f + stream_step*s + channel_step*c) mod n]
 * cmgpu_noise_fill: full-range deterministic noise, sample = (int16) splitmix64(seed ^ s<<40 ^ f<<4 ^ c)
 * (second data set), for streams s with s % every == phase (every >= 1). */
int cmgpu_tone_set_table(cmgpu_ctx_t *ctx, const int16_t *period, unsigned n);
int cmgpu_tone_fill(cmgpu_ctx_t *ctx, unsigned slot, uint64_t first_frame, unsigned first_stream,
                    unsigned stream_step, unsigned channel_step);
int cmgpu_noise_fill(cmgpu_ctx_t *ctx, unsigned slot, uint64_t first_frame, unsigned first_stream,
                     uint64_t seed, unsigned every, unsigned phase);

/* ---- EXTENSION: N -> M integer downmix (no counterpart in libcoolmic-dsp; PARITY UNPINNED) ------
 * BASELINE.json's config 4 names a downmix the reference does not implement.
 * Specified here in the reference's arithmetic style (transform.c:110-123):
 *     out[m] = clamp16(trunc(sum_c (int64)x[c] * weights[m][c] / scale))
 * with the vumeter rules applied to the in_channels input channels AND the out_channels outputs.
 * Checked only against our own CPU restatement (oracle_mix_process); reported separately. A mix
 * context is used like any other: cmgpu_submit (input geometry), cmgpu_process (flags ignored),
 * cmgpu_fetch (output geometry: cmgpu_out_block_stride), cmgpu_meter_* (OUTPUT channels). */
cmgpu_ctx_t *cmgpu_mix_ctx_create(int device, unsigned in_channels, unsigned out_channels,
                                  unsigned max_streams, unsigned ring_slots, unsigned block_frames,
                                  unsigned flags);
/* weights[out_channels][in_channels], scale 1..65535. */
int      cmgpu_stream_set_mix(cmgpu_ctx_t *ctx, unsigned stream, uint16_t scale, const uint16_t *weights);
int      cmgpu_mix_input_snapshot(cmgpu_ctx_t *ctx, unsigned first, unsigned count,
                                  cmgpu_meter_state_t *out, int reset);   /* INPUT-side meters */
unsigned cmgpu_out_channels(const cmgpu_ctx_t *ctx);        /* == cmgpu_channels for gain contexts */
size_t   cmgpu_out_block_stride(const cmgpu_ctx_t *ctx);    /* == cmgpu_block_stride for gain contexts */
void    *cmgpu_host_out_slot(cmgpu_ctx_t *ctx, unsigned slot); /* == cmgpu_host_slot for gain contexts */

/* ---- measurement ---------------------------------------------------------------- */
/* Runs `reps` ticks over slots first_slot .. first_slot+n_slots-1 (cyclically) and returns the
 * device time in milliseconds between CUDA events recorded on the compute stream around them. */
int cmgpu_time_process(cmgpu_ctx_t *ctx, unsigned first_slot, unsigned n_slots, unsigned reps,
                       unsigned flags, float *ms);
/* The same for `cycles` replays of the cached CUDA graph of n_slots ticks (cmgpu_process_cycle). */
int cmgpu_time_cycles(cmgpu_ctx_t *ctx, unsigned first_slot, unsigned n_slots, unsigned cycles,
                      unsigned flags, float *ms);
/* The latency of ONE tick issued alone, as a C caller sees it: `reps` times cmgpu_process + cmgpu_sync on
 * an idle context, wall clock from the call to the return of the wait (median and minimum, in
 * microseconds) -- the honest figure for 20 ms-sized blocks, where one tick is what arrives at a time. */
int cmgpu_time_single_tick(cmgpu_ctx_t *ctx, unsigned slot, unsigned flags, unsigned reps, float *median_us,
                           float *min_us);
/* What the host link of `device` gives for page-locked buffers of `bytes` bytes, `reps` times each:
 * upload alone, download alone, and both at once (GB/s per direction). The end-to-end path moves
 * every sample across the link twice, so `both` is its ceiling; with several ranks calling this at the
 * same moment (after a barrier) each sees its share of the links they have in common. */
int cmgpu_link_probe(int device, size_t bytes, unsigned reps, int write_combined, float *h2d_gbs, float *d2h_gbs,
                     float *both_gbs);
/* Bounds-checking build only (make -C csrc debug -> lib/libcoolmic_b200_dbg.so, for pools without
 * compute-sanitizer): every PCM vector access of the tick kernels is checked against the launching
 * context's rings; returns the number of stray accesses so far (they are dropped, and cmgpu_sync
 * fails). -1 in the normal build. */
int cmgpu_debug_violations(void);
/* Number of kernel launches this context has issued so far. */
uint64_t cmgpu_launch_count(const cmgpu_ctx_t *ctx);
/* How many cmgpu_sync calls saw the compute stream's end through the completion word of its last tick
 * launch (a word of mapped host memory the launch's last CTA writes) instead of asking the driver. */
uint64_t cmgpu_word_waits(const cmgpu_ctx_t *ctx);
/* Name of the kernel variant cmgpu_process would pick for the current shape (for logs). */
const char *cmgpu_kernel_name(const cmgpu_ctx_t *ctx);

/* ---- host-side gain arithmetic self-check (no device work) ------------------------ */
/* The kernel divides by multiplying with a per-(stream,channel) reciprocal. This evaluates
 * that integer recipe on the host for one sample so that tests can prove, exhaustively over
 * all 65,536 inputs, that it equals trunc(x*gain/scale) saturated. Not a data path. */
int cmgpu_recipe_eval(uint16_t gain, uint16_t scale, int16_t x);
/* The same for every x: out[i] = recipe(x = i - 32768), i = 0..65535. */
int cmgpu_recipe_table(uint16_t gain, uint16_t scale, int16_t out[65536]);

#ifdef __cplusplus
}
#endif
#endif
