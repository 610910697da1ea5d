/* oracle/coolmic_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement ("port") of the libcoolmic-dsp transform + vumeter hot path.
 * Every function cites the reference file:line it follows. Pinned against the
 * reference's own object code (oracle/_ref, built by oracle/Makefile) and against
 * tests/golden/ (generated from that object code by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may include, link or load this. The product (libcoolmic-dsp_b200/) never does.
 */
#ifndef COOLMIC_ORACLE_H
#define COOLMIC_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_CHANNELS 16
#define ORACLE_ERROR_NONE     0
#define ORACLE_ERROR_FAULT   (-9)
#define ORACLE_ERROR_INVAL  (-10)

/* Pull callback with the semantics of one iohandle `read` callback
 * (reference include/coolmic-dsp/iohandle.h:41-53): -1 error, 0 nothing now, >0 bytes. */
typedef ssize_t (*oracle_read_cb)(void *userdata, void *buffer, size_t len);

/* Flat result, same members as coolmic_vumeter_result_t (vumeter.h:48-83). */
typedef struct oracle_result {
    int32_t  rc;
    uint32_t rate;
    uint32_t channels;
    int32_t  global_peak;
    uint64_t frames;
    double   global_power;
    int32_t  channel_peak[ORACLE_MAX_CHANNELS];
    double   channel_power[ORACLE_MAX_CHANNELS];
} oracle_result_t;

/* ---- arithmetic kernels ------------------------------------------------------ */

/* transform.c:101-124 (__process) */
void oracle_gain_process(int16_t *samples, size_t frames, unsigned channels,
                         uint16_t scale, const uint16_t *gain);

/* transform.c:195-222 (coolmic_transform_set_master_gain): state in/out */
int oracle_gain_adapt(unsigned stream_channels, unsigned n, uint16_t scale, const uint16_t *gain,
                      uint16_t *state_scale, uint16_t state_gain[ORACLE_MAX_CHANNELS]);

/* Integer meter state between reset and result (vumeter.c:48-56). */
typedef struct oracle_meter {
    int64_t  power[ORACLE_MAX_CHANNELS];
    int16_t  channel_peak[ORACLE_MAX_CHANNELS];
    int16_t  global_peak;
    uint64_t frames;
} oracle_meter_t;

/* vumeter.c:161-177 */
void oracle_meter_accumulate(oracle_meter_t *m, const int16_t *samples, size_t frames, unsigned channels);
/* vumeter.c:189-218; resets *m on success */
int  oracle_meter_finalise(oracle_meter_t *m, uint32_t rate, unsigned channels, oracle_result_t *out);

/* ---- pull-chain objects (framing semantics) ------------------------------------ */

typedef struct oracle_transform {
    oracle_read_cb src;
    void *src_userdata;
    unsigned channels;
    uint16_t scale;
    uint16_t gain[ORACLE_MAX_CHANNELS];
    unsigned char carry[2 * ORACLE_MAX_CHANNELS - 1];
    size_t carry_fill;
} oracle_transform_t;

void    oracle_transform_init(oracle_transform_t *t, unsigned channels, oracle_read_cb src, void *userdata);
/* transform.c:126-165 (__read), including iohandle.c:74-104's read loop on the source */
ssize_t oracle_transform_read(oracle_transform_t *t, void *buffer, size_t len);

typedef struct oracle_vumeter {
    oracle_read_cb src;
    void *src_userdata;
    uint32_t rate;
    unsigned channels;
    unsigned char buffer[2 * ORACLE_MAX_CHANNELS * 32];
    size_t fill;
    oracle_meter_t meter;
} oracle_vumeter_t;

void    oracle_vumeter_init(oracle_vumeter_t *v, uint32_t rate, unsigned channels, oracle_read_cb src, void *userdata);
/* vumeter.c:112-187 */
ssize_t oracle_vumeter_read(oracle_vumeter_t *v, ssize_t maxlen);
/* vumeter.c:189-218 */
int     oracle_vumeter_result(oracle_vumeter_t *v, oracle_result_t *out);

/* ---- whole-buffer conveniences (ctypes entry points) --------------------------- */

/* mem -> transform -> caller, `pull`-byte reads, source hands out <= src_chunk bytes (0 = any). */
long oracle_run_transform(const void *in, size_t in_bytes, unsigned channels,
                          int set_gain, unsigned gain_n, unsigned scale, const uint16_t *gain,
                          size_t src_chunk, size_t pull, void *out, size_t out_cap, int *gain_rc);

/* mem -> vumeter; a result every `result_every` successful reads and one at the end. */
long oracle_run_vumeter(const void *in, size_t in_bytes, uint32_t rate, unsigned channels,
                        size_t src_chunk, long maxlen, unsigned result_every,
                        oracle_result_t *results, size_t results_cap);

/* Batch form used as the GPU parity checker: [n_streams][frames*channels] S16 in place,
 * per-stream scale/gains (already adapted), integer meter state out (accumulated onto *meters). */
void oracle_batch(int16_t *pcm, size_t n_streams, size_t stride_samples, const uint32_t *frames,
                  unsigned channels, const uint16_t *scale, const uint16_t *gain,
                  oracle_meter_t *meters);

/* EXTENSION CHECKER (parity unpinned: the reference has no downmix).
 * Restates include/cmgpu.h's specification of the N -> M mix in the reference's arithmetic style
 * (transform.c:110-123: int64 product, C division truncating toward zero, saturation):
 *   out[f][m] = clamp16(trunc(sum_c (int64)in[f][c] * w[m][c] / scale)),  w: [cout][cin] uint16.
 * Meters (vumeter.c:161-177 rules) accumulate onto *min (inputs) and *mout (outputs) if given. */
void oracle_mix_process(const int16_t *in, size_t frames, unsigned cin, unsigned cout, uint16_t scale,
                        const uint16_t *w, int16_t *out, oracle_meter_t *min, oracle_meter_t *mout);

/* enc_vorbis.c:108-117 (__vorbis_read_data): the encoder-side sample-format stage -- interleaved S16 ->
 * one float plane per channel, planes[c][f] = in[f * channels + c] / 32768.f (plane_stride floats apart).
 * Pinned against the reference's own enc_vorbis.c object code (oracle/_ref/libcoolmic_refenc.so,
 * tests/golden/planar.json). */
void oracle_planar(const int16_t *in, size_t frames, unsigned channels, float *planes, size_t plane_stride);

/* Multi-threaded form of oracle_batch for bench.py's cpu_baseline ("port"); returns seconds. */
double oracle_batch_threads(int16_t *pcm, size_t n_streams, size_t stride_samples, const uint32_t *frames,
                            unsigned channels, const uint16_t *scale, const uint16_t *gain,
                            oracle_meter_t *meters, unsigned n_threads);

/* snddev_sine.c:36-99,118-150: the 1 kHz one-period tables, read cyclically from phase 0.
 * Returns 0 for an unsupported rate. Restated from the table VALUES' generating rule is not
 * possible (they are literal arrays), so this entry point is only available in oracle/_ref;
 * the port exposes the FNV-1a hash used by the golden files instead. */
uint64_t oracle_fnv1a64(const void *data, size_t len);

#ifdef __cplusplus
}
#endif
#endif
