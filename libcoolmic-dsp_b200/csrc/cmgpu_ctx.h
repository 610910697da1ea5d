// cmgpu_ctx.h -- the engine's context object and error plumbing, shared by the translation units of
// the library (cmgpu.cu: ring, ticks, meters; cmgpu_comm.cu: the NCCL meter gather; cmgpu_post.cu:
// on-device consumers of meter results and the on-device tone source). Not installed; the public
// interface is include/cmgpu.h.
#pragma once

#include "cmgpu_tables.h"

#include "../../include/cmgpu.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>       // header-only; the ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached

#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

struct cmgpu_ctx;

namespace cmgpu {

// thread-local text behind cmgpu_last_error(); returns `code`
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

int take_rows_locked(struct ::cmgpu_ctx *c, unsigned first, unsigned count, int reset, bool device_db, uint32_t rate);
int ensure_take_buffers_locked(struct ::cmgpu_ctx *c);
// Host side of a result gather: raw rows -> integer states -> finalised results (each output optional),
// spread over a few threads when there are many rows (65,536 streams x log10/sqrt is ~20 ms on one).
void finalise_rows(const uint64_t *rows, size_t count, unsigned row_u64, unsigned channels, uint32_t rate,
                   cmgpu_result_t *results, cmgpu_meter_state_t *states, int *rcs);

}  // namespace cmgpu

// NVTX range over an entry point of the C ABI (tracing).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
#define CMGPU_TRACE(name) NvtxRange nvtx_range__(name)

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return cmgpu::fail(CMGPU_ERR_GENERIC, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

using cmgpu::GainRow;
using cmgpu::MixRow;

struct cmgpu_ctx {
    int device = 0;
    unsigned channels = 0, max_streams = 0, active = 0, slots = 0, block_frames = 0, flags = 0;
    size_t stride = 0, slot_bytes = 0;
    unsigned row_u64 = 0, pbits = 0;
    uint64_t launches = 0;
    int num_sms = 0;
    // environment hooks, read ONCE at cmgpu_ctx_create (never on the launch path)
    bool env_no_pdl = false, env_no_span = false, env_static = false, env_span_by_tick = false, env_span_by_stream = false, env_span_single = false;
    // work-claim counters (TickArgs::work): launch number n uses counter n % kWorkCounters; each only grows,
    // work_base[] is its value when the next launch that uses it starts (host arithmetic, no resets)
    static constexpr unsigned kWorkCounters = 8;
    unsigned int *d_work = nullptr;
    uint32_t work_base[kWorkCounters] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned work_next = 0;

    uint8_t *d_in = nullptr, *d_out = nullptr;     // rings
    uint8_t *h_ring = nullptr;                     // pinned staging ring
    float *d_planar = nullptr;                     // optional [slot][stream][channel][plane_stride] float
    size_t plane_stride = 0, planar_slot_floats = 0;

    // EXTENSION (downmix contexts only): N -> M mix, separate output geometry, input-side meters
    unsigned out_channels = 0;                     // 0: ordinary gain context
    size_t stride_out = 0, slot_bytes_out = 0;
    uint8_t *h_ring_out = nullptr;
    MixRow *d_mix = nullptr;
    std::vector<MixRow> h_mix;
    bool mix_dirty = false;
    unsigned long long *d_meters_in = nullptr;
    unsigned row_in_u64 = 0;
    GainRow *d_gains = nullptr;
    std::vector<GainRow> h_gains;
    std::vector<uint16_t> h_scale, h_gain;         // adapted settings, [stream], [stream][channels]
    unsigned dirty_lo = 0, dirty_hi = 0;           // gain rows to upload: [lo, hi)
    unsigned long long *d_meters = nullptr;
    unsigned long long *d_tick = nullptr;          // [0] tick sequence number, [1] CTA completion ticket
    uint32_t *d_frames = nullptr;                  // [slots][max_streams]
    std::vector<char> has_frames;
    std::vector<uint64_t> scratch;                 // snapshot staging
    // taking rows for many streams at once (cmgpu_post.cu): device staging, its pinned host mirror, and
    // the outputs of the optional on-device consumers; allocated on first use
    unsigned long long *d_take = nullptr;
    uint64_t *h_take = nullptr;
    cmgpu_result_t *d_results = nullptr;
    void *d_colors = nullptr;
    int16_t *d_tone = nullptr;                     // run-time one-period table of the tone source
    unsigned tone_len = 0;
    // per-slot pinned staging of the frame counts (cmgpu_slot_set_frames copies, then uploads from here)
    uint32_t *h_frames = nullptr;
    std::vector<cudaEvent_t> ev_frames;

    cudaStream_t s_up = nullptr, s_cmp = nullptr, s_down = nullptr;
    // Whoever queues anything on the compute stream takes it through cmp(): that forgets the completion
    // word of the last tick launch (tail_gen), because the stream's tail is then something else.
    cudaStream_t cmp()
    {
        tail_gen.store(0, std::memory_order_relaxed);
        return s_cmp;
    }
    // Completion word of tick launches (TickArgs::done_flag): one word of mapped host memory the last CTA
    // of a launch writes its generation to. tail_gen = generation of the launch that is the LAST thing
    // queued on s_cmp (0: the tail is something else, or nothing was launched): cmgpu_sync then polls the
    // word instead of the driver. up_seq / down_seq count what was queued on the side streams, *_synced
    // what cmgpu_sync has already waited for: idle streams are not asked at all.
    unsigned int *d_done_count = nullptr;
    volatile unsigned int *h_done = nullptr;
    unsigned int *d_done_flag = nullptr;              // device alias of h_done
    uint32_t done_gen_next = 1;
    std::atomic<uint32_t> tail_gen{0};
    std::atomic<uint64_t> word_waits{0};
    std::atomic<bool> idle_hint{true};                // a cmgpu_sync has returned since the last tick launch
    std::atomic<uint64_t> up_seq{0}, down_seq{0}, up_synced{0}, down_synced{0};
    std::vector<cudaEvent_t> ev_up, ev_cmp, ev_down;
    // Cross-stream ordering is queued only where it orders something: a tick waits for a slot's upload /
    // download only if one was issued since the slot's last tick, and a slot's "ticks done" event is
    // recorded when an upload or download first asks for it. Back-to-back ticks on resident data are
    // then back-to-back kernel launches, which is what lets them overlap (launch_begin).
    std::vector<uint8_t> up_pending, down_pending, cmp_unrecorded;
    // A slot whose device copy still equals what was uploaded from its pinned staging slot needs no
    // download: from_staging[slot] = the last submit came from the staging slot; slot_dirty[slot] = some
    // tick (or fill) has written the slot's PCM on the device since. Pass-through streams -- the
    // reference's default state, transform.c:107-108 -- are metered without a byte coming back.
    std::vector<uint8_t> from_staging, slot_dirty;
    uint64_t bytes_h2d = 0, bytes_d2h = 0;         // PCM bytes copied by cmgpu_submit / cmgpu_fetch so far
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    std::mutex mu;

    // cached CUDA graph of a cycle of ticks (launch-bound small-buffer regime): the ticks are
    // independent nodes spread over a few side streams, so their latencies overlap
    std::vector<cudaStream_t> s_fork;
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_join;
    cudaGraphExec_t graph = nullptr;
    unsigned graph_first = 0, graph_n = 0, graph_flags = 0;
    uint64_t graph_launches = 0;                   // kernel nodes in the cached graph
    uint64_t config_gen = 0, graph_gen = ~0ull;    // bumped whenever launch arguments may change
    // tick numbering: ticks issued by plain / span launches since the device counter was last bumped
    uint32_t pending_ticks = 0;
    // Overlap rule of programmatic dependent launch. A chain of overlapping launches is open on s_cmp
    // while every launch since the last full dependency was marked dependent-launchable; in_chain[slot]
    // says which slots those launches touch. A small grid does not fill the GPU, so ANY launch of the
    // chain may still be running when the next one starts -- not just the previous one.
    bool chain_open = false;
    std::vector<uint8_t> in_chain;
    unsigned last_first = ~0u;                     // slot of the last tick launch; ~0u: something else was queued last

    // how many active streams need which gain mode; the tick runs in the cheapest common one
    unsigned n_mode[3] = {0, 0, 0};                // GM_IDENTITY / GM_MASKED / GM_ADDALL
    bool classes_dirty = true;

    // launch plan (depends on shape only)
    int plan_g = 32;              // lanes per item (fast kernels); 0 = frame-per-lane generic kernel; -1 = any_tick
    int plan_lanes = 32;          // any_tick: lanes of a warp that take part
    // long stream-blocks, opt-in: TMA-staged kernel (cmgpu_tma.cuh) with its own item geometry
    bool tma = false;
    uint32_t tma_items = 1, tma_per_item = 0;
    int tma_grid_cap[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    uint32_t plan_items = 1, plan_per_item = 0;
    int grid_cap[3][2] = {{0, 0}, {0, 0}, {0, 0}};   // resident CTAs per (gain mode, meter) kernel
    int span_grid_cap[3][2] = {{0, 0}, {0, 0}, {0, 0}};   // ... of the stream-major span kernel
    char kname[64] = "";
    char kname_last[64] = "";     // set when the last launch used a kernel other than the context's plan (span_tick)
};


#ifdef __CUDACC__
#define CMGPU_HD __host__ __device__ inline
#else
#define CMGPU_HD inline
#endif

namespace cmgpu {

// One raw meter row { peak_key[C], power[C], frames, 0 } -> the integer state the reference keeps in
// struct coolmic_vumeter (vumeter.c:48-56). Host and device (the take kernel) run the same code.
CMGPU_HD void decode_row(const uint64_t *row, unsigned C, cmgpu_meter_state_t *st)
{
    st->frames = row[2 * C];
    st->global_peak = 0;
    st->reserved[0] = st->reserved[1] = st->reserved[2] = 0;
    uint32_t best_mag = 0;
    uint64_t best_order = 0;
    for (unsigned c = 0; c < CMGPU_MAX_CHANNELS; c++) {
        st->power[c] = 0;
        st->channel_peak[c] = 0;
    }
    for (unsigned c = 0; c < C; c++) {
        const uint64_t key = row[c];
        const uint32_t mag = (uint32_t)(key >> kKeyMagShift);
        const uint64_t pos = (~(key >> 1)) & kKeyPosMask;
        const int v = (key & 1ull) ? -(int)mag : (int)mag;
        st->channel_peak[c] = (int16_t)v;
        st->power[c] = (int64_t)row[C + c];
        // global peak (vumeter.c:163-168): first sample in interleaved order with the overall
        // largest magnitude = the channel winner with the smallest (frame, channel)
        if (mag) {
            const uint64_t order = pos * 16u + c;
            if (mag > best_mag || (mag == best_mag && order < best_order)) {
                best_mag = mag;
                best_order = order;
                st->global_peak = (int16_t)v;
            }
        }
    }
}

CMGPU_HD double power_db(double mean_square)
{
    // vumeter.c:204-205: p = 20*log10(sqrt(p)/32768); p = fmin(p, 0)
    double p = 20. * log10(sqrt(mean_square) / 32768.);
    return fmin(p, 0.);
}

// vumeter.c:198-212 on a decoded state. On the host this is the reference's expression with the
// reference's libm (bit-identical doubles); on the device the same expression with CUDA's fp64
// sqrt / log10.
CMGPU_HD int finalise_state(const cmgpu_meter_state_t *st, uint32_t rate, unsigned channels, cmgpu_result_t *out)
{
    if (!st->frames)
        return CMGPU_ERR_INVAL;                     // vumeter.c:198-199 (nothing written)
    static_assert(sizeof(cmgpu_result_t) == 192 && sizeof(cmgpu_result_t) % 8 == 0, "cmgpu_result_t layout");
#ifdef __CUDA_ARCH__
    for (unsigned i = 0; i < sizeof(cmgpu_result_t) / 8; i++)       // padding included, like the host's memset
        reinterpret_cast<uint64_t *>(out)[i] = 0;
#else
    memset(out, 0, sizeof(*out));
#endif
    out->rate = rate;
    out->channels = channels;
    out->frames = st->frames;
    out->global_peak = st->global_peak;
    int64_t all = 0;
    for (unsigned ch = 0; ch < channels; ch++) {
        all += st->power[ch];
        out->channel_peak[ch] = st->channel_peak[ch];
        // vumeter.c:203: signed integer division first, then the conversion to double
        out->channel_power[ch] = power_db((double)(st->power[ch] / (int64_t)st->frames));
    }
    // vumeter.c:209: unsigned division by frames * channels
    out->global_power = power_db((double)((uint64_t)all / (uint64_t)(st->frames * (uint64_t)channels)));
    return CMGPU_OK;
}

}  // namespace cmgpu
