// cmgpu_tables.h -- device table rows and launch arguments shared by the kernels (cmgpu_*.cuh) and
// the host engine (cmgpu_ctx.h). Plain structs, no device code.
#pragma once

#include <cstdint>

namespace cmgpu {

// ---- device tables ---------------------------------------------------------------------

// Per-stream gain recipe, one row per stream (device resident, 16-byte aligned).
// For channel c:   X  = x * mul[c]                      (mul = 2^pre, pre in 0..16)
//                  y  = clamp16((X * (int)mw[c] + ((X & addm[c]) << 32) + (X < 0 ? 2^32 - 1 : 0)) >> 32)
// which equals trunc(x*g/d) for every int16 x wherever the true quotient is inside the clamp
// range and clamps identically outside (proof: DESIGN.md "Exact division"; exhaustive test:
// tests/test_recipe.py through cmgpu_recipe_eval()).
struct GainRow {
    uint32_t mw[16];     // low 32 bits of M = floor(2^k * g/d) + 1
    uint32_t addm[16];   // all-ones when M >= 2^31 (then mulhi_s32 misses +X), else 0
    uint32_t mul[16];    // 2^pre
    uint32_t flags;      // bit 0: identity (scale == 0, or every gain[c] == scale)
    uint32_t pad[3];
};
static_assert(sizeof(GainRow) == 208, "GainRow layout");

constexpr uint32_t kGainIdentity = 1u;   // flags bit 0: y = x for every channel
constexpr uint32_t kGainAddAll = 2u;     // flags bit 1: addm is all-ones for every channel

// How a work item applies its stream's recipe (warp-uniform, chosen per item from flags).
enum GainMode { GM_IDENTITY = 0, GM_MASKED = 1, GM_ADDALL = 2 };

// Meter row per stream: { peak_key[C], power[C], frames, 0 } as uint64.
// peak_key = mag(17 bits) << 47 | (~position & (2^46-1)) << 1 | negative
// so that a 64-bit atomicMax keeps the largest magnitude and, among equals, the earliest
// position -- exactly the strict '>' update of vumeter.c:163. position = tick << pbits | frame.
constexpr int      kKeyMagShift = 47;
constexpr uint64_t kKeyPosMask = (1ull << 46) - 1ull;

struct TickArgs {
    const uint8_t *in;          // slot base (device)
    uint8_t *out;               // == in when working in place
    const uint32_t *frames;     // valid frames per stream, or nullptr = block_frames each
    const GainRow *gains;
    unsigned long long *meters;
    unsigned long long *tick;   // [0] tick sequence number as of the last bump_tick
    uint32_t pbits;             // position = (tick[0] + tick_offset) << pbits | frame
    uint32_t tick_offset;       // the launch's tick number relative to tick[0]
    uint32_t reserved0;
    uint32_t n_streams;
    uint32_t block_frames;
    uint32_t stride_bytes;      // bytes between stream-blocks (multiple of 16)
    uint32_t items_per_block;   // work items (chunks) per stream-block
    uint32_t per_item;          // vectors (fast kernels) or frames (generic kernel) per item
    uint32_t row_u64;           // meter row length in uint64
    uint32_t store;             // write PCM to `out` (0 only for identity streams in place)
    float *planar;              // optional second output: [stream][channel][plane_stride] float = y / 32768.f
    uint32_t plane_stride;      // floats per plane (block_frames rounded up to 4)
    // A span: ONE launch walks n_ticks consecutive ring slots (fused_tick only; 0 or 1 = a plain tick).
    // Work items then number (tick, stream, chunk); tick t of the span sits slot_bytes * t further
    // into both rings and frames_stride * t further into `frames`, and takes the position base
    // tick[0] + tick_offset + t, so the meter keys order its samples after those of tick t - 1.
    uint32_t n_ticks;
    uint32_t frames_stride;
    uint64_t slot_bytes;
    // Work distribution (fused_tick, 32-lane groups). nullptr: static -- group g takes items g, g + n,
    // g + 2n, ... (n = groups of the grid). Otherwise a counter that only ever grows: the first n items
    // are dealt out statically, every further one is claimed with an atomicAdd when a group starts on
    // its current item (claimed number = stride + counter value - work_base), so SMs that run a little
    // faster simply take more items and the launch has no tail of stragglers -- measured on config 5,
    // one slot in place: 4.38 -> 3.87 ms, 0.88 -> 0.99 of the copy rate. Every processed item makes
    // exactly one claim, so a launch advances the counter by its item count and the host knows the
    // next launch's base without ever resetting anything (no memset between launches: ticks stay bare
    // kernel launches and may overlap). Launches that may be in flight together use different counters.
    unsigned int *work;
    uint32_t work_base;
    // Completion word (tick_end): the launch's last CTA writes done_gen to done_flag, a word of mapped HOST
    // memory, so that a host thread waiting for this launch sees its end by polling its own memory instead
    // of asking the driver (cudaStreamQuery costs 1.4 us a call: for a 20 ms-block tick that is a tenth of
    // the launch-to-complete time). done_count counts the CTAs that have finished; nullptr = no word.
    unsigned int *done_count;
    unsigned int *done_flag;
    uint32_t done_gen;
};

// Per-stream mix recipe (device table row).
struct MixRow {
    uint16_t w[16][16];      // w[m][c]
    uint32_t magic;          // M = floor(2^(31+l) / scale) + 1, l = ceil(log2(scale))
    uint32_t shift;          // 31 + l
    uint32_t pad[2];
    // 8 -> 2 fast path: for output m and channel pair p the bytes { lo(w[2p]), lo(w[2p+1]), hi(w[2p]), hi(w[2p+1]) },
    // so that two dp2a per pair give sum(x * lo) and sum(x * hi) straight from the packed input word
    uint32_t packed[2][4];
};
static_assert(sizeof(MixRow) == 560, "MixRow layout");

struct MixArgs {
    const uint8_t *in;            // [stream][frames*CIN] S16, stride_in bytes apart
    uint8_t *out;                 // [stream][frames*COUT] S16, stride_out bytes apart
    const uint32_t *frames;
    const MixRow *rows;
    unsigned long long *meters_in;    // rows of (2*CIN+2) uint64
    unsigned long long *meters_out;   // rows of (2*COUT+2) uint64
    unsigned long long *tick;
    uint32_t pbits, tick_offset, reserved0;
    uint32_t n_streams, block_frames;
    uint32_t stride_in, stride_out;
    uint32_t items_per_block, per_item;   // frames per item
    uint32_t cin, cout;
    unsigned int *work;           // work-claim counter, see TickArgs::work
    uint32_t work_base;
    unsigned int *done_count;     // completion word, see TickArgs::done_flag
    unsigned int *done_flag;
    uint32_t done_gen;
};

}  // namespace cmgpu
