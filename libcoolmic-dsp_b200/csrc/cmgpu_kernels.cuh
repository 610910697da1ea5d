// cmgpu_kernels.cuh -- the fused transform + vumeter kernels (sm_100a).
//
// One launch ("tick") walks every active stream-block of a ring slot exactly once:
//   128-bit coalesced load -> per-sample gain (exact truncating division by multiplying with a
//   per-(stream,channel) reciprocal) -> saturate -> 128-bit store, and in the same pass the
//   meter's per-channel first-occurrence peak + exact sum of squares, reduced with warp
//   shuffles and merged into the per-stream state with 64-bit atomics.
//
// What is being computed (reference file:line):
//   transform.c:110-123   y = clamp16(trunc((int64)x * gain[c] / scale))
//   vumeter.c:161-175     peak[c] = first sample with the largest |y|; power[c] += y*y
//
// HBM-bound integer work: no tensor cores (no contraction exists), no shared memory (no reuse).
// Algorithmic traffic: 2 B read + 2 B written per sample; the meter adds none.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "cmgpu_tables.h"

#ifndef CMGPU_UNROLL
#define CMGPU_UNROLL 4      // vectors per load batch of a lane (two batches are held in registers)
#endif
#ifndef CMGPU_MIN_CTAS
#define CMGPU_MIN_CTAS 3    // resident 256-thread CTAs per SM the fast kernels are compiled for
#endif
// 4, 8 and 16 channels keep 4-8 recipes + 4-8 64-bit power sums per lane: more registers per thread
#ifndef CMGPU_UNROLL_WIDE
#define CMGPU_UNROLL_WIDE 4
#endif
#ifndef CMGPU_MIN_CTAS_WIDE
#define CMGPU_MIN_CTAS_WIDE 2
#endif
#ifndef CMGPU_SATPACK_ALL
#define CMGPU_SATPACK_ALL 1   // 0: the 80-register 1- and 2-channel kernels clamp with two VIMNMX and pack with PRMT instead of I2IP
#endif
// 8-lane-group kernels (stream-blocks of <= 1 KiB)
#ifndef CMGPU_G8_CTAS
#define CMGPU_G8_CTAS 4
#endif
#ifndef CMGPU_G8_CTAS_STEREO
#define CMGPU_G8_CTAS_STEREO 3   // 80 registers: at 4 (64 registers) the stereo 8-lane kernels spilled 52-96 bytes
#endif
#ifndef CMGPU_G8_CTAS_WIDE
#define CMGPU_G8_CTAS_WIDE 2
#endif
#ifndef CMGPU_PLANAR_CTAS
#define CMGPU_PLANAR_CTAS 2   // resident CTAs per SM of the kernels that also write float planes
#endif

namespace cmgpu {

// ---- small helpers -----------------------------------------------------------------------

// Streaming accesses. Measured on B200 (cfg2, tools/sweep_mode.sh): the evict-first
// hint (.cs) costs the LOADS 1-3 % against a plain or non-coherent load, while the stores keep it.
// `nc` = the input ring is not written by this launch (separate output ring): the read-only path
// (LDG.E.CONSTANT) is then legal and the fastest; in place the plain load is used.
#ifndef CMGPU_ST_HINT
#define CMGPU_ST_HINT 0      // 0 = .cs (evict first), 1 = default
#endif
// Written as volatile asm with a memory clobber on purpose: where the compiler is free to move
// them it sinks a batch of loads down to its first use (seen in SASS), which throws away the
// prefetch distance the kernels are built around and costs up to 8 %.
#ifdef CMGPU_BOUNDS_CHECK
// Debug build (make debug -> libcoolmic_b200_dbg.so; compute-sanitizer is not available on every pool):
// every PCM vector access is checked against the extents of the launching context's rings. A stray
// access is counted and dropped; cmgpu_sync / cmgpu_debug_violations report the count.
struct DebugBounds {
    unsigned long long lo[2], hi[2];
};
__device__ DebugBounds g_dbg_bounds;
__device__ unsigned int g_dbg_violations;
__device__ __forceinline__ bool dbg_ok(const void *p)
{
    const unsigned long long a = (unsigned long long)p;
    const bool ok = ((a & 15ull) == 0) && ((a >= g_dbg_bounds.lo[0] && a + 16 <= g_dbg_bounds.hi[0]) ||
                                           (a >= g_dbg_bounds.lo[1] && a + 16 <= g_dbg_bounds.hi[1]));
    if (!ok)
        atomicAdd(&g_dbg_violations, 1u);
    return ok;
}
#endif

__device__ __forceinline__ uint4 ld_stream(const uint8_t *p, bool nc = false)
{
    uint4 v;
#ifdef CMGPU_BOUNDS_CHECK
    if (!dbg_ok(p))
        return make_uint4(0, 0, 0, 0);
#endif
    if (nc)
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else
        asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_stream(uint8_t *p, uint4 v)
{
#ifdef CMGPU_BOUNDS_CHECK
    if (!dbg_ok(p))
        return;
#endif
#if CMGPU_ST_HINT == 0
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
}

__device__ __forceinline__ uint64_t shfl_xor64(unsigned mask, uint64_t v, int off)
{
    uint32_t lo = __shfl_xor_sync(mask, (uint32_t)v, off);
    uint32_t hi = __shfl_xor_sync(mask, (uint32_t)(v >> 32), off);
    return ((uint64_t)hi << 32) | lo;
}

// Asynchronous 16-byte copies from global to shared memory (L2 only), for the kernels that stage a
// lane's vectors through slots of its own instead of registers (span_tick, any_tick).
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// Hides where a value came from, so that recomputing something from it is really recomputed
// (and its first copy does not have to stay in registers across the hot loop).
__device__ __forceinline__ uint64_t opaque(uint64_t v)
{
    asm volatile("" : "+l"(v));
    return v;
}

__device__ __forceinline__ uint64_t make_key(uint32_t mag, uint64_t pos)
{
    return mag ? (((uint64_t)mag << kKeyMagShift) | (((~pos) & kKeyPosMask) << 1)) : 0ull;
}

// Which tick a launch is decides where its samples sit in the position order of the meter keys:
//     tick = tick[0] + tick_offset (+ the tick's index inside a span).
// tick[0] lives on the device and is advanced only by bump_tick (once per replay of a captured cycle,
// and before one when plain ticks were issued since); plain and span launches get their place from
// the host as tick_offset = ticks issued since the last bump. No launch modifies tick[0] itself, so
// consecutive launches may overlap (see launch_begin).
__device__ __forceinline__ uint64_t tick_pos_base(const unsigned long long *tick, uint32_t offset, uint32_t pbits)
{
    const unsigned long long t = *reinterpret_cast<const volatile unsigned long long *>(tick) + offset;
    return ((uint64_t)t << pbits) & kKeyPosMask;
}
__device__ __forceinline__ uint64_t tick_begin(const TickArgs &a) { return tick_pos_base(a.tick, a.tick_offset, a.pbits); }

// Programmatic dependent launch: every tick kernel lets the NEXT launch of its stream start as soon
// as its own CTAs have all started, so the next tick's CTAs take over SM by SM while this tick's last
// work items drain -- no idle tail, no launch gap, no ramp between ticks. A grid that does not fill
// the GPU lets the launch after the next start too, and so on: any number of consecutive launches
// may be in flight at once. Only launches the host marked as independent of EVERY launch that may
// still be running use it (slots that none of them touches, or a read-only input ring); meter
// updates are atomics on order-free position keys. launch_end keeps completion in stream order: a
// launch does not finish before the one before it.
__device__ __forceinline__ void launch_begin() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void launch_end() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// The end of a tick launch: completion in stream order, then the completion word (TickArgs::done_flag).
// The count is touched only after griddepcontrol.wait, i.e. after the previous launch has finished
// entirely, so consecutive launches -- overlapping or not -- can share one counter; every CTA's writes
// are fenced before it counts itself, so whoever sees the word sees the launch's results.
template <typename Args>          // TickArgs or MixArgs: both carry done_count / done_flag / done_gen
__device__ __forceinline__ void tick_end(const Args &a)
{
    launch_end();
    if (a.done_flag != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(a.done_count, 1u) == gridDim.x - 1u) {
                *a.done_count = 0;
                *reinterpret_cast<volatile unsigned int *>(a.done_flag) = a.done_gen;
            }
        }
    }
}

// Work distribution of the warp-per-item loops (TickArgs::work): the item after `item` for this warp --
// item + stride in the static order, or stride + (the number this warp claimed when it started on
// `item`). `claimed` is valid in lane 0.
__device__ __forceinline__ uint64_t next_item(uint64_t item, uint64_t stride, const unsigned int *work, uint32_t work_base,
                                              uint32_t claimed)
{
    if (work != nullptr)
        return stride + (uint32_t)(__shfl_sync(0xffffffffu, claimed, 0) - work_base);
    return item + stride;
}
__device__ __forceinline__ uint32_t claim_item(unsigned int *work, uint32_t lane)
{
    return (work != nullptr && lane == 0) ? atomicAdd(work, 1u) : 0u;
}

// One sample through the gain recipe (see GainRow).
struct Recipe {
    int mw;
    int addm;
    int mul;
};

// Returns the UNSATURATED result; callers saturate (cvt.pack.sat in the fast kernels).
// One 64-bit multiply-add, of which only the high word is kept (IMAD.HI with a register-pair addend):
//     y = (X * (int)mw + { hi: X or X & addm, lo: X >> 31 }) >> 32
// The high word of the addend supplies the "+X" that a signed multiply by mw >= 2^31 misses; the low
// word is 2^32-1 for negative X, which turns the floor of the shift into a ceiling, i.e. rounds
// toward zero like the reference's division (X*M/2^32 is never an integer for x != 0 inside the
// clamp range: DESIGN.md "Exact division"). Both halves of the addend come for one SHF, where a
// separate "+ (X >>> 31)" on the high word cost an add plus a zeroed pair register per sample.
template <int GM>
__device__ __forceinline__ int apply_gain_raw(int x, const Recipe &r)
{
    if (GM == GM_IDENTITY)
        return x;
    const int X = x * r.mul;
    const uint32_t hi = (GM == GM_ADDALL) ? (uint32_t)X : (uint32_t)(X & r.addm);
    const long long add = (long long)(((unsigned long long)hi << 32) | (uint32_t)(X >> 31));
    return (int)(((long long)X * (long long)r.mw + add) >> 32);
}

template <int GM>
__device__ __forceinline__ int apply_gain(int x, const Recipe &r)
{
    return max(min(apply_gain_raw<GM>(x, r), 32767), -32768);
}

// Two 32-bit results -> one word of two saturated int16 (I2IP.S16.S32.SAT: clamp and pack at once).
__device__ __forceinline__ uint32_t pack_sat16(int hi, int lo)
{
    uint32_t r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
}

// ---- fast kernels: channel counts that divide (or are a multiple of) one 16-byte vector -----
//
// Work item = (stream, chunk of `per_item` consecutive 16-byte vectors of its stream-block),
// owned by a group of G lanes (G = 8, 16 or 32; small stream-blocks use small groups so that no
// lane idles). Lane l of the group visits vectors v0 + l + G*i, i = 0,1,...: every load/store
// instruction of a group touches one contiguous G*16-byte run.
//
// Per lane the 8 sample slots of its vectors map to fixed channels, so the meter keeps 8
// (key, power) register pairs and never indexes dynamically. In-loop peak key (32 bit):
//   mag << 16 | (0xFFFF - i)   -> max() keeps the largest magnitude, then the earliest visit.
// Order of time inside an item is (i, lane, slot), which the reduction preserves by widening
// to the 64-bit position key before lanes are combined.

template <int C>
struct Shape {
    static constexpr int kPerLane = C < 8 ? C : 8;      // distinct channels a lane sees
    static constexpr int kLanesPerFrame = C <= 8 ? 1 : C / 8;
    static constexpr int kFramesPerVec8 = 8 / kPerLane; // frames per vector when C <= 8
};

// Per-kernel tuning: vectors per load batch and the resident CTAs per SM ptxas must fit.
// 8-lane groups only ever walk stream-blocks of <= 64 vectors (<= 8 per lane): short batches,
// and as many resident groups as possible so that one wave covers all streams of a tick.
template <int C, int G>
struct Tune {
    static constexpr int kUnroll = (G == 8) ? 4 : (C >= 4 ? CMGPU_UNROLL_WIDE : CMGPU_UNROLL);
    static constexpr int kMinCtas = (G == 8) ? (C >= 4 ? CMGPU_G8_CTAS_WIDE : (C == 2 ? CMGPU_G8_CTAS_STEREO : CMGPU_G8_CTAS)) : (C >= 4 ? CMGPU_MIN_CTAS_WIDE : CMGPU_MIN_CTAS);
    static constexpr bool kSatPack = CMGPU_SATPACK_ALL ? true : ((G == 8) || (C >= 4));
};

// SATPACK: saturate and pack two results with one I2IP and meter what was packed (one ALU
// instruction per sample less; needs a few more live registers, so only the 128-register and the
// 8-lane kernels use it -- in the 80-register stereo kernel it spills and loses 8 %).
// SIGNKEY: the in-loop peak key also carries the sample's sign in bit 0 (below the step bits, so it
// never decides a comparison); the caller passes radd = (0x7fff - step) << 1. Used where the
// written PCM cannot be re-read for the sign (the TMA kernel's stores are asynchronous).
// MERGE: all slots of a channel share ONE running key (kmax[channel], not kmax[slot]): equal keys of
// one vector tie, and item_publish finds the first of them by looking at the winning vector again.
// Fewer live registers (2 instead of 8 in the stereo kernel) and pairs of max() fold into VIMNMX3.
template <int C, int GM, bool METER, bool MASKED, bool SATPACK, bool SIGNKEY = false, bool MERGE = false>
__device__ __forceinline__ uint4 do_vector(uint4 w, const Recipe (&rc)[Shape<C>::kPerLane], uint32_t radd,
                                           uint32_t (&kmax)[8], uint64_t (&pacc)[Shape<C>::kPerLane], int nvalid)
{
    constexpr int P = Shape<C>::kPerLane;
    uint32_t in[4] = {w.x, w.y, w.z, w.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int k0 = 2 * j, k1 = 2 * j + 1;
        const int x0 = (int)(short)(in[j] & 0xffffu);
        const int x1 = (int)in[j] >> 16;
        int m0, m1;
        if (GM == GM_IDENTITY) {
            o[j] = in[j];
            m0 = x0;
            m1 = x1;
        } else if (!SATPACK) {
            int y0 = apply_gain<GM>(x0, rc[k0 % P]);
            int y1 = apply_gain<GM>(x1, rc[k1 % P]);
            if (MASKED) {
                if (k0 >= nvalid) y0 = x0;
                if (k1 >= nvalid) y1 = x1;
            }
            o[j] = ((uint32_t)y0 & 0xffffu) | ((uint32_t)y1 << 16);
            m0 = y0;
            m1 = y1;
        } else {
            int r0 = apply_gain_raw<GM>(x0, rc[k0 % P]);
            int r1 = apply_gain_raw<GM>(x1, rc[k1 % P]);
            if (MASKED) {
                // samples past the valid frames pass through untouched
                if (k0 >= nvalid) r0 = x0;
                if (k1 >= nvalid) r1 = x1;
            }
            o[j] = pack_sat16(r1, r0);                      // saturate both, pack: one instruction
            m0 = (int)(short)(o[j] & 0xffffu);              // what was actually written, for the meter
            m1 = (int)o[j] >> 16;
        }
        if (MASKED) {
            // ... and are invisible to the meter
            if (k0 >= nvalid) m0 = 0;
            if (k1 >= nvalid) m1 = 0;
        }
        if (METER) {
            const uint32_t a0 = (uint32_t)abs(m0), a1 = (uint32_t)abs(m1);
            const int q0 = MERGE ? k0 % P : k0, q1 = MERGE ? k1 % P : k1;
            if (SIGNKEY) {
                kmax[q0] = max(kmax[q0], (a0 << 16) + radd + ((uint32_t)m0 >> 31));
                kmax[q1] = max(kmax[q1], (a1 << 16) + radd + ((uint32_t)m1 >> 31));
            } else {
                kmax[q0] = max(kmax[q0], (a0 << 16) + radd);
                kmax[q1] = max(kmax[q1], (a1 << 16) + radd);
            }
            // exact: |y| <= 32768, so y*y <= 2^30; one IMAD.WIDE with 64-bit accumulate per sample
            pacc[k0 % P] += (uint64_t)((int64_t)m0 * (int64_t)m0);
            pacc[k1 % P] += (uint64_t)((int64_t)m1 * (int64_t)m1);
        }
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// Optional second output: the encoder-side sample-format stage of
// enc_vorbis.c:108-117 -- de-interleave and `sample / 32768.f` into one float plane per channel,
// which is what vorbis_analysis_buffer() wants. The division by 2^15 is exact in binary32, so
// multiplying by 2^-15 gives bit-identical floats (checked against the reference's own enc_vorbis.c
// object code: tests/golden/planar.json, oracle/ref_enc_harness.c). Only the PLANAR instantiations of the kernel call
// it (a run-time flag in the plain kernels cost them 2 %).
template <int C>
__device__ __forceinline__ void store_planar(float *planar, uint32_t plane_stride, uint32_t s, uint32_t v, uint4 o,
                                             int nvalid)
{
    constexpr int P = Shape<C>::kPerLane;
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        f[2 * j] = (float)(int)(short)(w[j] & 0xffffu) * (1.0f / 32768.0f);
        f[2 * j + 1] = (float)((int)w[j] >> 16) * (1.0f / 32768.0f);
    }
    const uint32_t cbase = (C == 16) ? (v & 1u) * 8u : 0u;
    const uint32_t frame0 = (C <= 8) ? v * (uint32_t)Shape<C>::kFramesPerVec8 : (v >> 1);
    float *base = planar + ((size_t)s * C + cbase) * plane_stride + frame0;
    if (nvalid >= 8 && C == 1) {
        reinterpret_cast<float4 *>(base)[0] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4 *>(base)[1] = make_float4(f[4], f[5], f[6], f[7]);
    } else if (nvalid >= 8 && C == 2) {
        *reinterpret_cast<float4 *>(base) = make_float4(f[0], f[2], f[4], f[6]);
        *reinterpret_cast<float4 *>(base + plane_stride) = make_float4(f[1], f[3], f[5], f[7]);
    } else if (nvalid >= 8 && C == 4) {
#pragma unroll
        for (int c = 0; c < 4; c++)
            *reinterpret_cast<float2 *>(base + (size_t)c * plane_stride) = make_float2(f[c], f[4 + c]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < nvalid)
                base[(size_t)(k % P) * plane_stride + (k / P)] = f[k];
    }
}

// What a lane needs to know about one work item.
struct Item {
    const uint8_t *src;      // the lane's first vector of the item
    size_t base;             // byte offset of the item's stream-block in the rings (slot of the span included)
    uint32_t tk;             // which tick of the span the item belongs to (0 for a plain tick)
    uint32_t s;              // stream
    uint32_t first;          // index of that vector in the stream-block
    uint32_t n_i;            // how many entirely valid vectors the lane visits
    uint32_t tail_vec;       // vector straddling the end of the valid frames, if this lane owns it
    uint32_t tail_step;      // ... its step number in the item
    int tail_valid;          // ... its number of valid samples (0: the lane owns no such vector)
    uint32_t count_frames;   // frames to add to the stream's counter (first chunk, lane 0 only)
};

template <int C, int G>
__device__ __forceinline__ bool item_setup(const TickArgs &a, uint64_t item, uint64_t n_items, uint32_t gl, Item &it)
{
    it.n_i = 0;
    it.tail_valid = 0;
    it.count_frames = 0;
    it.s = 0;
    it.tk = 0;
    it.base = 0;
    it.first = 0;
    it.tail_vec = it.tail_step = 0;
    it.src = a.in;
    if (item >= n_items)
        return false;
    // (the host keeps n_ticks * n_streams * items_per_block below 2^32: 32-bit divisions, none for
    //  one item per block and a plain tick)
    const uint32_t item32 = (uint32_t)item;
    const uint32_t sv = a.items_per_block == 1 ? item32 : item32 / a.items_per_block;   // (tick, stream)
    const uint32_t chunk = item32 - sv * a.items_per_block;
    uint32_t tk = 0, s = sv;
    if (a.n_ticks > 1) {
        tk = sv / a.n_streams;
        s = sv - tk * a.n_streams;
    }
    const uint32_t nfr = a.frames ? min(__ldg(a.frames + (size_t)tk * a.frames_stride + s), a.block_frames) : a.block_frames;
    const uint32_t valid_bytes = nfr * (uint32_t)(2 * C);
    const uint32_t nvec = (valid_bytes + 15u) >> 4;
    const uint32_t v0 = chunk * a.per_item;
    const uint32_t v1 = min(v0 + a.per_item, nvec);
    const size_t base = (size_t)tk * a.slot_bytes + (size_t)s * a.stride_bytes;
    it.s = s;
    it.tk = tk;
    it.base = base;
    it.first = v0 + gl;
    it.src = a.in + base + (size_t)it.first * 16;
    it.count_frames = (chunk == 0 && gl == 0) ? nfr : 0;
    if (v0 < v1) {
        const uint32_t vfull = min(v1, valid_bytes >> 4);        // vectors [v0, vfull) are entirely valid
        it.n_i = it.first < vfull ? (vfull - it.first + (G - 1)) / G : 0;
        if (vfull < v1 && (vfull << 4) < valid_bytes && ((vfull - v0) % G) == gl) {
            it.tail_vec = vfull;
            it.tail_step = (vfull - v0) / G;
            it.tail_valid = (int)((valid_bytes - (vfull << 4)) >> 1);
        }
    }
    return true;
}

template <int C, int GM>
__device__ __forceinline__ void load_recipes(const TickArgs &a, uint32_t s, uint32_t gl,
                                             Recipe (&rc)[Shape<C>::kPerLane])
{
    constexpr int P = Shape<C>::kPerLane;
#pragma unroll
    for (int c = 0; c < P; c++) {
        rc[c].mw = rc[c].addm = 0;
        rc[c].mul = 1;
    }
    if (GM != GM_IDENTITY) {
        const GainRow *g = a.gains + s;
        // C == 16: even lanes own channels 0-7, odd lanes 8-15 (v0 and G are even)
        const int cbase = (C == 16) ? (int)(gl & 1u) * 8 : 0;
#pragma unroll
        for (int c = 0; c < P; c++) {
            rc[c].mw = (int)__ldg(&g->mw[cbase + c]);
            rc[c].mul = (int)__ldg(&g->mul[cbase + c]);
            if (GM == GM_MASKED)
                rc[c].addm = (int)__ldg(&g->addm[cbase + c]);
        }
    }
}

// Meter epilogue of one item: combine the group's lanes with shuffles, publish one channel per lane.
// The lanes' keys are combined while they are still 32 bits wide -- magnitude << 16 | ~(index of the
// vector among the group's vectors of this item), one SHFL and one max per round instead of two and
// a 64-bit compare -- and only the winner is widened to a position key. For frames narrower than a
// vector (C < 8) that key names the winning VECTOR; the publishing lane looks at it again and takes
// the channel's first sample with the winning magnitude (and, always, the sample's sign).
// STRIDE: vectors between a lane's consecutive steps (G in fused_tick; the consumer count in tma_tick,
// where G is only the width of the shuffle tree). Needs per_item + STRIDE <= 65,536 vectors.
template <int C, int G, bool MERGE = false, int STRIDE = G>
__device__ __forceinline__ void item_publish(const TickArgs &a, const Item &it, uint32_t gl, unsigned gmask,
                                             const uint32_t (&kmax)[8], uint64_t (&pacc)[Shape<C>::kPerLane])
{
    static_assert(MERGE || C >= 8, "narrow frames keep one running key per channel (do_vector<MERGE>)");
    constexpr int P = Shape<C>::kPerLane;
    constexpr int S = (C <= 8) ? 8 / P : 1;      // slots of one channel in a vector (= frames per vector)
    // The tick number is requested first, so that its latency hides behind the shuffle rounds instead of
    // adding to the re-read's (cfg5: 3 %). The 8-lane groups used to read it late, in the publishing
    // lanes only (6 % faster when they walked config 3's spans); spans with many streams now run in
    // span_tick, and what is left for them -- single ticks, where the dependent latency chain of the
    // epilogue is the kernel's tail -- wants the early read as well.
    constexpr bool kEarlyTick = true;
    uint64_t pos_base = 0;
    if (kEarlyTick)
        pos_base = tick_pos_base(a.tick, a.tick_offset + it.tk, a.pbits);
    uint32_t k32[P];
#pragma unroll
    for (int c = 0; c < P; c++) {
        const uint32_t best = kmax[c];
        const uint32_t idx = (0xffffu - (best & 0xffffu)) * (uint32_t)STRIDE + gl;
        k32[c] = (best >> 16) ? ((best & 0xffff0000u) | (0xffffu - idx)) : 0u;
    }
#pragma unroll
    for (int off = G / 2; off >= Shape<C>::kLanesPerFrame; off >>= 1) {
#pragma unroll
        for (int c = 0; c < P; c++) {
            k32[c] = max(k32[c], __shfl_xor_sync(gmask, k32[c], off));
            pacc[c] += shfl_xor64(gmask, pacc[c], off);
        }
    }
    // lane L publishes one channel: C <= 8 -> channel L; C == 16 -> channel 8*(L&1) + (L>>1)
    uint32_t key32 = 0;
    uint64_t pw = 0;
    const int sel = (C == 16) ? (int)(gl >> 1) : (int)gl;
#pragma unroll
    for (int c = 0; c < P; c++) {
        if (sel == c) {
            key32 = k32[c];
            pw = pacc[c];
        }
    }
    const int ch = (C == 16) ? (int)((gl & 1u) * 8u + (gl >> 1)) : (int)gl;
    __syncwarp(gmask);       // make the item's stores visible to the lane that re-reads a sample
    if ((int)gl < C) {
        unsigned long long *row = a.meters + (size_t)it.s * a.row_u64;
        if (key32) {
            const uint32_t mag = key32 >> 16;
            const uint32_t v = (it.first - gl) + (0xffffu - (key32 & 0xffffu));       // the winning vector
            const uint32_t frame = (C <= 8) ? v * (uint32_t)Shape<C>::kFramesPerVec8 : (v >> 1);
            if (!kEarlyTick)
                pos_base = tick_pos_base(a.tick, a.tick_offset + it.tk, a.pbits);
            const uint64_t pos = pos_base + frame;
            // the winning sample, re-read from where it was written: its sign, and for narrow frames
            // which of the vector's frames it is (slots past the valid frames of a tail vector come
            // after every valid one, so they can only match after the real winner)
            const volatile int16_t *y = reinterpret_cast<const volatile int16_t *>(a.out + it.base);
            uint32_t j = 0;
            int yv = y[(size_t)frame * C + ch];
            if (S > 1) {
                bool found = (uint32_t)abs(yv) == mag;
#pragma unroll
                for (int q = 1; q < S; q++) {
                    const int cand = y[((size_t)frame + q) * C + ch];
                    if (!found && (uint32_t)abs(cand) == mag) {
                        found = true;
                        j = (uint32_t)q;
                        yv = cand;
                    }
                }
            }
            atomicMax(row + ch, (unsigned long long)(make_key(mag, pos + j) | (yv < 0 ? 1ull : 0ull)));
        }
        if (pw)
            atomicAdd(row + C + ch, (unsigned long long)pw);
    }
    if (it.count_frames)
        atomicAdd(a.meters + (size_t)it.s * a.row_u64 + 2 * C, (unsigned long long)it.count_frames);
}

// One kernel per (channel shape, group width, gain mode, meter on/off): every work item of a
// launch runs the same straight-line code. The host picks the cheapest mode that is exact for
// every active stream (identity rows carry the unity recipe, add-all rows an all-ones mask), so
// a tick is always exactly ONE launch.
//
// Each group walks its items (grid-stride) as a software pipeline:
//   * inside an item, double-buffered batches: a lane requests UNROLL consecutive vectors of its
//     stride back-to-back (the warp's requests then cover UNROLL*G*16 contiguous bytes at once,
//     which is what HBM rows like), one whole batch ahead of the batch it computes on; two
//     register sets alternate roles, nothing is copied;
//   * across items, the next item's recipes and first batch are requested BEFORE the current
//     item's meter epilogue (shuffles, one dependent re-read, atomics), whose latency they hide.
// NC: the launch does not write its input ring (separate output ring), so the loads may take the
// read-only path; in place they are plain loads. A compile-time choice: as a run-time flag it cost
// the 8-channel kernel 5 %.
// (the instantiations with the float-plane second output get 128 registers: at 80 they spilled 100-250
//  bytes, and a kernel that writes three times what it reads is bound by its stores, not by occupancy)
template <int C, int G, int GM, bool METER, bool PLANAR = false, bool NC = false>
__global__ void __launch_bounds__(256, PLANAR ? CMGPU_PLANAR_CTAS : Tune<C, G>::kMinCtas) fused_tick(const __grid_constant__ TickArgs a)
{
    constexpr int P = Shape<C>::kPerLane;
    constexpr int UNROLL = Tune<C, G>::kUnroll;
    constexpr size_t kStep = (size_t)G * 16;              // bytes between a lane's consecutive vectors
    // Pass-through streams in place -- the reference's default state, transform.c:107-108 -- write nothing:
    // the host only picks this instantiation (identity, input ring == output ring, no planes) with store
    // == 0, so the stores and the registers they keep alive are compiled out of it.
    constexpr bool kStores = !(GM == GM_IDENTITY && !NC && !PLANAR);
    launch_begin();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gl = threadIdx.x & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(uint32_t)(G - 1)));
    const uint32_t groups_per_cta = 256 / G;
    const uint64_t n_items = (uint64_t)a.n_streams * a.items_per_block * (a.n_ticks > 1 ? a.n_ticks : 1u);
    const uint64_t stride = (uint64_t)gridDim.x * groups_per_cta;

    uint32_t kmax[8];
    uint64_t pacc[P];
#pragma unroll
    for (int k = 0; k < 8; k++)
        kmax[k] = 0;
#pragma unroll
    for (int c = 0; c < P; c++)
        pacc[c] = 0;

    // the output ring mirrors the input ring: one uniform distance instead of a second pointer per lane
    const ptrdiff_t out_delta = a.out - a.in;
    constexpr bool nc = NC;
    uint4 bufA[UNROLL], bufB[UNROLL];
#define CMGPU_LOAD_BATCH(buf, it, b)                                                    \
    _Pragma("unroll") for (int u = 0; u < UNROLL; u++)                                  \
        buf[u] = ld_stream((it).src + (size_t)((b) * UNROLL + u) * kStep, nc);
#define CMGPU_DO_BATCH(buf, it, b)                                                      \
    _Pragma("unroll") for (int u = 0; u < UNROLL; u++) {                                \
        const uint32_t iu = (b) * UNROLL + u;                                           \
        const uint4 o = do_vector<C, GM, METER, false, Tune<C, G>::kSatPack, false, true>(buf[u], rc, 0xffffu - iu, kmax, pacc, 8); \
        if (kStores && a.store)                                                                    \
            st_stream(dstp + (size_t)iu * kStep, o);                                    \
        if (PLANAR)                                                                     \
            store_planar<C>(a.planar, a.plane_stride, (it).s, (it).first + iu * G, o, 8); \
    }

#define CMGPU_LOAD_REM(buf, it, b)                                                      \
    _Pragma("unroll") for (int u = 0; u < UNROLL - 1; u++)                              \
        if ((b) * UNROLL + u < (it).n_i)                                                \
            buf[u] = ld_stream((it).src + (size_t)((b) * UNROLL + u) * kStep, nc);
#define CMGPU_DO_REM(buf, it, b)                                                        \
    _Pragma("unroll") for (int u = 0; u < UNROLL - 1; u++) {                            \
        const uint32_t iu = (b) * UNROLL + u;                                           \
        if (iu < (it).n_i) {                                                            \
            const uint4 o = do_vector<C, GM, METER, false, Tune<C, G>::kSatPack, false, true>(buf[u], rc, 0xffffu - iu, kmax, pacc, 8); \
            if (kStores && a.store)                                                                \
                st_stream(dstp + (size_t)iu * kStep, o);                                \
            if (PLANAR)                                                                 \
                store_planar<C>(a.planar, a.plane_stride, (it).s, (it).first + iu * G, o, 8); \
        }                                                                               \
    }

    uint64_t item = (uint64_t)blockIdx.x * groups_per_cta + threadIdx.x / G;
    Item cur;
    Recipe rc[P];
    bool have = item_setup<C, G>(a, item, n_items, gl, cur);
    // 8-lane groups walk stream-blocks of at most 64 vectors: a lane's (at most 8) vectors are all
    // requested at once, predicated, into the two register sets -- one memory latency per item.
#define CMGPU_LOAD_ALL(it)                                                              \
    _Pragma("unroll") for (int u = 0; u < UNROLL; u++) {                                \
        if ((uint32_t)u < (it).n_i)                                                     \
            bufA[u] = ld_stream((it).src + (size_t)u * kStep, nc);                          \
        if ((uint32_t)(UNROLL + u) < (it).n_i)                                          \
            bufB[u] = ld_stream((it).src + (size_t)(UNROLL + u) * kStep, nc);               \
    }
#define CMGPU_DO_ALL(buf, it, base)                                                     \
    _Pragma("unroll") for (int u = 0; u < UNROLL; u++) {                                \
        const uint32_t iu = (base) + u;                                                 \
        if (iu < (it).n_i) {                                                            \
            const uint4 o = do_vector<C, GM, METER, false, Tune<C, G>::kSatPack, false, true>(buf[u], rc, 0xffffu - iu, kmax, pacc, 8); \
            if (kStores && a.store)                                                                \
                st_stream(dstp + (size_t)iu * kStep, o); \
            if (PLANAR)                                                                 \
                store_planar<C>(a.planar, a.plane_stride, (it).s, (it).first + iu * G, o, 8); \
        }                                                                               \
    }
    if (have) {
        load_recipes<C, GM>(a, cur.s, gl, rc);
        if (G == 8) {
            CMGPU_LOAD_ALL(cur)
        } else if (cur.n_i >= (uint32_t)UNROLL) {
            CMGPU_LOAD_BATCH(bufA, cur, 0u)
        } else if (cur.n_i) {
            CMGPU_LOAD_REM(bufA, cur, 0u)
        }
    }
    while (have) {
        // (opaque: otherwise ptxas re-derives the address from src at every store, 8 instructions each)
        uint8_t *const dstp = reinterpret_cast<uint8_t *>(opaque(reinterpret_cast<uint64_t>(cur.src) + (uint64_t)out_delta));
        const uint32_t nb = (G == 8) ? 0u : cur.n_i / UNROLL;   // full batches of this lane; batch 0 is in flight
        // what is left of the lane's vectors after them (< UNROLL) is one more, predicated, batch of the
        // same pipeline: requested while the last full batch is worked on, into the register set that is
        // free then (measured on the 48,000-frame stream-blocks of config 5, whose eighth work item ends
        // in such a remainder: requested only after the loop, its latency was exposed once per item)
        const bool rem = (G != 8) && nb * UNROLL < cur.n_i;
        // work distribution, see TickArgs::work: the next item's number is requested now and used after the loops
        uint32_t claimed = 0;
        if (G != 8 && a.work != nullptr && lane == 0)
            claimed = atomicAdd(a.work, 1u);
        if (G == 8) {
            CMGPU_DO_ALL(bufA, cur, 0u)
            CMGPU_DO_ALL(bufB, cur, (uint32_t)UNROLL)
        }
        for (uint32_t b = 0; b < nb; b += 2) {
            if (b + 1 < nb) {
                CMGPU_LOAD_BATCH(bufB, cur, b + 1)
            } else if (rem) {
                CMGPU_LOAD_REM(bufB, cur, b + 1)
            }
            CMGPU_DO_BATCH(bufA, cur, b)
            if (b + 2 < nb) {
                CMGPU_LOAD_BATCH(bufA, cur, b + 2)
            } else if (b + 2 == nb && rem) {
                CMGPU_LOAD_REM(bufA, cur, b + 2)
            }
            if (b + 1 < nb) {
                CMGPU_DO_BATCH(bufB, cur, b + 1)
            }
        }
        if (rem) {
            if (nb & 1u) {
                CMGPU_DO_REM(bufB, cur, nb)
            } else {
                CMGPU_DO_REM(bufA, cur, nb)
            }
        }
        // Everything about the item that the loops above did not need (stream, tail vector, frame
        // count) is derived again here instead of being carried through them in registers.
        if (!PLANAR)
            item_setup<C, G>(a, opaque(item), n_items, gl, cur);
        if (cur.tail_valid) {
            // the one vector that straddles the end of the valid frames
            const size_t off = cur.base + (size_t)cur.tail_vec * 16;
            const uint4 w = ld_stream(a.in + off, nc);
            const uint4 o = do_vector<C, GM, METER, true, Tune<C, G>::kSatPack, false, true>(w, rc, 0xffffu - cur.tail_step, kmax, pacc, cur.tail_valid);
            if (kStores && a.store)
                st_stream(a.out + off, o);
            if (PLANAR)
                store_planar<C>(a.planar, a.plane_stride, cur.s, cur.tail_vec, o, cur.tail_valid);
        }

        // next item: recipes and first batch go out before this item's epilogue
        if (G != 8 && a.work != nullptr)
            item = stride + (uint32_t)(__shfl_sync(0xffffffffu, claimed, 0) - a.work_base);   // the first `stride` items were dealt out statically
        else
            item += stride;
        Item nxt;
        Recipe rcn[P];
        have = item_setup<C, G>(a, item, n_items, gl, nxt);
        if (have) {
            load_recipes<C, GM>(a, nxt.s, gl, rcn);
            if (G == 8) {
                CMGPU_LOAD_ALL(nxt)
            } else if (nxt.n_i >= (uint32_t)UNROLL) {
                CMGPU_LOAD_BATCH(bufA, nxt, 0u)
            } else if (nxt.n_i) {
                CMGPU_LOAD_REM(bufA, nxt, 0u)
            }
        }
        if (METER) {
            item_publish<C, G, true>(a, cur, gl, gmask, kmax, pacc);
#pragma unroll
            for (int k = 0; k < 8; k++)
                kmax[k] = 0;
#pragma unroll
            for (int c = 0; c < P; c++)
                pacc[c] = 0;
        }
        cur = nxt;
#pragma unroll
        for (int c = 0; c < P; c++)
            rc[c] = rcn[c];
    }
#undef CMGPU_LOAD_BATCH
#undef CMGPU_DO_BATCH
#undef CMGPU_LOAD_REM
#undef CMGPU_DO_REM
#undef CMGPU_LOAD_ALL
#undef CMGPU_DO_ALL
    tick_end(a);
}

// Float planes for any channel count: one scalar store per sample (secondary path).
__device__ __noinline__ void store_planar_any(float *planar, uint32_t plane_stride, uint32_t s, int C, uint32_t v,
                                              uint4 o, int nvalid)
{
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
    const uint64_t sample0 = (uint64_t)v * 8u;
    uint32_t frame = (uint32_t)(sample0 / (uint32_t)C);
    int ch = (int)(sample0 - (uint64_t)frame * (uint32_t)C);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int y = (k & 1) ? ((int)w[k >> 1] >> 16) : (int)(short)(w[k >> 1] & 0xffffu);
        if (k < nvalid)
            planar[((size_t)s * C + ch) * plane_stride + frame] = (float)y * (1.0f / 32768.0f);
        if (++ch == C) {
            ch = 0;
            frame++;
        }
    }
}

// ---- any_tick: channel counts that do not tile a 16-byte vector (3, 5, 6, 7, 9..15) ------------
//
// Same vector pipeline as fused_tick, for frames that straddle vectors. A warp uses L <= 32 lanes,
// L a multiple of C / gcd(8, C) (30 for 3/5/6/10/12/15 channels, 28 for 7/14, 27 for 9, 26 for 13,
// 22 for 11): the sample offset between a lane's consecutive vectors, 8*L, is then a multiple of
// C, so each of the lane's 8 sample slots keeps ONE channel for the whole item -- the hot loop is
// the 8-channel one (do_vector<8>: a recipe, a peak key and a power sum per slot), only the
// recipes are gathered per lane and the epilogue folds slots into channels by a per-lane map.
constexpr int kAnyDepth = 8;                               // vectors of a lane in flight
constexpr int kAnySmemBytes = kAnyDepth * 256 * 16;        // [kAnyDepth][256] uint4: slot (i mod depth) of thread tid

// Round 2: the lane's vectors are no longer held in two register batches but streamed through a ring of
// kAnyDepth private shared-memory slots with cp.async -- 8 vectors per lane in flight whatever the
// register budget, and the ring simply runs on into the NEXT item before the current item's epilogue,
// which the register version could not afford (it spilled).
template <int GM, bool METER, bool NC = false>
__global__ void __launch_bounds__(256, 2) any_tick(const __grid_constant__ TickArgs a, const int C, const int L)
{
    extern __shared__ uint4 any_ring[];
    launch_begin();
    const uint32_t lane = threadIdx.x & 31u;
    const bool active = (int)lane < L;
    const size_t kStep = (size_t)L * 16;
    const uint64_t n_items = (uint64_t)a.n_streams * a.items_per_block;
    const uint64_t stride = (uint64_t)gridDim.x * 8u;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(any_ring) + threadIdx.x * 16u;
    const ptrdiff_t out_delta = a.out - a.in;

    struct AnyItem {
        const uint8_t *src;       // the lane's first vector
        size_t base;              // byte offset of the stream-block
        uint32_t s, v0, v1, vfull, first, n_i, valid_bytes;
    };
    auto setup = [&](uint64_t item, AnyItem &it) -> bool {
        it.src = a.in;
        it.base = 0;
        it.s = it.v0 = it.v1 = it.vfull = it.first = it.n_i = it.valid_bytes = 0;
        if (item >= n_items)
            return false;
        const uint32_t s = (uint32_t)(item / a.items_per_block);
        const uint32_t chunk = (uint32_t)(item - (uint64_t)s * a.items_per_block);
        const uint32_t nfr = a.frames ? min(__ldg(a.frames + s), a.block_frames) : a.block_frames;
        it.valid_bytes = nfr * (uint32_t)(2 * C);
        const uint32_t nvec = (it.valid_bytes + 15u) >> 4;
        it.s = s;
        it.v0 = chunk * a.per_item;
        it.v1 = min(it.v0 + a.per_item, nvec);
        if (METER && chunk == 0 && lane == 0 && nfr)
            atomicAdd(a.meters + (size_t)s * a.row_u64 + 2 * C, (unsigned long long)nfr);
        it.base = (size_t)s * a.stride_bytes;
        it.first = it.v0 + lane;
        it.vfull = min(it.v1, it.valid_bytes >> 4);
        it.n_i = (active && it.v0 < it.v1 && it.first < it.vfull) ? (it.vfull - it.first + (uint32_t)L - 1u) / (uint32_t)L : 0u;
        it.src = a.in + it.base + (size_t)it.first * 16;
        return true;
    };
    // vector i of the item into ring slot i mod depth; always one commit, so that "all but the newest
    // depth - 1 groups have landed" means "vector i is there" at step i
    auto request = [&](const AnyItem &it, uint32_t i) {
        if (i < it.n_i) {
            const uint8_t *p = it.src + (size_t)i * kStep;
#ifdef CMGPU_BOUNDS_CHECK
            if (dbg_ok(p))
#endif
                cp_async16(smem0 + (i & (uint32_t)(kAnyDepth - 1)) * 4096u, p);
        }
        cp_async_commit();
    };
    // the lane's slot -> channel map and recipes
    auto gather = [&](const AnyItem &it, int (&chan)[8], Recipe (&rc)[8]) {
        int ch = (int)(((uint64_t)it.first * 8u) % (uint32_t)C);
        const GainRow *g = a.gains + it.s;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            chan[k] = ch;
            rc[k].mw = rc[k].addm = 0;
            rc[k].mul = 1;
            if (GM != GM_IDENTITY) {
                rc[k].mw = (int)__ldg(&g->mw[ch]);
                rc[k].mul = (int)__ldg(&g->mul[ch]);
                if (GM == GM_MASKED)
                    rc[k].addm = (int)__ldg(&g->addm[ch]);
            }
            ch = ch + 1 == C ? 0 : ch + 1;
        }
    };

    uint64_t item = (uint64_t)blockIdx.x * 8u + (threadIdx.x >> 5);
    AnyItem cur;
    int chan[8];
    Recipe rc[8];
    bool have = setup(item, cur);
    if (have) {
#pragma unroll
        for (int i = 0; i < kAnyDepth; i++)
            request(cur, (uint32_t)i);
        gather(cur, chan, rc);
    }
    while (have) {
        const uint32_t claimed = claim_item(a.work, lane);
        uint32_t kmax[8];
        uint64_t pacc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            kmax[k] = 0;
            pacc[k] = 0;
        }
        // The ring is walked in rounds of kAnyDepth vectors, fully unrolled: slot offsets are constants and
        // the two addresses run along as pointers (the first version of this loop, one vector per
        // iteration with computed slots, spent 30 instructions per vector on its own bookkeeping and was
        // issue-bound at the old kernel's speed).
        const uint8_t *ld_p = cur.src + (size_t)kAnyDepth * kStep;      // the next vector to request
        uint8_t *st_p = const_cast<uint8_t *>(cur.src) + out_delta;
        const uint4 *const mine = any_ring + threadIdx.x;
        uint32_t i = 0;
        const uint32_t n_round = cur.n_i & ~(uint32_t)(kAnyDepth - 1);
        for (; i < n_round; i += kAnyDepth) {
#pragma unroll
            for (int u = 0; u < kAnyDepth; u++) {
                cp_async_wait<kAnyDepth - 1>();
                const uint4 w = mine[u * 256];
                const uint4 o = do_vector<8, GM, METER, false, true>(w, rc, 0xffffu - i - (uint32_t)u, kmax, pacc, 8);
                if (a.store)
                    st_stream(st_p, o);
                if (a.planar)
                    store_planar_any(a.planar, a.plane_stride, cur.s, C, cur.first + (i + (uint32_t)u) * (uint32_t)L, o, 8);
                if (i + (uint32_t)(kAnyDepth + u) < cur.n_i) {
#ifdef CMGPU_BOUNDS_CHECK
                    if (dbg_ok(ld_p))
#endif
                        cp_async16(smem0 + (uint32_t)u * 4096u, ld_p);
                }
                cp_async_commit();
                ld_p += kStep;
                st_p += kStep;
            }
        }
        for (; i < cur.n_i; i++) {          // what is left of the lane's vectors (< kAnyDepth): nothing more to request
            cp_async_wait<0>();
            const uint4 w = mine[(i & (uint32_t)(kAnyDepth - 1)) * 256u];
            const uint4 o = do_vector<8, GM, METER, false, true>(w, rc, 0xffffu - i, kmax, pacc, 8);
            if (a.store)
                st_stream(st_p, o);
            if (a.planar)
                store_planar_any(a.planar, a.plane_stride, cur.s, C, cur.first + i * (uint32_t)L, o, 8);
            st_p += kStep;
        }
        if (active && cur.v0 < cur.v1 && cur.vfull < cur.v1 && (cur.vfull << 4) < cur.valid_bytes &&
            ((cur.vfull - cur.v0) % (uint32_t)L) == lane) {
            // the one vector that straddles the end of the valid frames
            const uint32_t step = (cur.vfull - cur.v0) / (uint32_t)L;
            const int nvalid = (int)((cur.valid_bytes - (cur.vfull << 4)) >> 1);
            const uint4 w = ld_stream(a.in + cur.base + (size_t)cur.vfull * 16, NC);
            const uint4 o = do_vector<8, GM, METER, true, true>(w, rc, 0xffffu - step, kmax, pacc, nvalid);
            if (a.store)
                st_stream(a.out + cur.base + (size_t)cur.vfull * 16, o);
            if (a.planar)
                store_planar_any(a.planar, a.plane_stride, cur.s, C, cur.vfull, o, nvalid);
        }

        // the next item: its first vectors go out before this item's epilogue
        item = next_item(item, stride, a.work, a.work_base, claimed);
        AnyItem nxt;
        const bool have_n = setup(item, nxt);
        if (have_n) {
#pragma unroll
            for (int i = 0; i < kAnyDepth; i++)
                request(nxt, (uint32_t)i);
        }

        if (METER && cur.v0 < cur.v1) {
            // ---- epilogue: fold the lane's slots by its channel map, combine lanes, publish ----
            // Keys stay 32 bits wide until a channel's winner is known: magnitude << 16 | ~(index of the
            // sample's vector among the item's vectors * 8 + slot) -- one SHFL + max per round, and ONE
            // division by C (sample index -> frame) per channel instead of one per slot.
            const uint64_t pos_base = tick_begin(a);
            uint32_t key8[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t idx = ((0xffffu - (kmax[k] & 0xffffu)) * (uint32_t)L + lane) * 8u + (uint32_t)k;
                key8[k] = (kmax[k] >> 16) ? ((kmax[k] & 0xffff0000u) | (0xffffu - idx)) : 0u;
            }
            uint32_t key32 = 0;
            uint64_t pw = 0;
            for (int c = 0; c < C; c++) {
                uint32_t kc = 0;
                uint64_t pc = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (chan[k] == c) {
                        kc = max(kc, key8[k]);
                        pc += pacc[k];
                    }
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    kc = max(kc, __shfl_xor_sync(0xffffffffu, kc, off));
                    pc += shfl_xor64(0xffffffffu, pc, off);
                }
                if ((int)lane == c) {
                    key32 = kc;
                    pw = pc;
                }
            }
            __syncwarp();
            if ((int)lane < C) {
                unsigned long long *row = a.meters + (size_t)cur.s * a.row_u64;
                if (key32) {
                    const uint32_t mag = key32 >> 16;
                    const uint64_t sample = (uint64_t)cur.v0 * 8u + (0xffffu - (key32 & 0xffffu));
                    const uint32_t frame = (uint32_t)(sample / (uint32_t)C);
                    const volatile int16_t *y = reinterpret_cast<const volatile int16_t *>(a.out + cur.base);
                    const int yv = y[(size_t)frame * C + lane];
                    atomicMax(row + lane, (unsigned long long)(make_key(mag, pos_base + frame) | (yv < 0 ? 1ull : 0ull)));
                }
                if (pw)
                    atomicAdd(row + C + lane, (unsigned long long)pw);
            }
        }
        if (have_n)
            gather(nxt, chan, rc);
        cur = nxt;
        have = have_n;
    }
    cp_async_wait<0>();
    tick_end(a);
}

// Advances the tick sequence number: at the end of a captured cycle (the graph orders it after
// every tick node) and, from the host, before a cycle when plain ticks were issued since the last one.
static __global__ void bump_tick(unsigned long long *tick, unsigned n)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
        atomicAdd(tick, (unsigned long long)n);
}

// ---- generic kernel: any channel count 1..16 -------------------------------------------------
//
// Work item = (stream, chunk of `per_item` frames), one warp per item, lane l visits frames
// f0 + l + 32*i and walks the frame's channels with 16-bit accesses. Used for channel counts
// that do not tile a 16-byte vector (3, 5, 6, 7, 9..15) and as an independently written
// cross-check of the fast kernels (CMGPU_FORCE_GENERIC).

template <int GM, bool METER>
__device__ __forceinline__ void run_item_generic(const TickArgs &a, int C, uint32_t s, uint32_t f0, uint32_t f1,
                                                 uint32_t lane)
{
    const size_t base = (size_t)s * a.stride_bytes;
    const int16_t *in = reinterpret_cast<const int16_t *>(a.in + base);
    int16_t *out = reinterpret_cast<int16_t *>(a.out + base);

    Recipe rc[16];
    uint32_t kmax[16];
    uint64_t pacc[16];
#pragma unroll
    for (int c = 0; c < 16; c++) {
        kmax[c] = 0;
        pacc[c] = 0;
        rc[c].mw = rc[c].addm = 0;
        rc[c].mul = 1;
        if (GM != GM_IDENTITY && c < C) {
            const GainRow *g = a.gains + s;
            rc[c].mw = (int)__ldg(&g->mw[c]);
            rc[c].addm = (int)__ldg(&g->addm[c]);
            rc[c].mul = (int)__ldg(&g->mul[c]);
        }
    }

    uint32_t i = 0;
    for (uint32_t f = f0 + lane; f < f1; f += 32, i++) {
        const size_t o = (size_t)f * C;
        const uint32_t radd = 0xffffu - i;
#pragma unroll
        for (int c = 0; c < 16; c++) {
            if (c < C) {
                const int x = in[o + c];
                const int y = apply_gain<GM>(x, rc[c]);
                if (a.store)
                    out[o + c] = (int16_t)y;
                if (a.planar)       // planes of lane = frame are coalesced as they are
                    a.planar[((size_t)s * C + c) * a.plane_stride + f] = (float)y * (1.0f / 32768.0f);
                if (METER) {
                    const uint32_t m = (uint32_t)abs(y);
                    kmax[c] = max(kmax[c], (m << 16) + radd);
                    pacc[c] += (uint64_t)(m * m);
                }
            }
        }
    }
    if (!METER)
        return;

    const uint64_t pos_base = tick_begin(a);
    uint64_t key = 0, pw = 0;
#pragma unroll
    for (int c = 0; c < 16; c++) {
        if (c < C) {
            const uint32_t mag = kmax[c] >> 16;
            const uint32_t it = 0xffffu - (kmax[c] & 0xffffu);
            uint64_t k = make_key(mag, pos_base + (f0 + lane + 32u * it));
            uint64_t p = pacc[c];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                k = max(k, shfl_xor64(0xffffffffu, k, off));
                p += shfl_xor64(0xffffffffu, p, off);
            }
            if ((int)lane == c) {
                key = k;
                pw = p;
            }
        }
    }
    __syncwarp();
    if ((int)lane < C) {
        unsigned long long *row = a.meters + (size_t)s * a.row_u64;
        if (key) {
            const uint64_t pos = (~(key >> 1)) & kKeyPosMask;
            const uint32_t frame = (uint32_t)(pos - pos_base);
            const volatile int16_t *y = out;
            const int yv = y[(size_t)frame * C + lane];
            atomicMax(row + lane, (unsigned long long)(key | (yv < 0 ? 1ull : 0ull)));
        }
        if (pw)
            atomicAdd(row + C + lane, (unsigned long long)pw);
    }
}

template <int GM, bool METER>
__global__ void __launch_bounds__(128) generic_tick(const __grid_constant__ TickArgs a, const int C)
{
    launch_begin();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_per_cta = 128 / 32;
    const uint64_t n_items = (uint64_t)a.n_streams * a.items_per_block;
    const uint64_t stride = (uint64_t)gridDim.x * warps_per_cta;

    uint32_t claimed = 0;
    for (uint64_t item = (uint64_t)blockIdx.x * warps_per_cta + threadIdx.x / 32; item < n_items;
         item = next_item(item, stride, a.work, a.work_base, claimed)) {
        claimed = claim_item(a.work, lane);
        const uint32_t si = (uint32_t)(item / a.items_per_block);
        const uint32_t chunk = (uint32_t)(item - (uint64_t)si * a.items_per_block);
        const uint32_t s = si;
        const uint32_t nfr = a.frames ? min(__ldg(a.frames + s), a.block_frames) : a.block_frames;
        const uint32_t f0 = chunk * a.per_item;
        const uint32_t f1 = min(f0 + a.per_item, nfr);

        if (METER && chunk == 0 && lane == 0 && nfr)
            atomicAdd(a.meters + (size_t)s * a.row_u64 + 2 * C, (unsigned long long)nfr);
        if (f0 < f1)
            run_item_generic<GM, METER>(a, C, s, f0, f1, lane);
    }
    tick_end(a);
}

}  // namespace cmgpu
