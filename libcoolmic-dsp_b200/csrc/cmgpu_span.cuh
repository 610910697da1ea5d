// cmgpu_span.cuh -- span_tick: the small-buffer regime (BASELINE config 3), stream-major.
//
// A span is ONE launch over n_ticks consecutive ring slots (cmgpu_process_cycle). fused_tick walks a
// span as independent work items (tick, stream): for 640-byte stream-blocks that is 5 vectors of
// arithmetic per lane followed by a full meter epilogue (shuffle tree, a dependent re-read, three
// atomics) -- a third of all instructions. Here an 8-lane group owns one STREAM for the whole span:
//   * the stream's recipe is loaded once, not once per tick;
//   * the meter partials (one merged peak key and one 64-bit power sum per channel of the lane) stay in
//     registers across the ticks; ONE epilogue per stream and span publishes them -- the reference's
//     meter window is many reads long anyway (vumeter.c:170-177 accumulates until result());
//   * a tick's vectors are staged through shared memory with cp.async, a whole tick ahead of the one
//     being worked on (below), so every lane always has a full tick of loads in flight.
// Per tick the arithmetic is fused_tick's (do_vector). Peak order: the in-loop key counts steps through
// the whole span (tick * 8 + vector of the lane), the epilogue turns the winner into the 64-bit
// position key of its tick, so "first occurrence" holds across ticks, launches and GPUs as before.
// Used when the span has enough streams to fill the machine with 8-lane groups (the host decides).
#pragma once

#include "cmgpu_kernels.cuh"

namespace cmgpu {

constexpr uint32_t kSpanMaxTicks = 1024;      // (tick * 8 + step) * 8 + lane must fit 16 bits

// Shared-memory staging of a tick's vectors (cp.async, 16 bytes per copy, L2 only): every lane copies
// exactly the vectors it will work on into slots of its own, so no lane ever reads another's data and
// the only synchronisation is the lane's own cp.async.wait_group. Two stages: while tick t is worked
// on from stage t & 1, ALL of tick t + 1 is in flight into the other one. Measured on config 3
// (profiles/r2_span_tick_cfg3_*): with the vectors held in two register half-buffers a lane had 1.6
// vectors in flight on average and the launch ran at the rate Little's law gives for that (4.4 TB/s,
// half the warp cycles on long_scoreboard); staged, a whole tick per lane is in flight all the time.
template <int C, int GM, bool METER, bool NC>
__global__ void __launch_bounds__(256, Tune<C, 8>::kMinCtas) span_tick(const __grid_constant__ TickArgs a, const uint32_t vmax)
{
    constexpr int G = 8;
    constexpr int P = Shape<C>::kPerLane;
    constexpr int S = (C <= 8) ? 8 / P : 1;               // frames per vector
    constexpr size_t kStep = (size_t)G * 16;
    static_assert(C <= 8, "16-channel frames span two lanes: they use the 32-lane kernels");
    extern __shared__ uint4 span_stage[];                 // [2][vmax][256]: slot (stage, u) of thread tid
    launch_begin();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gl = threadIdx.x & (G - 1);
    const unsigned gmask = ((1u << G) - 1u) << (lane & ~(uint32_t)(G - 1));
    const uint32_t n_groups = gridDim.x * (256 / G);
    const ptrdiff_t out_delta = a.out - a.in;
    const uint32_t n_ticks = a.n_ticks ? a.n_ticks : 1u;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(span_stage) + threadIdx.x * 16u;
    const uint32_t stage_bytes = vmax * 256u * 16u;

    for (uint32_t s = blockIdx.x * (256 / G) + threadIdx.x / G; s < a.n_streams; s += n_groups) {
        Recipe rc[P];
        load_recipes<C, GM>(a, s, gl, rc);
        uint32_t kmax[8];
        uint64_t pacc[P];
#pragma unroll
        for (int k = 0; k < 8; k++)
            kmax[k] = 0;
#pragma unroll
        for (int c = 0; c < P; c++)
            pacc[c] = 0;
        uint64_t frames_total = 0;

        // what a lane has to do in tick t: n full vectors (its vectors gl, gl + 8, ...), and possibly
        // the one vector that straddles the end of the valid frames
        auto shape = [&](uint32_t nfr, uint32_t &n_i, uint32_t &tail_vec, int &tail_valid) {
            const uint32_t valid_bytes = nfr * (uint32_t)(2 * C);
            const uint32_t vfull = valid_bytes >> 4;
            n_i = gl < vfull ? (vfull - gl + (G - 1)) / G : 0u;
            tail_valid = 0;
            tail_vec = vfull;
            if ((vfull << 4) < valid_bytes && (vfull % G) == gl)
                tail_valid = (int)((valid_bytes - (vfull << 4)) >> 1);
        };
        auto frames_of = [&](uint32_t t) -> uint32_t {
            return a.frames ? min(__ldg(a.frames + (size_t)t * a.frames_stride + s), a.block_frames) : a.block_frames;
        };
        auto request = [&](const uint8_t *src_t, uint32_t n, uint32_t stage) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#ifdef CMGPU_BOUNDS_CHECK
                if ((uint32_t)u < n && !dbg_ok(src_t + (size_t)u * kStep))
                    continue;
#endif
                if ((uint32_t)u < n)
                    cp_async16(smem0 + stage * stage_bytes + (uint32_t)u * 4096u, src_t + (size_t)u * kStep);
            }
            cp_async_commit();                             // (an empty group when n == 0: the count stays uniform)
        };

        const uint8_t *src = a.in + (size_t)s * a.stride_bytes + (size_t)gl * 16;      // tick 0
        uint32_t nfr = frames_of(0);
        uint32_t nfr_next = n_ticks > 1 ? frames_of(1) : 0u;
        uint32_t n_i, tail_vec;
        int tail_valid;
        shape(nfr, n_i, tail_vec, tail_valid);
        request(src, n_i, 0u);
        for (uint32_t t = 0; t < n_ticks; t++) {
            uint8_t *const dstp = const_cast<uint8_t *>(src) + out_delta;
            const uint32_t radd0 = 0xffffu - t * 8u;
            const uint32_t stage = t & 1u;
            // all of the next tick goes out before this one is worked on
            const uint8_t *src_n = src + a.slot_bytes;
            uint32_t n_n = 0, tail_vec_n = 0;
            int tail_valid_n = 0;
            if (t + 1 < n_ticks)
                shape(nfr_next, n_n, tail_vec_n, tail_valid_n);
            request(src_n, n_n, stage ^ 1u);
            cp_async_wait<1>();                            // this tick's copies have landed (the next tick's may be in flight)
            const uint4 *mine = span_stage + (size_t)stage * vmax * 256u + threadIdx.x;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if ((uint32_t)u < n_i) {
                    const uint4 w = mine[(size_t)u * 256u];
                    const uint4 o = do_vector<C, GM, METER, false, true, false, true>(w, rc, radd0 - (uint32_t)u, kmax, pacc, 8);
                    if (a.store)
                        st_stream(dstp + (size_t)u * kStep, o);
                }
            }
            if (tail_valid) {
                // the one vector that straddles the end of the valid frames (ragged ticks only)
                const size_t off = (size_t)t * a.slot_bytes + (size_t)s * a.stride_bytes + (size_t)tail_vec * 16;
                const uint4 w = ld_stream(a.in + off, NC);
                const uint4 o = do_vector<C, GM, METER, true, true, false, true>(w, rc, radd0 - tail_vec / G, kmax, pacc, tail_valid);
                if (a.store)
                    st_stream(a.out + off, o);
            }
            if (gl == 0)
                frames_total += nfr;
            src = src_n;
            nfr = nfr_next;
            n_i = n_n;
            tail_vec = tail_vec_n;
            tail_valid = tail_valid_n;
            nfr_next = t + 2 < n_ticks ? frames_of(t + 2) : 0u;
        }
        cp_async_wait<0>();
        if (!METER)
            continue;

        // ---- one epilogue per stream and span ----
        uint32_t k32[P];
#pragma unroll
        for (int c = 0; c < P; c++) {
            const uint32_t best = kmax[c];
            const uint32_t idx = (0xffffu - (best & 0xffffu)) * (uint32_t)G + gl;     // (tick * 8 + step) * 8 + lane
            k32[c] = (best >> 16) ? ((best & 0xffff0000u) | (0xffffu - idx)) : 0u;
        }
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
#pragma unroll
            for (int c = 0; c < P; c++) {
                k32[c] = max(k32[c], __shfl_xor_sync(gmask, k32[c], off));
                pacc[c] += shfl_xor64(gmask, pacc[c], off);
            }
        }
        uint32_t key32 = 0;
        uint64_t pw = 0;
#pragma unroll
        for (int c = 0; c < P; c++) {
            if ((int)gl == c) {
                key32 = k32[c];
                pw = pacc[c];
            }
        }
        __syncwarp(gmask);       // make the span's stores visible to the lane that re-reads a sample
        if ((int)gl < C) {
            const int ch = (int)gl;
            unsigned long long *row = a.meters + (size_t)s * a.row_u64;
            if (key32) {
                const uint32_t mag = key32 >> 16;
                const uint32_t idx = 0xffffu - (key32 & 0xffffu);
                const uint32_t t = idx >> 6;                                          // idx = tick * 64 + vector
                const uint32_t v = idx & 63u;
                const uint32_t frame = v * (uint32_t)S;
                const uint64_t pos = tick_pos_base(a.tick, a.tick_offset + t, a.pbits) + frame;
                const volatile int16_t *y = reinterpret_cast<const volatile int16_t *>(
                    a.out + (size_t)t * a.slot_bytes + (size_t)s * a.stride_bytes);
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
        if (gl == 0 && frames_total)
            atomicAdd(a.meters + (size_t)s * a.row_u64 + 2 * C, (unsigned long long)frames_total);
    }
    tick_end(a);
}

}  // namespace cmgpu
