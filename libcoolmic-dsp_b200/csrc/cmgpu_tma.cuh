// cmgpu_tma.cuh -- the fused tick with its loads staged through shared memory by TMA bulk copies
// (cp.async.bulk global -> shared, SASS UBLKCP) and its stores written straight from registers.
//
// Why it exists. fused_tick keeps its loads in flight in registers: one batch of 4 vectors per lane,
// 48 KiB per SM at 24 warps. ncu shows what that costs (profiles/): 38 % of warp-cycles wait on the
// first use of a batch (long_scoreboard) and the tick sits at the copy rate of that access pattern;
// making the per-sample code SHORTER made the tick slower, not faster -- the bytes in flight, not the
// instructions, set the pace. Registers cannot hold more; shared memory can: here a ring of
// CMGPU_TMA_STAGES x 16 KiB tiles per CTA (192 KiB per SM) is kept full by one producer thread, and
// the consumer warps never wait on global memory at all.
//
// One CTA = NW consumer warps + 1 producer warp (one elected thread).
//   producer   walks the CTA's work items (stream, chunk of per_item vectors) tile by tile: waits for
//              the slot's `empty` barrier (every consumer warp has read the previous tile out of it),
//              arms `full` with the tile's byte count and issues ONE bulk load.
//   consumers  wait for `full`, pull their CMGPU_TMA_VPT vectors of the tile into registers
//              (conflict-free LDS.128: thread t owns vectors t, t+256, ...), release the slot at once
//              (one arrive per warp), then run the same per-sample code as fused_tick
//              (do_vector) and write the result with coalesced 128-bit streaming stores.
// Meter state stays in registers across the tiles of an item; at its end every warp publishes its
// own partial result with the shuffle tree + atomics of fused_tick (item_publish): position keys
// make the merge exact in any order. Nothing is written back to shared memory, so there is no
// proxy fence and no bulk store to wait for. The next item's recipes are fetched one item ahead.
//
// The first version of this file (round 1, in-place in shared memory + bulk stores, 4 x 16 KiB) is in
// the history: 0.684 ms fused on cfg2 against 0.650 for fused_tick; DESIGN.md 4.6 has both.
#pragma once

#include "cmgpu_kernels.cuh"

namespace cmgpu {

#ifndef CMGPU_TMA_NW
#define CMGPU_TMA_NW 8          // consumer warps per CTA
#endif
#ifndef CMGPU_TMA_VPT
#define CMGPU_TMA_VPT 4         // vectors of a tile each consumer thread owns
#endif
#ifndef CMGPU_TMA_STAGES
#define CMGPU_TMA_STAGES 6      // tiles in the shared-memory ring
#endif
#ifndef CMGPU_TMA_CTAS
#define CMGPU_TMA_CTAS 2        // resident CTAs per SM the kernel is compiled for
#endif
constexpr int kTmaStages = CMGPU_TMA_STAGES;
constexpr int kTmaConsumerWarps = CMGPU_TMA_NW;
constexpr int kTmaConsumers = kTmaConsumerWarps * 32;
constexpr int kTmaThreads = kTmaConsumers + 32;
constexpr int kTmaVecsPerThread = CMGPU_TMA_VPT;
constexpr int kTmaTileVecs = kTmaConsumers * kTmaVecsPerThread;
constexpr int kTmaTileBytes = kTmaTileVecs * 16;
constexpr int kTmaSmemBytes = kTmaStages * kTmaTileBytes;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 lds128(const void *p)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
    return v;
}

// The geometry of one work item, computed identically by producer and consumers.
struct TmaItem {
    uint32_t s, chunk, nfr, valid_bytes, v0, v1, ntile;
};
template <int C>
__device__ __forceinline__ TmaItem tma_item(const TickArgs &a, uint64_t item)
{
    TmaItem it;
    const uint32_t item32 = (uint32_t)item;
    it.s = a.items_per_block == 1 ? item32 : item32 / a.items_per_block;
    it.chunk = item32 - it.s * a.items_per_block;
    it.nfr = a.frames ? min(__ldg(a.frames + it.s), a.block_frames) : a.block_frames;
    it.valid_bytes = it.nfr * (uint32_t)(2 * C);
    const uint32_t nvec = (it.valid_bytes + 15u) >> 4;
    it.v0 = it.chunk * a.per_item;
    it.v1 = min(it.v0 + a.per_item, nvec);
    it.ntile = it.v0 < it.v1 ? (it.v1 - it.v0 + kTmaTileVecs - 1) / kTmaTileVecs : 0;
    return it;
}

template <int C, int GM, bool METER>
__global__ void __launch_bounds__(kTmaThreads, CMGPU_TMA_CTAS) tma_tick(const __grid_constant__ TickArgs a)
{
    constexpr int P = Shape<C>::kPerLane;
    constexpr bool kSat = C >= 4;
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ uint64_t full[kTmaStages], empty[kTmaStages];

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    const uint64_t n_items = (uint64_t)a.n_streams * a.items_per_block;
    launch_begin();

    if (tid == 0) {
        for (int s = 0; s < kTmaStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTmaConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kTmaConsumerWarps) {
        // ---------------------------------------------------------------- producer (one thread)
        if (lane == 0) {
            uint32_t k = 0;
            for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const TmaItem it = tma_item<C>(a, item);
                const uint8_t *src = a.in + (size_t)it.s * a.stride_bytes;
                for (uint32_t t = 0; t < it.ntile; t++, k++) {
                    const uint32_t slot = k % kTmaStages;
                    const uint32_t tv0 = it.v0 + t * kTmaTileVecs;
                    const uint32_t bytes = min((uint32_t)kTmaTileVecs, it.v1 - tv0) * 16u;
                    if (k >= kTmaStages)          // every consumer warp has read the slot's previous tile
                        mbar_wait(&empty[slot], (k / kTmaStages - 1) & 1u);
                    mbar_expect_tx(&full[slot], bytes);
                    bulk_load(ring + (size_t)slot * kTmaTileBytes, src + (size_t)tv0 * 16, bytes, &full[slot]);
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- consumers
        uint32_t kmax[8];
        uint64_t pacc[P];
#pragma unroll
        for (int q = 0; q < 8; q++)
            kmax[q] = 0;
#pragma unroll
        for (int c = 0; c < P; c++)
            pacc[c] = 0;
        uint32_t k = 0;
        uint64_t item = blockIdx.x;
        Recipe rc[P], rcn[P];
        TmaItem it;
        if (item < n_items) {
            it = tma_item<C>(a, item);
            load_recipes<C, GM>(a, it.s, tid, rc);
        }
        while (item < n_items) {
            // the next item's stream and recipes, one item ahead of their use
            const uint64_t item_n = item + gridDim.x;
            TmaItem itn;
            if (item_n < n_items) {
                itn = tma_item<C>(a, item_n);
                load_recipes<C, GM>(a, itn.s, tid, rcn);
            }
            const size_t base = (size_t)it.s * a.stride_bytes;
            uint8_t *const dst = a.out + base + ((size_t)it.v0 + tid) * 16;
            for (uint32_t t = 0; t < it.ntile; t++, k++) {
                const uint32_t slot = k % kTmaStages;
                const uint8_t *tile = ring + (size_t)slot * kTmaTileBytes;
                const uint32_t tv0 = it.v0 + t * kTmaTileVecs;
                const uint32_t nv = min((uint32_t)kTmaTileVecs, it.v1 - tv0);
                // the block's very last vector may straddle the end of the valid frames
                const bool partial = ((tv0 + nv) << 4) > it.valid_bytes;
                const uint32_t nfull = partial ? nv - 1 : nv;
                mbar_wait(&full[slot], (k / kTmaStages) & 1u);
                uint4 vbuf[kTmaVecsPerThread];
#pragma unroll
                for (int u = 0; u < kTmaVecsPerThread; u++) {
                    const uint32_t vi = (uint32_t)u * kTmaConsumers + tid;
                    if (vi < nv)
                        vbuf[u] = lds128(tile + (size_t)vi * 16);
                }
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&empty[slot]);         // the slot may be refilled while we compute
#pragma unroll
                for (int u = 0; u < kTmaVecsPerThread; u++) {
                    const uint32_t vi = (uint32_t)u * kTmaConsumers + tid;
                    const uint32_t step = t * kTmaVecsPerThread + (uint32_t)u;
                    if (vi < nfull) {
                        const uint4 o = do_vector<C, GM, METER, false, kSat, false, true>(vbuf[u], rc, 0xffffu - step, kmax, pacc, 8);
                        if (a.store)
                            st_stream(dst + (size_t)step * (kTmaConsumers * 16), o);
                    } else if (partial && vi == nv - 1) {
                        const int nvalid = (int)((it.valid_bytes - ((tv0 + vi) << 4)) >> 1);
                        const uint4 o = do_vector<C, GM, METER, true, kSat, false, true>(vbuf[u], rc, 0xffffu - step, kmax, pacc, nvalid);
                        if (a.store)
                            st_stream(dst + (size_t)step * (kTmaConsumers * 16), o);
                    }
                }
            }
            if (METER) {
                // ---- item epilogue, per warp: shuffle tree, one lane per channel issues the atomics
                Item pub;
                pub.src = nullptr;
                pub.base = base;
                pub.tk = 0;
                pub.s = it.s;
                pub.first = it.v0 + tid;
                pub.n_i = 0;
                pub.tail_vec = pub.tail_step = 0;
                pub.tail_valid = 0;
                pub.count_frames = (it.chunk == 0 && tid == 0) ? it.nfr : 0;
                item_publish<C, 32, true, kTmaConsumers>(a, pub, lane, 0xffffffffu, kmax, pacc);
#pragma unroll
                for (int q = 0; q < 8; q++)
                    kmax[q] = 0;
#pragma unroll
                for (int c = 0; c < P; c++)
                    pacc[c] = 0;
            }
            item = item_n;
            it = itn;
#pragma unroll
            for (int c = 0; c < P; c++)
                rc[c] = rcn[c];
        }
    }
    tick_end(a);
}

}  // namespace cmgpu
