// cmgpu_tma.cuh -- the fused tick for long mono / stereo stream-blocks, staged through shared
// memory with TMA bulk copies (cp.async.bulk, SASS UBLKCP).
//
// Why it exists: a plain device copy reaches 6.10 TB/s with the batched-LDG pattern of fused_tick,
// 6.47 TB/s with 16 KiB TMA bulk tiles and 6.55 TB/s with cudaMemcpy (tools/copy_probe.cu on this
// B200); the LDG kernel already sits on the first number, so only a different way of moving the
// bytes could lift it further.
// What was measured (cfg2): in copy mode this kernel is faster than fused_tick
// (0.632-0.647 ms vs 0.650), with gain + meter it is slower (0.684 vs 0.650): by then the tick is
// bound by instruction issue, and this design adds a sign op per sample, LDS/STS and per-tile
// barrier traffic, and makes all 8 consumer warps wait on the same tile. It is therefore OPT-IN
// (CMGPU_TMA=1), kept bit-exact by the same tests, as the starting point for round 2.
//
// One CTA = 8 consumer warps + 1 producer warp (one elected thread). The producer walks the CTA's
// work items (stream, chunk) tile by tile: 16 KiB bulk load global -> shared into a 4-slot ring
// (mbarrier complete_tx), and, once the consumers have finished a slot, bulk store shared -> global
// of the transformed tile. The consumers take a slot when its barrier flips, each thread owning 4 of
// the tile's 1,024 vectors (conflict-free LDS.128 / STS.128), run the same per-sample code as
// fused_tick in place in shared memory, fence to the async proxy and arrive on the slot's
// "consumed" barrier. Meter state stays in registers across an item's tiles; at the item's end the
// warps combine through shuffles and a small shared array, and one thread per channel issues the
// atomics. Because the stores are asynchronous the winner's sign cannot be re-read from memory: it
// rides in bit 0 of the in-loop key instead (do_vector<SIGNKEY>).
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
#define CMGPU_TMA_STAGES 4      // tiles in the shared-memory ring
#endif
#ifndef CMGPU_TMA_CTAS
#define CMGPU_TMA_CTAS 3        // resident CTAs per SM the kernel is compiled for
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
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumers) : "memory"); }

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
    static_assert(C == 1 || C == 2, "tma_tick: mono and stereo");
    constexpr int P = Shape<C>::kPerLane;
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ uint64_t full[kTmaStages], consumed[kTmaStages];
    __shared__ unsigned long long red_key[2][kTmaConsumerWarps][2], red_pow[2][kTmaConsumerWarps][2];

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    const uint64_t n_items = (uint64_t)a.n_streams * a.items_per_block;

    if (tid == 0) {
        for (int s = 0; s < kTmaStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&consumed[s], kTmaConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kTmaConsumerWarps) {
        // ---------------------------------------------------------------- producer (one thread)
        if (lane == 0) {
            uint8_t *pend_dst[kTmaStages];
            uint32_t pend_bytes[kTmaStages];
            uint32_t k = 0;
            for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const TmaItem it = tma_item<C>(a, item);
                const size_t base = (size_t)it.s * a.stride_bytes;
                for (uint32_t t = 0; t < it.ntile; t++, k++) {
                    const uint32_t slot = k % kTmaStages;
                    const uint32_t tv0 = it.v0 + t * kTmaTileVecs;
                    const uint32_t bytes = min((uint32_t)kTmaTileVecs, it.v1 - tv0) * 16u;
                    if (k >= kTmaStages) {
                        // the tile that used this slot: wait for the consumers, send it home, free the slot
                        mbar_wait(&consumed[slot], (k / kTmaStages - 1) & 1u);
                        if (a.store) {
                            bulk_store(pend_dst[slot], ring + (size_t)slot * kTmaTileBytes, pend_bytes[slot]);
                            bulk_wait_read_all();
                        }
                    }
                    pend_dst[slot] = a.out + base + (size_t)tv0 * 16;
                    pend_bytes[slot] = bytes;
                    mbar_expect_tx(&full[slot], bytes);
                    bulk_load(ring + (size_t)slot * kTmaTileBytes, a.in + base + (size_t)tv0 * 16, bytes, &full[slot]);
                }
            }
            // drain: the last min(k, stages) tiles are still in their slots
            for (uint32_t j = k > kTmaStages ? k - kTmaStages : 0; j < k; j++) {
                const uint32_t slot = j % kTmaStages;
                mbar_wait(&consumed[slot], (j / kTmaStages) & 1u);
                if (a.store)
                    bulk_store(pend_dst[slot], ring + (size_t)slot * kTmaTileBytes, pend_bytes[slot]);
            }
            bulk_wait_all();
        }
    } else {
        // ---------------------------------------------------------------- consumers (8 warps)
        uint32_t kmax[8];
        uint64_t pacc[P];
#pragma unroll
        for (int q = 0; q < 8; q++)
            kmax[q] = 0;
#pragma unroll
        for (int c = 0; c < P; c++)
            pacc[c] = 0;
        uint32_t k = 0, par = 0;
        for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const TmaItem it = tma_item<C>(a, item);
            Recipe rc[P];
            load_recipes<C, GM>(a, it.s, 0, rc);
            for (uint32_t t = 0; t < it.ntile; t++, k++) {
                const uint32_t slot = k % kTmaStages;
                uint8_t *tile = ring + (size_t)slot * kTmaTileBytes;
                const uint32_t tv0 = it.v0 + t * kTmaTileVecs;
                const uint32_t nv = min((uint32_t)kTmaTileVecs, it.v1 - tv0);
                // the block's very last vector may straddle the end of the valid frames
                const bool partial = ((tv0 + nv) << 4) > it.valid_bytes;
                const uint32_t nfull = partial ? nv - 1 : nv;
                mbar_wait(&full[slot], (k / kTmaStages) & 1u);
#pragma unroll
                for (int u = 0; u < kTmaVecsPerThread; u++) {
                    const uint32_t vi = (uint32_t)u * kTmaConsumers + tid;
                    if (vi < nfull) {
                        const uint32_t radd = (0x7fffu - (t * kTmaVecsPerThread + (uint32_t)u)) << 1;
                        uint4 *p = reinterpret_cast<uint4 *>(tile + (size_t)vi * 16);
                        const uint4 o = do_vector<C, GM, METER, false, true, true>(*p, rc, radd, kmax, pacc, 8);
                        if (a.store)
                            *p = o;
                    }
                }
                if (partial && ((nv - 1) % kTmaConsumers) == tid) {
                    const uint32_t vi = nv - 1, u = vi / kTmaConsumers;
                    const uint32_t radd = (0x7fffu - (t * kTmaVecsPerThread + u)) << 1;
                    const int nvalid = (int)((it.valid_bytes - ((tv0 + vi) << 4)) >> 1);
                    uint4 *p = reinterpret_cast<uint4 *>(tile + (size_t)vi * 16);
                    const uint4 o = do_vector<C, GM, METER, true, true, true>(*p, rc, radd, kmax, pacc, nvalid);
                    if (a.store)
                        *p = o;
                }
                fence_async_proxy();          // generic-proxy writes -> visible to the bulk store
                mbar_arrive(&consumed[slot]);
            }
            if (!METER)
                continue;
            // ---- item epilogue: lane -> warp (shuffles) -> CTA (shared) -> one atomic per channel
            const uint64_t pos_base = tick_begin(a);
            uint64_t kc[P];
#pragma unroll
            for (int c = 0; c < P; c++) {
                uint32_t best = kmax[c];
                uint32_t sub = 0;
#pragma unroll
                for (int j = 1; j < 8 / P; j++) {
                    // compare without the sign bit: an equal key in a later slot must not win
                    if ((kmax[c + j * P] >> 1) > (best >> 1)) {
                        best = kmax[c + j * P];
                        sub = (uint32_t)j;
                    }
                }
                const uint32_t mag = best >> 16;
                const uint32_t step = 0x7fffu - ((best >> 1) & 0x7fffu);
                const uint32_t v = it.v0 + (step / kTmaVecsPerThread) * kTmaTileVecs + (step % kTmaVecsPerThread) * kTmaConsumers + tid;
                const uint32_t frame = v * (uint32_t)Shape<C>::kFramesPerVec8 + sub;
                kc[c] = mag ? (make_key(mag, pos_base + frame) | (uint64_t)(best & 1u)) : 0ull;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                for (int c = 0; c < P; c++) {
                    kc[c] = max(kc[c], shfl_xor64(0xffffffffu, kc[c], off));
                    pacc[c] += shfl_xor64(0xffffffffu, pacc[c], off);
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < P; c++) {
                    red_key[par][warp][c] = kc[c];
                    red_pow[par][warp][c] = pacc[c];
                }
            }
            consumer_barrier();
            if (tid < (uint32_t)C) {
                unsigned long long key = 0, pw = 0;
                for (int w = 0; w < kTmaConsumerWarps; w++) {
                    key = max(key, red_key[par][w][tid]);
                    pw += red_pow[par][w][tid];
                }
                unsigned long long *row = a.meters + (size_t)it.s * a.row_u64;
                if (key)
                    atomicMax(row + tid, key);
                if (pw)
                    atomicAdd(row + C + tid, pw);
                if (tid == 0 && it.chunk == 0 && it.nfr)
                    atomicAdd(row + 2 * C, (unsigned long long)it.nfr);
            }
            par ^= 1;                 // the next item writes the other half of the scratch
#pragma unroll
            for (int q = 0; q < 8; q++)
                kmax[q] = 0;
#pragma unroll
            for (int c = 0; c < P; c++)
                pacc[c] = 0;
        }
    }
    tick_end(a);
}

}  // namespace cmgpu
