// tools/copy_probe.cu -- how fast can a plain device-to-device copy go on this B200, by access pattern?
// Decides whether TMA / shared-memory staging could lift the fused kernel above the ceiling of its
// current batched-LDG pattern (north_star: "TMA or shared-memory staging only where ncu shows it
// lifts achieved HBM GB/s"). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o copy_probe copy_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda/barrier>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// (1) our pattern: each warp walks a contiguous 32 KiB chunk, 4 x 128-bit loads per lane per batch,
//     double buffered (what fused_tick does in --mode copy)
__global__ void __launch_bounds__(256, 3) copy_warp_chunks(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t nvec, uint32_t per_item)
{
    const uint32_t lane = threadIdx.x & 31;
    const size_t n_items = (nvec + per_item - 1) / per_item;
    const size_t stride = (size_t)gridDim.x * 8;
    for (size_t item = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); item < n_items; item += stride) {
        const size_t v0 = item * per_item;
        const uint4 *src = in + v0 + lane;
        uint4 *dst = out + v0 + lane;
        const uint32_t n_i = per_item / 32;
        uint4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) a[u] = __ldcs(src + (size_t)u * 32);
        for (uint32_t i = 0; i < n_i; i += 8) {
#pragma unroll
            for (int u = 0; u < 4; u++) b[u] = __ldcs(src + (size_t)(i + 4 + u) * 32);
#pragma unroll
            for (int u = 0; u < 4; u++) __stcs(dst + (size_t)(i + u) * 32, a[u]);
            if (i + 8 < n_i) {
#pragma unroll
                for (int u = 0; u < 4; u++) a[u] = __ldcs(src + (size_t)(i + 8 + u) * 32);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) __stcs(dst + (size_t)(i + 4 + u) * 32, b[u]);
        }
    }
}


// (1b) the same loads, but the stores go through a per-warp shared-memory staging tile and leave as
//      ONE bulk store (cp.async.bulk shared -> global) of BATCHES x 2 KiB per warp: does the write side
//      of the pattern get cheaper when it arrives in bulk?
template <int BATCHES>
__global__ void __launch_bounds__(256, 3) copy_warp_chunks_bulk_store(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t nvec, uint32_t per_item)
{
    extern __shared__ __align__(128) uint8_t stage[];       // [warp][2 tiles][BATCHES * 2048]
    constexpr uint32_t kTile = BATCHES * 2048;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *mine = stage + (size_t)warp * 2 * kTile;
    const size_t n_items = (nvec + per_item - 1) / per_item;
    const size_t stride = (size_t)gridDim.x * 8;
    uint32_t t = 0;
    for (size_t item = (size_t)blockIdx.x * 8 + warp; item < n_items; item += stride) {
        const size_t v0 = item * per_item;
        const uint4 *src = in + v0 + lane;
        const uint32_t n_b = per_item / 128;               // batches of 4 vectors per lane
        uint4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) a[u] = __ldg(src + (size_t)u * 32);
        for (uint32_t bi = 0; bi < n_b; bi++) {
            if (bi + 1 < n_b) {
#pragma unroll
                for (int u = 0; u < 4; u++) b[u] = __ldg(src + (size_t)((bi + 1) * 4 + u) * 32);
            }
            const uint32_t sub = bi % BATCHES;
            uint8_t *tile = mine + (size_t)(t & 1u) * kTile;
            if (sub == 0) {
                // the tile is free again once the bulk store issued two tiles ago has read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                *reinterpret_cast<uint4 *>(tile + (size_t)sub * 2048 + (size_t)u * 512 + lane * 16) = a[u];
            if (sub == BATCHES - 1 || bi + 1 == n_b) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const uint32_t bytes = (sub + 1) * 2048;
                    uint8_t *dst = reinterpret_cast<uint8_t *>(out + v0) + (size_t)(bi - sub) * 2048;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                                 "r"((uint32_t)__cvta_generic_to_shared(tile)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                t++;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) a[u] = b[u];
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// (2) classic grid-stride elementwise copy, 4 x 128-bit per thread per step, whole grid sweeps memory in lockstep
__global__ void __launch_bounds__(256) copy_grid_stride(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t nvec)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (size_t base = tid; base + 3 * nthreads < nvec; base += 4 * nthreads) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldcs(in + base + u * nthreads);
#pragma unroll
        for (int u = 0; u < 4; u++) __stcs(out + base + u * nthreads, v[u]);
    }
}

// (3) TMA 1-D bulk copies through shared memory: one elected thread per CTA moves 16 KiB tiles
//     global -> shared -> global with cp.async.bulk, 4 stages
__global__ void __launch_bounds__(128) copy_tma_bulk(const uint8_t *in, uint8_t *out, size_t bytes, uint32_t tile)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int STAGES = 4;
    __shared__ uint64_t full[STAGES];
    const size_t n_tiles = bytes / tile;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full[s]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x != 0)
        return;
    uint32_t phase[STAGES] = {0, 0, 0, 0};
    size_t t = blockIdx.x;
    size_t issued[STAGES];
    int n_inflight = 0, head = 0, tail = 0;
    auto issue_load = [&](int s, size_t tileidx) {
        uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full[s]);
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + (size_t)s * tile);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(in + tileidx * tile), "r"(tile), "r"(bar) : "memory");
        issued[s] = tileidx;
    };
    while (n_inflight < STAGES && t < n_tiles) {
        // the stage's previous store must have finished reading shared memory
        issue_load(head, t);
        head = (head + 1) % STAGES;
        n_inflight++;
        t += gridDim.x;
    }
    while (n_inflight) {
        const int s = tail;
        uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full[s]);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar), "r"(phase[s]) : "memory");
        phase[s] ^= 1;
        uint32_t src = (uint32_t)__cvta_generic_to_shared(smem + (size_t)s * tile);
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + issued[s] * tile), "r"(src), "r"(tile) : "memory");
        asm volatile("cp.async.bulk.commit_group;");
        tail = (tail + 1) % STAGES;
        n_inflight--;
        if (t < n_tiles) {
            // wait until the store that used this stage has read it, then refill the stage
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue_load(s, t);
            head = (s + 1) % STAGES;
            n_inflight++;
            t += gridDim.x;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv)
{
    // size in GiB of EACH of the two buffers (default 2): the footprint matters -- see DESIGN.md 4.1
    const double gib = argc > 1 ? atof(argv[1]) : 2.0;
    const size_t bytes = ((size_t)(gib * (double)((size_t)1 << 30))) & ~(size_t)0xffff;
    printf("buffers: 2 x %.2f GiB\n", (double)bytes / (double)((size_t)1 << 30));
    uint8_t *in, *out;
    CK(cudaMalloc(&in, bytes));
    CK(cudaMalloc(&out, bytes));
    CK(cudaMemset(in, 1, bytes));
    CK(cudaMemset(out, 0, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const size_t nvec = bytes / 16;
    auto report = [&](const char *name, float ms) {
        printf("%-44s %8.3f ms  %8.1f GB/s (read+write)\n", name, ms, 2.0 * bytes / (ms * 1e-3) / 1e9);
    };
    float ms, best;
    // cudaMemcpyAsync D2D as the reference point
    best = 1e9;
    for (int r = 0; r < 6; r++) {
        cudaEventRecord(e0);
        cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    report("cudaMemcpyAsync D2D", best);
    {
        // the same copy back to back for >= 2 s: what the part sustains once it runs into its power cap
        // (the yardstick for bench.py's roofline.sustained)
        const int reps = (int)(2200.0 / best) + 1;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++)
            cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        char name[64];
        snprintf(name, sizeof(name), "cudaMemcpyAsync D2D, %d back to back (%.1f s)", reps, ms * 1e-3);
        report(name, ms / reps);
    }
    for (uint32_t per : {1024u, 2048u, 4096u}) {
        best = 1e9;
        for (int r = 0; r < 6; r++) {
            cudaEventRecord(e0);
            copy_warp_chunks<<<sms * 3, 256>>>((const uint4 *)in, (uint4 *)out, nvec, per);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        char name[64];
        snprintf(name, sizeof(name), "warp chunks, %u vectors per item", per);
        report(name, best);
    }

    {
        auto run_bulk = [&](auto kern, int batches, const char *label) {
            const int smem = 8 * 2 * batches * 2048;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
            float b2 = 1e9, m2;
            for (int r = 0; r < 6; r++) {
                cudaEventRecord(e0);
                kern<<<sms * occ, 256, smem>>>((const uint4 *)in, (uint4 *)out, nvec, 2048u);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&m2, e0, e1);
                if (r && m2 < b2) b2 = m2;
            }
            char name[96];
            snprintf(name, sizeof(name), "warp chunks, LDG + %s bulk stores, %d CTA/SM", label, occ);
            report(name, b2);
        };
        run_bulk(copy_warp_chunks_bulk_store<1>, 1, "2 KiB");
        run_bulk(copy_warp_chunks_bulk_store<2>, 2, "4 KiB");
        run_bulk(copy_warp_chunks_bulk_store<4>, 4, "8 KiB");
    }
    for (int mult : {4, 8, 16}) {
        best = 1e9;
        for (int r = 0; r < 6; r++) {
            cudaEventRecord(e0);
            copy_grid_stride<<<sms * mult, 256>>>((const uint4 *)in, (uint4 *)out, nvec);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        char name[64];
        snprintf(name, sizeof(name), "grid-stride, %d CTAs per SM", mult);
        report(name, best);
    }
    for (uint32_t tile : {8192u, 16384u, 32768u}) {
        const int smem = 4 * tile;
        cudaFuncSetAttribute(copy_tma_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int mult : {1, 2}) {
            if ((size_t)smem * mult > 200 * 1024) continue;
            best = 1e9;
            for (int r = 0; r < 6; r++) {
                cudaEventRecord(e0);
                copy_tma_bulk<<<sms * mult, 128, smem>>>(in, out, bytes, tile);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
                if (r && ms < best) best = ms;
            }
            char name[64];
            snprintf(name, sizeof(name), "TMA bulk, %u B tiles x 4 stages, %d CTA/SM", tile, mult);
            report(name, best);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    // sanity: the last copy really copied
    uint8_t h[4];
    cudaMemcpy(h, out + bytes - 4, 4, cudaMemcpyDeviceToHost);
    printf("tail bytes %d %d %d %d\n", h[0], h[1], h[2], h[3]);
    return e != cudaSuccess;
}
