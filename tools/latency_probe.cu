// tools/latency_probe.cu -- where the microseconds of ONE small tick go (BASELINE config 3: 16,384 mono
// streams x 320 frames). Launch-to-complete of:
//   (a) an empty kernel, 1 CTA and a GPU-filling grid, completion seen by polling cudaStreamQuery;
//   (b) the same with completion seen through a flag the kernel's last CTA writes to mapped host memory;
//   (c) cmgpu_process + cmgpu_sync (the product's path), with the time cmgpu_process itself takes;
//   (d) the tick between CUDA events.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -I include -o tools/latency_probe tools/latency_probe.cu
//        -L libcoolmic-dsp_b200/lib -lcoolmic_b200 -Xlinker -rpath=$PWD/libcoolmic-dsp_b200/lib
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "cmgpu.h"

static long long now_ns()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}

__global__ void empty_kernel() {}
struct BigArgs { unsigned long long w[30]; };      // 240 bytes, the size of the tick kernels' argument block
__global__ void big_arg_kernel(const __grid_constant__ BigArgs a, unsigned *sink)
{
    if (a.w[29] == 0x1234567ull)
        *sink = 1;
}

__global__ void flag_kernel(unsigned *count, volatile unsigned *host_flag, unsigned gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(count, 1u) == gridDim.x - 1) {
            *count = 0;
            *host_flag = gen;
            __threadfence_system();
        }
    }
}

static void report(const char *what, std::vector<float> &us)
{
    std::sort(us.begin(), us.end());
    printf("%-64s median %6.2f  min %6.2f  p90 %6.2f us\n", what, us[us.size() / 2], us[0], us[us.size() * 9 / 10]);
}

int main()
{
    const unsigned reps = 400;
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    unsigned *d_count;
    cudaMalloc(&d_count, 4);
    cudaMemset(d_count, 0, 4);
    volatile unsigned *h_flag;
    cudaHostAlloc((void **)&h_flag, 4, cudaHostAllocMapped);
    *h_flag = 0;
    unsigned *d_flag;
    cudaHostGetDevicePointer((void **)&d_flag, (void *)h_flag, 0);
    std::vector<float> us(reps);

    for (unsigned grid : {1u, 592u}) {
        for (unsigned r = 0; r < reps; r++) {
            cudaStreamSynchronize(st);
            const long long t0 = now_ns();
            empty_kernel<<<grid, 256, 0, st>>>();
            while (cudaStreamQuery(st) == cudaErrorNotReady) {}
            us[r] = (now_ns() - t0) * 1e-3f;
        }
        char name[96];
        snprintf(name, sizeof name, "empty kernel, grid %u, cudaStreamQuery polling", grid);
        report(name, us);
        for (unsigned r = 0; r < reps; r++) {
            cudaStreamSynchronize(st);
            const long long t0 = now_ns();
            empty_kernel<<<grid, 256, 0, st>>>();
            cudaStreamSynchronize(st);
            us[r] = (now_ns() - t0) * 1e-3f;
        }
        snprintf(name, sizeof name, "empty kernel, grid %u, cudaStreamSynchronize", grid);
        report(name, us);
        for (unsigned r = 0; r < reps; r++) {
            cudaStreamSynchronize(st);
            const unsigned gen = r + 1;
            const long long t0 = now_ns();
            flag_kernel<<<grid, 256, 0, st>>>(d_count, d_flag, gen);
            while (*h_flag != gen) {}
            us[r] = (now_ns() - t0) * 1e-3f;
        }
        snprintf(name, sizeof name, "flag kernel, grid %u, mapped host flag polling", grid);
        report(name, us);
    }
    {
        for (unsigned r = 0; r < reps; r++) {
            cudaStreamSynchronize(st);
            const long long t0 = now_ns();
            empty_kernel<<<1, 32, 0, st>>>();
            us[r] = (now_ns() - t0) * 1e-3f;
        }
        report("empty kernel: the launch call alone (host)", us);
    }
    {
        BigArgs big = {};
        for (unsigned r = 0; r < reps; r++) {
            cudaStreamSynchronize(st);
            const long long t0 = now_ns();
            big_arg_kernel<<<1, 32, 0, st>>>(big, d_count);
            us[r] = (now_ns() - t0) * 1e-3f;
        }
        report("240-byte argument block, <<<>>>: the launch call alone (host)", us);
        for (int pdl = 0; pdl < 2; pdl++) {
            for (unsigned r = 0; r < reps; r++) {
                cudaStreamSynchronize(st);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(512);
                cfg.blockDim = dim3(256);
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr;
                cfg.numAttrs = pdl;
                const long long t0 = now_ns();
                cudaLaunchKernelEx(&cfg, big_arg_kernel, big, d_count);
                us[r] = (now_ns() - t0) * 1e-3f;
            }
            report(pdl ? "240-byte block, cudaLaunchKernelEx + PDL attribute, grid 512: call" : "240-byte block, cudaLaunchKernelEx, grid 512: call alone", us);
        }
    }
    {
        // completion through a stream memory operation: the front end writes the flag when the stream gets there
        typedef CUresult (*write32_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
        write32_t write32 = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", (void **)&write32, cudaEnableDefault, &qr) != cudaSuccess || !write32) {
            printf("cuStreamWriteValue32 not available\n");
        } else {
            for (unsigned grid : {1u, 592u}) {
                unsigned bad = 0;
                for (unsigned r = 0; r < reps; r++) {
                    cudaStreamSynchronize(st);
                    const unsigned gen = 1000000u + r;
                    const long long t0 = now_ns();
                    empty_kernel<<<grid, 256, 0, st>>>();
                    if (write32((CUstream)st, (CUdeviceptr)d_flag, gen, 0) != CUDA_SUCCESS)
                        bad++;
                    while (*h_flag != gen && !bad) {}
                    us[r] = (now_ns() - t0) * 1e-3f;
                }
                char name[96];
                snprintf(name, sizeof name, "empty kernel, grid %u, cuStreamWriteValue32 + host polling%s", grid, bad ? " (FAILED)" : "");
                report(name, us);
            }
            for (unsigned r = 0; r < reps; r++) {
                cudaStreamSynchronize(st);
                const long long t0 = now_ns();
                write32((CUstream)st, (CUdeviceptr)d_flag, 5u, 0);
                us[r] = (now_ns() - t0) * 1e-3f;
            }
            report("cuStreamWriteValue32: the call alone (host)", us);
        }
        for (unsigned r = 0; r < reps; r++) {
            const long long t0 = now_ns();
            cudaStreamQuery(st);
            us[r] = (now_ns() - t0) * 1e-3f;
        }
        report("cudaStreamQuery on an idle stream: the call alone", us);
        for (unsigned r = 0; r < reps; r++) {
            const long long t0 = now_ns();
            cudaSetDevice(0);
            us[r] = (now_ns() - t0) * 1e-3f;
        }
        report("cudaSetDevice (unchanged): the call alone", us);
    }

    // the product's single tick, config 3
    const unsigned streams = 16384, frames = 320;
    cmgpu_ctx_t *c = cmgpu_ctx_create(0, 1, streams, 1, frames, 0);
    if (!c) {
        printf("cmgpu_ctx_create: %s\n", cmgpu_last_error());
        return 1;
    }
    std::vector<uint16_t> scale(streams), gain(streams);
    for (unsigned s = 0; s < streams; s++) {
        scale[s] = (uint16_t)(1000 + s % 9000);
        gain[s] = (uint16_t)(scale[s] * 3 / 4 + 37 * (s % 64));
    }
    cmgpu_set_gain_table(c, 0, streams, scale.data(), gain.data());
    int16_t *h = (int16_t *)cmgpu_host_slot(c, 0);
    for (size_t i = 0; i < (size_t)streams * frames; i++)
        h[i] = (int16_t)((i * 2654435761u) >> 16);
    cmgpu_submit(c, 0, nullptr);
    cmgpu_process(c, 0, CMGPU_FUSED);
    cmgpu_sync(c);
    std::vector<float> call(reps);
    for (unsigned r = 0; r < reps; r++) {
        cmgpu_sync(c);
        const long long t0 = now_ns();
        cmgpu_process(c, 0, CMGPU_FUSED);
        const long long t1 = now_ns();
        cmgpu_sync(c);
        us[r] = (now_ns() - t0) * 1e-3f;
        call[r] = (t1 - t0) * 1e-3f;
    }
    report("cmgpu_process + cmgpu_sync (config 3 single tick)", us);
    report("  of which the cmgpu_process call", call);
    printf("kernel: %s\n", cmgpu_kernel_name(c));
    float med = 0, mn = 0;
    cmgpu_time_single_tick(c, 0, CMGPU_FUSED, reps, &med, &mn);
    printf("cmgpu_time_single_tick: median %.2f min %.2f us\n", med, mn);
    float ms = 0;
    cmgpu_time_process(c, 0, 1, 1, CMGPU_FUSED, &ms);
    printf("one tick between CUDA events: %.2f us\n", ms * 1e3f);
    cmgpu_time_process(c, 0, 1, 100, CMGPU_FUSED, &ms);
    printf("100 ticks back to back between CUDA events: %.2f us per tick\n", ms * 10.f);
    cmgpu_ctx_destroy(c);
    return 0;
}
