// pcie_probe.cu -- what the host link of this box gives: pinned H2D, D2H and both at once, by
// transfer size. The end-to-end figure of bench.py (PCM up, transformed PCM down, every sample
// crosses the link twice) is bounded by the bidirectional number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_probe pcie_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main()
{
    const size_t total = 1ull << 30;
    char *h_in, *h_out, *d_in, *d_out;
    CK(cudaMallocHost(&h_in, total));
    CK(cudaMallocHost(&h_out, total));
    CK(cudaMalloc(&d_in, total));
    CK(cudaMalloc(&d_out, total));
    for (size_t i = 0; i < total; i += 4096) h_in[i] = (char)i;
    cudaStream_t up, down;
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const size_t chunks[] = {1u << 20, 8u << 20, 48u << 20, 192u << 20, 1u << 30};
    printf("%-12s %12s %12s %22s\n", "chunk", "H2D GB/s", "D2H GB/s", "both at once (each way)");
    for (size_t chunk : chunks) {
        float ms_up = 0, ms_down = 0, ms_both = 0;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0, up));
            for (size_t o = 0; o < total; o += chunk) CK(cudaMemcpyAsync(d_in + o, h_in + o, chunk < total - o ? chunk : total - o, cudaMemcpyHostToDevice, up));
            CK(cudaEventRecord(e1, up)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_up, e0, e1));
            CK(cudaEventRecord(e0, down));
            for (size_t o = 0; o < total; o += chunk) CK(cudaMemcpyAsync(h_out + o, d_out + o, chunk < total - o ? chunk : total - o, cudaMemcpyDeviceToHost, down));
            CK(cudaEventRecord(e1, down)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_down, e0, e1));
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, up));
            CK(cudaStreamWaitEvent(down, e0, 0));
            for (size_t o = 0; o < total; o += chunk) {
                const size_t n = chunk < total - o ? chunk : total - o;
                CK(cudaMemcpyAsync(d_in + o, h_in + o, n, cudaMemcpyHostToDevice, up));
                CK(cudaMemcpyAsync(h_out + o, d_out + o, n, cudaMemcpyDeviceToHost, down));
            }
            CK(cudaEventRecord(e2, down));
            CK(cudaStreamWaitEvent(up, e2, 0));
            CK(cudaEventRecord(e1, up)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_both, e0, e1));
        }
        printf("%-12zu %12.1f %12.1f %22.1f\n", chunk, total / ms_up / 1e6, total / ms_down / 1e6, total / ms_both / 1e6);
    }
    return 0;
}
