// tools/pipe_probe.cu -- measures per-SM issue throughput of the integer instructions the fused
// kernel is built from (sm_100a). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
// Output: lane-ops per clock per SM for each instruction, from clock64() deltas.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define ITERS 512
#define CHAINS 8
#define REPS 8

enum Op { OP_IMAD, OP_IMADHI, OP_IMADWIDE, OP_LEAHI, OP_VIMNMX, OP_VIMNMX3, OP_VIADDMNMX, OP_VIMNMX16x2, OP_I2IP,
          OP_IABS, OP_PRMT, OP_LOP3, OP_SHF, OP_IADD3, OP_MIX_ALU_FMA, OP_COUNT };
static const char *names[] = {"IMAD", "IMAD.HI", "IMAD.WIDE(acc64)", "LEA.HI", "VIMNMX+IADD alternating", "VIMNMX3", "VIADDMNMX",
                              "VIMNMX.S16x2", "I2IP.S16.S32.SAT", "IABS+IADD alternating", "PRMT", "LOP3", "SHF", "IADD3",
                              "1 VIMNMX + 1 IMAD (dual pipe)"};

template <int OP>
__global__ void __launch_bounds__(256) probe(int *out, long long *cyc, int a0, int b0)
{
    int v[CHAINS];
    int u[CHAINS];
    unsigned long long w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { v[c] = a0 + threadIdx.x * (c + 1); w[c] = v[c]; u[c] = v[c] ^ 7; }
    int b = b0 + (int)threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int rc = 0; rc < CHAINS * REPS; rc++) {
            const int c = rc % CHAINS;
            if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(b), "r"(a0));
            if (OP == OP_IMADHI) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(b), "r"(a0));
            if (OP == OP_IMADWIDE) asm volatile("{ .reg .s32 lo; cvt.u32.u64 lo, %0; mad.wide.s32 %0, lo, %1, %0; }" : "+l"(w[c]) : "r"(b));
            if (OP == OP_LEAHI) asm volatile("{ .reg .u32 t; shr.u32 t, %0, 31; add.s32 %0, t, %1; }" : "+r"(v[c]) : "r"(b));
            if (OP == OP_VIMNMX) { if ((rc / CHAINS) & 1) asm volatile("max.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(b)); else asm volatile("add.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(a0)); }
            if (OP == OP_VIMNMX3) asm volatile("{ .reg .s32 t; max.s32 t, %0, %1; max.s32 %0, t, %2; }" : "+r"(v[c]) : "r"(b), "r"(a0));
            if (OP == OP_VIADDMNMX) asm volatile("{ .reg .s32 t; add.s32 t, %0, %1; min.s32 %0, t, %2; }" : "+r"(v[c]) : "r"(b), "r"(a0));
            if (OP == OP_VIMNMX16x2) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(b));
            if (OP == OP_I2IP) asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(b));
            if (OP == OP_IABS) { if ((rc / CHAINS) & 1) asm volatile("abs.s32 %0, %0;" : "+r"(v[c])); else asm volatile("sub.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(b)); }
            if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x9910;" : "+r"(v[c]) : "r"(b));
            if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(b), "r"(a0));
            if (OP == OP_SHF) asm volatile("shf.r.clamp.b32 %0, %0, %1, 5;" : "+r"(v[c]) : "r"(b));
            if (OP == OP_IADD3) asm volatile("{ .reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2; }" : "+r"(v[c]) : "r"(b), "r"(a0));
            if (OP == OP_MIX_ALU_FMA) {
                asm volatile("max.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(b));
                asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(u[c]) : "r"(b), "r"(a0));
            }
        }
    }
    long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc += v[c] + u[c] + (int)w[c] + (int)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, int *d_out, long long *d_cyc)
{
    const int ctas_per_sm = 4;            // 1024 threads per SM: 8 warps per scheduler
    const int grid = sms * ctas_per_sm;
    probe<OP><<<grid, 256>>>(d_out, d_cyc, 3, 5);
    probe<OP><<<grid, 256>>>(d_out, d_cyc, 3, 5);
    cudaDeviceSynchronize();
    std::vector<long long> c(grid);
    cudaMemcpy(c.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(c.begin(), c.end());
    double med = (double)c[grid / 2];
    double per_thread = (double)ITERS * CHAINS * REPS * (OP == OP_MIX_ALU_FMA ? 2 : 1);
    double lane_ops_per_clk_sm = per_thread * 256.0 * ctas_per_sm / med;
    printf("%-32s %8.1f lane-ops/clk/SM   (median %.0f cycles)\n", names[OP], lane_ops_per_clk_sm, med);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int *d_out; long long *d_cyc;
    cudaMalloc(&d_out, sizeof(int) * sms * 4 * 256);
    cudaMalloc(&d_cyc, sizeof(long long) * sms * 4);
    printf("%s, %d SMs\n", p.name, sms);
    run<OP_IMAD>(sms, d_out, d_cyc);
    run<OP_IMADHI>(sms, d_out, d_cyc);
    run<OP_IMADWIDE>(sms, d_out, d_cyc);
    run<OP_LEAHI>(sms, d_out, d_cyc);
    run<OP_VIMNMX>(sms, d_out, d_cyc);
    run<OP_VIMNMX3>(sms, d_out, d_cyc);
    run<OP_VIADDMNMX>(sms, d_out, d_cyc);
    run<OP_VIMNMX16x2>(sms, d_out, d_cyc);
    run<OP_I2IP>(sms, d_out, d_cyc);
    run<OP_IABS>(sms, d_out, d_cyc);
    run<OP_PRMT>(sms, d_out, d_cyc);
    run<OP_LOP3>(sms, d_out, d_cyc);
    run<OP_SHF>(sms, d_out, d_cyc);
    run<OP_IADD3>(sms, d_out, d_cyc);
    run<OP_MIX_ALU_FMA>(sms, d_out, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
