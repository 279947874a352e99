// Micro-benchmark: scalar FADD/FFMA vs packed add.f32x2 / fma.f32x2 issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
template <int MODE, int ILP>
__global__ void k(float* out, int iters, float s) {
    float a[ILP], b[ILP];
    u64 p[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = a[i] * 0.5f; p[i] = pk(a[i], b[i]); }
    const u64 ps = pk(s, s * 1.0001f), pm = pk(0.999f, 1.0001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) { a[i] += s; b[i] += s; }                                   // 2 FADD
                if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));      // 1 FADD2
                if (MODE == 2) { a[i] = fmaf(a[i], 0.999f, s); b[i] = fmaf(b[i], 1.0001f, s); }      // 2 FFMA
                if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pm), "l"(ps));
            }
    }
    float r = 0.f;
    for (int i = 0; i < ILP; ++i) r += (MODE & 1) ? lo(p[i]) : a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE, int ILP>
void run(const char* name, int warps_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 1024);
    const int iters = 20000, threads = warps_per_sm * 32;
    k<MODE, ILP><<<sms, threads>>>(out, 100, 1e-3f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, ILP><<<sms, threads>>>(out, iters, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)sms * threads * iters * 8.0 * ILP * 2.0;   // scalar-equivalent add/fma per lane
    const double per_clk_sm = lane_ops / (ms * 1e-3) / (clk * 1e3) / sms;
    printf("%-8s ILP %d warps/SM %2d: %8.3f ms  %6.1f lane-ops/clk/SM (at nominal %d MHz)  %.2f T lane-ops/s\n", name, ILP,
           warps_per_sm, ms, per_clk_sm, clk / 1000, lane_ops / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0, 8>("FADD", w); run<1, 8>("FADD2", w); run<2, 8>("FFMA", w); run<3, 8>("FFMA2", w);
    }
    run<0, 1>("FADD", 4); run<1, 1>("FADD2", 4); run<2, 1>("FFMA", 4); run<3, 1>("FFMA2", 4);   // latency
    return 0;
}
