// Micro-benchmark: do warp shuffles and shared-memory loads share one throughput budget on sm_100a?
// Modes: 0 = LDS.64 only (conflict free), 1 = SHFL only, 2 = both interleaved 1:1 (same totals as 0 plus 1).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
    __shared__ float2 buf[1][64 * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* b = buf[0]; (void)warp;
    for (int i = lane; i < 512; i += 32) b[i] = make_float2(i, -i);
    __syncthreads();
    float acc = 0.f, s = lane;
    const float2* p = b + lane;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0 || MODE == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p + 32 * u))); acc += v.x; }
            if (MODE == 1 || MODE == 2) { s = __shfl_xor_sync(0xffffffffu, s, 16); }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + s;
}
template <int MODE> void run(const char* name, int warps) {
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 4 * sms * 1024);
    const int iters = 20000;
    k<MODE><<<sms, warps * 32>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * clk * 1e3;                       // SM cycles at nominal clock
    const double per_warp_ops = (double)iters * 16 * warps;         // per SM: LDS.64 count (2 wavefronts each) and/or SHFL count
    printf("%-10s warps/SM %2d: %7.3f ms  cycles/SM per (LDS.64|SHFL|pair) = %.3f\n", name, warps, ms, cyc / per_warp_ops);
    cudaFree(out);
}
int main() {
    for (int w : {8, 16, 32}) { run<0>("LDS.64", w); run<1>("SHFL", w); run<2>("LDS+SHFL", w); }
    return 0;
}
