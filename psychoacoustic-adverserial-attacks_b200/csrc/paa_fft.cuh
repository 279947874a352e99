// Warp-per-frame real FFT for n_fft in {512, 1024} on sm_100a.
//
// A real frame of n_fft samples is transformed as a complex FFT of N = n_fft/2 points
// (z[m] = x[2m] + i x[2m+1]) followed by the usual split into the n_fft/2+1 one-sided bins.
// The complex FFT is a three-stage Stockham autosort (radix 8*8*8 for N=512, 4*8*8 for N=256):
// every lane owns N/32 points per stage in registers, stages exchange through a warp-private
// shared-memory buffer of N float2, and only __syncwarp is needed.  The buffer index is XOR
// swizzled so that every stage's strided stores and contiguous loads are bank-conflict free for
// 64-bit accesses (checked exhaustively by tools/bank_search.py).
#pragma once
#include <cuda_runtime.h>

namespace paa {

template <int NFFT> struct Plan;
template <> struct Plan<1024> {
    static constexpr int N = 512, R0 = 8, R1 = 8, R2 = 8;
    __device__ __forceinline__ static int swz(int i) {
        return i ^ (((i >> 4) & 1) | (((i >> 5) & 1) << 1) | (((i >> 6) & 1) * 12));
    }
};
template <> struct Plan<512> {
    static constexpr int N = 256, R0 = 4, R1 = 8, R2 = 8;
    __device__ __forceinline__ static int swz(int i) {
        return i ^ (((i >> 4) & 1) | (((i >> 5) & 1) * 6) | (((i >> 6) & 1) * 8));
    }
};

// per-lane twiddle table: stage 1 block then stage 2 block, each [NB][R-1][32] float2
template <int NFFT> struct TwLayout {
    using P = Plan<NFFT>;
    static constexpr int NB1 = P::N / P::R1 / 32, NB2 = P::N / P::R2 / 32;
    static constexpr int kStage1 = NB1 * (P::R1 - 1) * 32;
    static constexpr int kStage2 = NB2 * (P::R2 - 1) * 32;
    static constexpr int kTotal = kStage1 + kStage2;     // float2 entries
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by w (forward, DIR=-1) or conj(w) (inverse, DIR=+1); tables hold forward twiddles (cos, -sin)
template <int DIR>
__device__ __forceinline__ float2 cmul_tw(float2 v, float2 w) {
    if (DIR < 0) return make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
    return make_float2(v.x * w.x + v.y * w.y, v.y * w.x - v.x * w.y);
}
// multiply by -i (forward) / +i (inverse)
template <int DIR>
__device__ __forceinline__ float2 rot90(float2 v) {
    return DIR < 0 ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);
}

template <int DIR>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = rot90<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}

template <int DIR>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    dft4<DIR>(v[0], v[2], v[4], v[6]);        // even samples -> E[0..3] in v[0],v[2],v[4],v[6]
    dft4<DIR>(v[1], v[3], v[5], v[7]);        // odd  samples -> O[0..3] in v[1],v[3],v[5],v[7]
    const float h = 0.70710678118654752440f;
    float2 o1 = v[3], o2 = v[5], o3 = v[7];
    if (DIR < 0) {
        o1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));       // * (1-i)/sqrt2
        o2 = make_float2(o2.y, -o2.x);                               // * -i
        o3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));     // * (-1-i)/sqrt2
    } else {
        o1 = make_float2(h * (o1.x - o1.y), h * (o1.x + o1.y));       // * (1+i)/sqrt2
        o2 = make_float2(-o2.y, o2.x);                               // * +i
        o3 = make_float2(-h * (o3.x + o3.y), h * (o3.x - o3.y));     // * (-1+i)/sqrt2
    }
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

template <int R, int DIR> struct Dft;
template <int DIR> struct Dft<4, DIR> {
    __device__ __forceinline__ static void run(float2 (&v)[4]) { dft4<DIR>(v[0], v[1], v[2], v[3]); }
};
template <int DIR> struct Dft<8, DIR> {
    __device__ __forceinline__ static void run(float2 (&v)[8]) { dft8<DIR>(v); }
};

// One Stockham stage on register data: twiddle (k = j mod Ns), R-point DFT.
// v[b][r] holds input j + r*N/R of butterfly j = lane + 32 b.
template <int N, int R, int Ns, int DIR>
__device__ __forceinline__ void stage_compute(float2 (&v)[N / R / 32][R], const float2* __restrict__ tw, int lane) {
    constexpr int NB = N / R / 32;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[b][r] = cmul_tw<DIR>(v[b][r], tw[(b * (R - 1) + (r - 1)) * 32 + lane]);
        }
        Dft<R, DIR>::run(v[b]);
    }
}
template <class P, int R>
__device__ __forceinline__ void stage_load(float2 (&v)[P::N / R / 32][R], const float2* buf, int lane) {
#pragma unroll
    for (int b = 0; b < P::N / R / 32; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) v[b][r] = buf[P::swz(lane + 32 * b + r * (P::N / R))];
}
// output r of butterfly j goes to (j/Ns)*Ns*R + (j%Ns) + r*Ns
template <class P, int R, int Ns>
__device__ __forceinline__ void stage_store(const float2 (&v)[P::N / R / 32][R], float2* buf, int lane) {
#pragma unroll
    for (int b = 0; b < P::N / R / 32; ++b) {
        const int j = lane + 32 * b;
        const int base = (j / Ns) * Ns * R + (j % Ns);
#pragma unroll
        for (int r = 0; r < R; ++r) buf[P::swz(base + r * Ns)] = v[b][r];
    }
}

// Complex FFT of N points.  First-stage inputs come from `first` (functor m -> float2), the
// result of the last stage is handed to `last` (functor (m, float2)) in natural order m.
template <int NFFT, int DIR, class First, class Last>
__device__ __forceinline__ void fft_warp(float2* buf, const float2* __restrict__ tw, int lane, First first, Last last) {
    using P = Plan<NFFT>;
    using L = TwLayout<NFFT>;
    constexpr int N = P::N;
    {   // stage 0: Ns = 1, no twiddles
        float2 v[N / P::R0 / 32][P::R0];
#pragma unroll
        for (int b = 0; b < N / P::R0 / 32; ++b)
#pragma unroll
            for (int r = 0; r < P::R0; ++r) v[b][r] = first(lane + 32 * b + r * (N / P::R0));
        stage_compute<N, P::R0, 1, DIR>(v, tw, lane);
        __syncwarp();                       // buf may still be read by the previous user
        stage_store<P, P::R0, 1>(v, buf, lane);
    }
    __syncwarp();
    {   // stage 1: Ns = R0
        float2 v[N / P::R1 / 32][P::R1];
        stage_load<P, P::R1>(v, buf, lane);
        stage_compute<N, P::R1, P::R0, DIR>(v, tw, lane);
        __syncwarp();
        stage_store<P, P::R1, P::R0>(v, buf, lane);
    }
    __syncwarp();
    {   // stage 2: Ns = R0*R1 = N/R2, outputs j + r*Ns in natural order
        float2 v[N / P::R2 / 32][P::R2];
        stage_load<P, P::R2>(v, buf, lane);
        stage_compute<N, P::R2, P::R0 * P::R1, DIR>(v, tw + L::kStage1, lane);
        __syncwarp();
#pragma unroll
        for (int b = 0; b < N / P::R2 / 32; ++b)
#pragma unroll
            for (int r = 0; r < P::R2; ++r) last(lane + 32 * b + r * (P::R0 * P::R1), v[b][r]);
    }
}

}  // namespace paa
