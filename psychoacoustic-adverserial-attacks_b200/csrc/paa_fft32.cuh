// Half-warp real FFT for n_fft = 1024 on sm_100a: one frame per 16 lanes, 32 complex points per lane.
//
// A real frame of 1024 samples is a complex FFT of N = 512 points z[m] = x[2m] + i x[2m+1] (paa_fft.cuh).  Here the
// 512 points live in a HALF warp: lane l (0..15) holds z[l + 16 q], q = 0..31, as 32 packed register pairs, so a warp
// instruction works on two frames and the transform needs ONE exchange through shared memory instead of two:
//
//   forward   DFT-32 over q in registers -> twiddle W_512^{l k1} -> exchange -> DFT-16 over l  (k = k1 + 32 k2)
//   inverse   IDFT-16 over k2            -> exchange -> twiddle conj    -> IDFT-32 over k1     (m = l + 16 q)
//
// After the exchange lane lam holds the two DFT-16 butterflies k1 = lam and k1 = 32 - lam (lane 0: 0 and 16), so both
// members of every conjugate pair (k, N - k) of the real-FFT split sit in ONE thread's registers (the "paired"
// assignment of paa_fft.cuh, with 32 in place of 64) and the spectral middle needs no shared memory.
// Per frame this is 128 shared-memory wavefronts of exchange traffic instead of 256, about two thirds of the
// instructions (index arithmetic, loop control and twiddle loads are shared by two frames) and the same flops.
//
// Exchange buffer: 32 rows (k1) x 16 columns (l), row stride 17 complex: rows are written / read by the 16 lanes as
// one contiguous 128-byte run, columns (stride 17) hit 16 distinct 8-byte slots mod 16 -- every access is conflict
// free and is (one per-lane base register) + (compile-time offset).  tools/emulate_fft32.py replays the index math.
#pragma once
#include "paa_fft.cuh"

namespace paa {

constexpr int kHwRow = 17;                       // row stride of the exchange buffer (complex)
constexpr int kHwBuf = 560;                      // complex per half-warp buffer: 32 x 17 = 544, padded so that the 8 buffers
                                                 // of a CTA are exactly one staged input span (8960 floats)

__host__ __device__ constexpr int hw_j1(int lam) { return lam ? 32 - lam : 16; }

// cos / sin (2 pi n / 32), n = 0..31, as literals (immediate operands after unrolling)
__host__ __device__ constexpr float c32(int n) {
    constexpr float t[9] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                            0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.f};
    n &= 31;
    if (n > 16) n = 32 - n;
    return n <= 8 ? t[n] : -t[16 - n];
}
__host__ __device__ constexpr float s32(int n) { return c32((n & 31) + 24); }          // sin x = cos(x - pi/2)

// v *= W_32^n (forward, DIR = -1) or its conjugate (inverse)
template <int DIR, int n>
__device__ __forceinline__ cpx tw32(cpx v) {
    constexpr int m = n & 31;
    if (m == 0) return v;
    if (m == 8) return rot90<DIR>(v);
    if (m == 16) return neg2(v);
    if (m == 24) return rot90<-DIR>(v);
    constexpr float wx = c32(m), wy = -s32(m);
    return cmul_tw<DIR>(v, wx, wy);
}

// 16-point DFT, natural order in and out:  n = 4a + b, k = c + 4d;  DFT-4 over a -> W_16^{bc} -> DFT-4 over b
template <int DIR>
__device__ __forceinline__ void dft16(cpx (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4<DIR>(v[b], v[4 + b], v[8 + b], v[12 + b]);      // Y[b][c] in v[4c + b]
    v[5] = tw32<DIR, 2>(v[5]);   v[6] = tw32<DIR, 4>(v[6]);    v[7] = tw32<DIR, 6>(v[7]);
    v[9] = tw32<DIR, 4>(v[9]);   v[10] = tw32<DIR, 8>(v[10]);  v[11] = tw32<DIR, 12>(v[11]);
    v[13] = tw32<DIR, 6>(v[13]); v[14] = tw32<DIR, 12>(v[14]); v[15] = tw32<DIR, 18>(v[15]);
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4<DIR>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);   // X[c + 4d] in v[4c + d]
    // to natural order (register renaming: (c, d) -> (d, c) is a transposition of the 4 x 4 array)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) { const cpx t = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = t; }
}

// The part of the 32-point DFT behind the first radix-4 level:  input Y[b][c] in v[8c + b] (b = 0..7 the low three bits
// of the input index q = 8a + b, c the output of the DFT-4 over a);  W_32^{bc} -> DFT-8 over b;  output X[c + 4d] in v[8c + d].
template <int DIR>
__device__ __forceinline__ void dft32_tail(cpx (&v)[32]) {
#define PAA_TW(c, b) v[8 * c + b] = tw32<DIR, c * b>(v[8 * c + b]);
    PAA_TW(1, 1) PAA_TW(1, 2) PAA_TW(1, 3) PAA_TW(1, 4) PAA_TW(1, 5) PAA_TW(1, 6) PAA_TW(1, 7)
    PAA_TW(2, 1) PAA_TW(2, 2) PAA_TW(2, 3) PAA_TW(2, 4) PAA_TW(2, 5) PAA_TW(2, 6) PAA_TW(2, 7)
    PAA_TW(3, 1) PAA_TW(3, 2) PAA_TW(3, 3) PAA_TW(3, 4) PAA_TW(3, 5) PAA_TW(3, 6) PAA_TW(3, 7)
#undef PAA_TW
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        cpx w[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) w[b] = v[8 * c + b];
        dft8<DIR>(w);
#pragma unroll
        for (int d = 0; d < 8; ++d) v[8 * c + d] = w[d];
    }
}
// index of output k (natural order) in the array dft32_tail leaves behind
__host__ __device__ constexpr int hw_out32(int k) { return 8 * (k & 3) + (k >> 2); }

// Plain 32-point DFT of v[q] (first level included), result in the hw_out32 order.
template <int DIR>
__device__ __forceinline__ void dft32(cpx (&v)[32]) {
#pragma unroll
    for (int b = 0; b < 8; ++b) dft4<DIR>(v[b], v[8 + b], v[16 + b], v[24 + b]);
    dft32_tail<DIR>(v);
}

}  // namespace paa
