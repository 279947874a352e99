// Warp-per-frame real FFT for n_fft in {512, 1024} on sm_100a.
//
// A real frame of n_fft samples is transformed as a complex FFT of N = n_fft/2 points
// (z[m] = x[2m] + i x[2m+1]) followed by the usual split into the n_fft/2+1 one-sided bins.
// The complex FFT is a three-stage Stockham autosort (radix 8*8*8 for N=512, 4*8*8 for N=256):
// every lane owns N/32 points per stage in registers -- each complex point one 64-bit register pair, all arithmetic
// in sm_100a's packed fp32 instructions (FADD2 / FMUL2 / FFMA2, see "packed complex arithmetic" below) -- stages
// exchange through a warp-private shared-memory buffer, and only __syncwarp is needed.
//
// Buffer layout: the two exchanges use two additive paddings of the point index (padc, Plan::padB).  Additive means
// every access of a stage is  (one per-lane base register) + (compile-time offset)  and costs no address arithmetic;
// both make every load and store of their exchange bank-conflict free (tools/bank_search.py conventions: a 64-bit
// access is served per half-warp, conflict free when 16 lanes hit 16 distinct 8-byte slots mod 16).
//
// n_fft 1024 additionally has a "paired" form of the last forward / first inverse stage (paired_j1 below) that
// leaves the spectrum in registers with both members of every conjugate pair in the same thread.
#pragma once
#include <cuda_runtime.h>

namespace paa {

template <int NFFT> struct Plan;
// padB: the layout of the second exchange (stage-1 stores / stage-2 loads).  Each exchange is a complete
// write-then-read of the buffer, so it may use its own padding: padc (= i + i/16) makes the stride-R0 stores of
// stage 0 conflict free, padB makes the (j/R0)*R0*R1 + j%R0 + r*R0 stores of stage 1 conflict free, and both keep
// the contiguous loads conflict free and every address = per-lane base + compile-time offset.
template <> struct Plan<1024> {
    static constexpr int N = 512, R0 = 8, R1 = 8, R2 = 8;
    // 8 spare slots per 64: conflict free for the stage-1 stores, the contiguous stage-2 loads AND the descending
    // runs {32, 63..49}, {48..33} (+64 r) of the paired butterflies (search: tools/bank_search.py conventions)
    __host__ __device__ static constexpr int padB(int i) { return i + 8 * (i >> 6); }
};
template <> struct Plan<512> {
    static constexpr int N = 256, R0 = 4, R1 = 8, R2 = 8;
    __host__ __device__ static constexpr int padB(int i) { return i + 4 * (i >> 5); }
};

__host__ __device__ constexpr int padc(int i) { return i + (i >> 4); }
template <int NFFT> struct BufLayout {
    static constexpr int kFloat2 = ((Plan<NFFT>::padB(Plan<NFFT>::N - 1) + 2) / 2) * 2;      // padB is the wider of the two
};

// per-lane twiddle tables.  Stage 1: one [4][32] block of float4 (k = j mod R0 does not depend on b), the forward
// twiddles (cos, -sin) of r = 2q and r = 2q+1, kept in registers by the kernel.  Stage 2 (compact): per block
// (b = 0, b = 1, and for n_fft 1024 the paired b = 1) only W^k, W^2k, W^4k as float2 [3][32]; the other four powers
// are products of two of them (W^3k = W^k W^2k, W^5k = W^k W^4k, W^6k = W^2k W^4k, W^7k = W^3k W^4k: at most two
// extra roundings) -- 8 packed instructions per block instead of 10 more shared-memory wavefronts.
template <int NFFT> struct TwLayout {
    using P = Plan<NFFT>;
    static constexpr int NB1 = P::N / P::R1 / 32, NB2 = P::N / P::R2 / 32;
    static constexpr int kStage1 = 4 * 32;               // float4 entries
    static constexpr int kStage2Blocks = NB2 + ((NFFT == 1024) ? 1 : 0);
    static constexpr int kStage2 = kStage2Blocks * 3 * 32 / 2;          // float4 entries (3 x 32 float2 per block)
    static constexpr int kTotal = kStage1 + kStage2;
};

// Paired butterfly assignment (n_fft 1024, N = 512 = 8*8*8).  The last forward stage and the first inverse stage
// run 64 radix-8 butterflies j = 0..63 on the points j + 64 r.  Bin k = j + 64 r has its conjugate partner N - k in
// butterfly 64 - j, so lane l takes the two butterflies  j0 = l  and  j1 = 64 - l  and every (k, N-k) pair of the
// real-FFT split / merge lives in ONE thread's registers: the spectral middle needs no shared memory at all.
// Lane 0 takes the two self-paired butterflies 0 and 32.
__host__ __device__ constexpr int paired_j1(int lane) { return lane ? 64 - lane : 32; }

// ---- packed complex arithmetic -------------------------------------------------------------------
// A complex value is one 64-bit register pair (re, im).  sm_100a has packed fp32 instructions (FADD2 / FMUL2 / FFMA2,
// PTX add/mul/fma.rn.f32x2) whose operands take a half swap, per-half negation and scalar broadcast for free, so a
// complex add, a +-i rotation folded into an add, or half a twiddle multiply is ONE issue slot instead of two
// (same IEEE roundings as the scalar forms; tools/micro/f32x2_bench.cu: same lane throughput, half the issue).
// The pk()/up() repacks below never cost an instruction when they feed a packed op: ptxas folds them into the
// operand modifiers (.LO_HI, .NP, .F32).
typedef unsigned long long cpx;
__device__ __forceinline__ cpx pk(float x, float y) { cpx r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void up(cpx v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ float cre(cpx v) { float x, y; up(v, x, y); return x; }
__device__ __forceinline__ float cim(cpx v) { float x, y; up(v, x, y); return y; }
__device__ __forceinline__ cpx add2(cpx a, cpx b) { cpx r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ cpx sub2(cpx a, cpx b) { cpx r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ cpx mul2(cpx a, cpx b) { cpx r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ cpx fma2(cpx a, cpx b, cpx c) { cpx r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ cpx bcast(float s) { return pk(s, s); }
__device__ __forceinline__ cpx conj2(cpx v) { float x, y; up(v, x, y); return pk(x, -y); }
__device__ __forceinline__ cpx swap2(cpx v) { float x, y; up(v, x, y); return pk(y, x); }
__device__ __forceinline__ cpx neg2(cpx v) { float x, y; up(v, x, y); return pk(-x, -y); }
// multiply by -i (DIR < 0, the forward transform's rotation) / +i (DIR > 0)
template <int DIR>
__device__ __forceinline__ cpx rot90(cpx v) {
    float x, y;
    up(v, x, y);
    return DIR < 0 ? pk(y, -x) : pk(-y, x);
}
// multiply by w (forward, DIR=-1) or conj(w) (inverse, DIR=+1); tables hold forward twiddles (cos, -sin):
//   forward (vx wx - vy wy, vx wy + vy wx) = v*wx + (+i v)*wy,   inverse (vx wx + vy wy, vy wx - vx wy) = v*wx + (-i v)*wy
template <int DIR>
__device__ __forceinline__ cpx cmul_tw(cpx v, float wx, float wy) {
    return fma2(rot90<-DIR>(v), bcast(wy), mul2(v, bcast(wx)));
}

// product of two complex numbers (used to derive twiddle powers: the product of two forward twiddles is one)
__device__ __forceinline__ cpx cmul(cpx a, cpx b) {
    return fma2(rot90<+1>(a), bcast(cim(b)), mul2(a, bcast(cre(b))));
}

template <int DIR>
__device__ __forceinline__ void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
    const cpx t0 = add2(a0, a2), t1 = sub2(a0, a2), t2 = add2(a1, a3), t3 = sub2(a1, a3);
    a0 = add2(t0, t2); a2 = sub2(t0, t2);
    a1 = add2(t1, rot90<DIR>(t3)); a3 = sub2(t1, rot90<DIR>(t3));
}

template <int DIR>
__device__ __forceinline__ void dft8(cpx (&v)[8]) {
    dft4<DIR>(v[0], v[2], v[4], v[6]);        // even samples -> E[0..3] in v[0],v[2],v[4],v[6]
    dft4<DIR>(v[1], v[3], v[5], v[7]);        // odd  samples -> O[0..3] in v[1],v[3],v[5],v[7]
    const float h = 0.70710678118654752440f;
    // O1 * (1 -+ i)/sqrt2 = h (O1 + rot O1),  O3 * (-1 -+ i)/sqrt2 = -h (O3 - rot O3),  O2 * (-+i) = rot O2;
    // the factor h rides the final butterfly as an FFMA2
    const cpx s1 = add2(v[3], rot90<DIR>(v[3])), s3 = sub2(v[7], rot90<DIR>(v[7]));
    const cpx e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o2 = v[5];
    v[0] = add2(e0, o0);                  v[4] = sub2(e0, o0);
    v[1] = fma2(s1, bcast(h), e1);        v[5] = fma2(s1, bcast(-h), e1);
    v[2] = add2(e2, rot90<DIR>(o2));      v[6] = sub2(e2, rot90<DIR>(o2));
    v[3] = fma2(s3, bcast(-h), e3);       v[7] = fma2(s3, bcast(h), e3);
}

template <int R, int DIR> struct Dft;
template <int DIR> struct Dft<4, DIR> {
    __device__ __forceinline__ static void run(cpx (&v)[4]) { dft4<DIR>(v[0], v[1], v[2], v[3]); }
};
template <int DIR> struct Dft<8, DIR> {
    __device__ __forceinline__ static void run(cpx (&v)[8]) { dft8<DIR>(v); }
};

// Per-lane base offsets into the padded buffer (float2 units), computed once per warp.
//   ld : contiguous accesses  lane + c, c a multiple of 32          -> ld + padc(c)
//   s0 : stage-0 stores       R0*j + r,  j = lane + 32 b            -> s0 + r + padc(32*R0*b)
//   s1 : stage-1 stores       (j/R0)*R0*R1 + j%R0 + r*R0            -> s1 + padB(r*R0) + padB(32*R1*b)   (second layout)
//   ldB: stage-2 loads        lane + c                              -> ldB + padB(c)
//   ldB1, s01: the same for the paired butterfly j1 (b = 1 of the paired stages)
template <int NFFT>
struct LaneBase {
    int ld, ldB, s0, s1, ldB1, s01;
    __device__ __forceinline__ explicit LaneBase(int lane) {
        using P = Plan<NFFT>;
        ld = padc(lane);
        ldB = P::padB(lane);
        s0 = padc(P::R0 * lane);
        const int i1 = (lane / P::R0) * P::R0 * P::R1 + (lane % P::R0);
        s1 = P::padB(i1);
        ldB1 = P::padB(paired_j1(lane));
        s01 = padc(P::R0 * paired_j1(lane));
    }
};

// One Stockham stage on register data: twiddle (k = j mod Ns), R-point DFT.
// v[b][r] holds input j + r*N/R of butterfly j = lane + 32 b.
template <int N, int R, bool TWIDDLE, bool PER_B, int DIR>
__device__ __forceinline__ void stage_compute(cpx (&v)[N / R / 32][R], const float4* __restrict__ tw, int lane) {
    constexpr int NB = N / R / 32;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (TWIDDLE) {
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
                const float4 w = tw[((PER_B ? b : 0) * 4 + q) * 32 + lane];
                if (q > 0) v[b][2 * q] = cmul_tw<DIR>(v[b][2 * q], w.x, w.y);
                v[b][2 * q + 1] = cmul_tw<DIR>(v[b][2 * q + 1], w.z, w.w);
            }
        }
        Dft<R, DIR>::run(v[b]);
    }
}

// The same with the twiddles of the stage in registers: stage 1 has k = j mod R0 = lane mod R0 for every b, so one
// set of R-1 twiddles serves the whole tile (loaded once per warp, 32 shared-memory wavefronts per frame saved).
// W: anything indexable as w[q] -> float4 (a register array, or Tw1Shared re-reading the table per transform)
struct Tw1Shared {
    const float4* p;                    // &table[lane]
    __device__ __forceinline__ float4 operator[](int q) const { return p[q * 32]; }
};
template <int N, int R, int DIR, class W>
__device__ __forceinline__ void stage_compute_reg(cpx (&v)[N / R / 32][R], const W& w) {
    float4 t[R / 2];
#pragma unroll
    for (int q = 0; q < R / 2; ++q) t[q] = w[q];
#pragma unroll
    for (int b = 0; b < N / R / 32; ++b) {
#pragma unroll
        for (int q = 0; q < R / 2; ++q) {
            if (q > 0) v[b][2 * q] = cmul_tw<DIR>(v[b][2 * q], t[q].x, t[q].y);
            v[b][2 * q + 1] = cmul_tw<DIR>(v[b][2 * q + 1], t[q].z, t[q].w);
        }
        Dft<R, DIR>::run(v[b]);
    }
}

// ---- the three Stockham stages as separate pieces ---------------------------------------------------------------
// stage 0 (Ns = 1, no twiddles): v[b][r] = input j + r*N/R0 of butterfly j; outputs go to buf at R0*j + r (first layout)
template <int NFFT, int DIR, bool PAIRED>
__device__ __forceinline__ void fft_stage0(cpx* buf, cpx (&v)[Plan<NFFT>::N / Plan<NFFT>::R0 / 32][Plan<NFFT>::R0], int lane,
                                           const LaneBase<NFFT>& lb) {
    using P = Plan<NFFT>;
    constexpr int N = P::N, R = P::R0, NB = N / R / 32;
    stage_compute<N, R, false, false, DIR>(v, nullptr, lane);
    __syncwarp();                       // buf may still be read by the previous user
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (PAIRED && b == 1) buf[lb.s01 + r] = v[b][r];
            else buf[lb.s0 + r + padc(32 * R * b)] = v[b][r];
        }
    __syncwarp();
}
// stage 1 (Ns = R0): buf -> registers -> buf (second layout)
template <int NFFT, int DIR, class W>
__device__ __forceinline__ void fft_stage1(cpx* buf, const W& tw1, const LaneBase<NFFT>& lb) {
    using P = Plan<NFFT>;
    constexpr int N = P::N, R = P::R1, NB = N / R / 32;
    cpx v[NB][R];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) v[b][r] = buf[lb.ld + padc(32 * b + r * (N / R))];
    stage_compute_reg<N, R, DIR>(v, tw1);
    __syncwarp();
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r) buf[lb.s1 + P::padB(r * P::R0) + P::padB(32 * R * b)] = v[b][r];
    __syncwarp();
}
// stage 2 (Ns = R0*R1 = N/R2): buf -> registers; v[b][r] = output j + r*Ns of butterfly j (natural order).
// PAIRED: b = 1 is butterfly j1(lane) with its own twiddle block (table index 2).
template <int NFFT, int DIR, bool PAIRED>
__device__ __forceinline__ void fft_stage2(const cpx* buf, const float4* __restrict__ tw, cpx (&v)[Plan<NFFT>::N / Plan<NFFT>::R2 / 32][Plan<NFFT>::R2],
                                           int lane, const LaneBase<NFFT>& lb) {
    using P = Plan<NFFT>;
    using L = TwLayout<NFFT>;
    constexpr int N = P::N, R = P::R2, NB = N / R / 32;
    const float2* t2 = reinterpret_cast<const float2*>(tw + L::kStage1);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            v[b][r] = (PAIRED && b == 1) ? buf[lb.ldB1 + P::padB(r * (N / R))] : buf[lb.ldB + P::padB(32 * b + r * (N / R))];
        // twiddles W^{k r}, r = 1..7, from W^k, W^2k, W^4k
        const float2* tb = t2 + ((PAIRED && b == 1) ? 2 : b) * 96 + lane;
        const float2 f1 = tb[0], f2 = tb[32], f4 = tb[64];
        const cpx w1 = pk(f1.x, f1.y), w2 = pk(f2.x, f2.y), w4 = pk(f4.x, f4.y);
        const cpx w3 = cmul(w1, w2), w5 = cmul(w1, w4), w6 = cmul(w2, w4), w7 = cmul(w3, w4);
        v[b][1] = cmul_tw<DIR>(v[b][1], cre(w1), cim(w1));
        v[b][2] = cmul_tw<DIR>(v[b][2], cre(w2), cim(w2));
        v[b][3] = cmul_tw<DIR>(v[b][3], cre(w3), cim(w3));
        v[b][4] = cmul_tw<DIR>(v[b][4], cre(w4), cim(w4));
        v[b][5] = cmul_tw<DIR>(v[b][5], cre(w5), cim(w5));
        v[b][6] = cmul_tw<DIR>(v[b][6], cre(w6), cim(w6));
        v[b][7] = cmul_tw<DIR>(v[b][7], cre(w7), cim(w7));
        Dft<R, DIR>::run(v[b]);
    }
}

// Complex FFT of N points.  First-stage inputs come from `first` (functor (m, c) -> cpx) and the
// result of the last stage is handed to `last` (functor (m, c, cpx)) in natural order; m = lane + c
// with c a compile-time multiple of 32, so callers can address  base(lane) + c.
template <int NFFT, int DIR, class W, class First, class Last>
__device__ __forceinline__ void fft_warp(cpx* buf, const float4* __restrict__ tw, const W& tw1,
                                         int lane, const LaneBase<NFFT>& lb, First first, Last last) {
    using P = Plan<NFFT>;
    constexpr int N = P::N;
    {
        constexpr int R = P::R0, NB = N / R / 32;
        cpx v[NB][R];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int r = 0; r < R; ++r) v[b][r] = first(lane + 32 * b + r * (N / R), 32 * b + r * (N / R));
        fft_stage0<NFFT, DIR, false>(buf, v, lane, lb);
    }
    fft_stage1<NFFT, DIR>(buf, tw1, lb);
    {
        constexpr int R = P::R2, NB = N / R / 32;
        cpx v[NB][R];
        fft_stage2<NFFT, DIR, false>(buf, tw, v, lane, lb);
        __syncwarp();                   // `last` may write buf
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int r = 0; r < R; ++r) last(lane + 32 * b + r * (P::R0 * P::R1), 32 * b + r * (P::R0 * P::R1), v[b][r]);
    }
}

// n_fft 1024: forward transform that leaves the half-length spectrum in registers in the paired assignment,
//   z[0][r] = Z[lane + 64 r],   z[1][r] = Z[j1(lane) + 64 r]
template <class W, class First>
__device__ __forceinline__ void fft_forward_paired(cpx* buf, const float4* __restrict__ tw, const W& tw1, int lane,
                                                   const LaneBase<1024>& lb, First first, cpx (&z)[2][8]) {
    {
        cpx v[2][8];
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int r = 0; r < 8; ++r) v[b][r] = first(lane + 32 * b + r * 64, 32 * b + r * 64);
        fft_stage0<1024, -1, false>(buf, v, lane, lb);
    }
    fft_stage1<1024, -1>(buf, tw1, lb);
    fft_stage2<1024, -1, true>(buf, tw, z, lane, lb);
}
// ... and the inverse that starts from those registers; `last` as in fft_warp (natural order, m = lane + c)
template <class W, class Last>
__device__ __forceinline__ void fft_inverse_paired(cpx* buf, const float4* __restrict__ tw, const W& tw1, int lane,
                                                   const LaneBase<1024>& lb, cpx (&z)[2][8], Last last) {
    fft_stage0<1024, +1, true>(buf, z, lane, lb);
    fft_stage1<1024, +1>(buf, tw1, lb);
    cpx v[2][8];
    fft_stage2<1024, +1, false>(buf, tw, v, lane, lb);
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int r = 0; r < 8; ++r) last(lane + 32 * b + r * 64, 32 * b + r * 64, v[b][r]);
}

}  // namespace paa
