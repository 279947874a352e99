// STFT-domain half of the hot path for sm_100a (reference: src/training_utils/train.py:38-66,
// src/core/fourier_transforms.py:4-41, src/core/projections.py:68-159).
//
// One kernel template covers every use:
//   SRC_TIME + SINK_TIME   fused  [PGD step ->] STFT -> per-bin op -> ISTFT -> align   (spectrum never in HBM)
//   SRC_TIME + SINK_REDUCE [PGD step ->] STFT -> fletcher_munson weighted power, per-block partials
//   SRC_TIME + SINK_SPEC   compute_stft          SRC_SPEC + SINK_TIME   compute_istft
//
// Geometry.  centre=True: frame t covers samples [t*hop - n/2, t*hop + n/2) of the reflect-padded
// row; output sample n receives the R = n_fft/hop frames t in (n/hop + R/2 - R, n/hop + R/2].
// A CTA of 8 warps owns S = FT-R+1 consecutive output hop-blocks and transforms the FT = 8*R*Q
// frames that touch them (R-1 halo frames are recomputed by the neighbour tile).  The input span is staged once in
// shared memory -- interior tiles by two TMA bulk copies (cp.async.bulk + mbarrier: the perturbation span into the
// tile buffer, the gradient span into the still idle FFT buffers, then a shared-memory pass applies the PGD step),
// row-end tiles by float4 loads with reflect padding -- while the CTA prefetches the next tile of its slot into L2.
// Every warp runs whole frames (paa_fft.cuh, packed-fp32 complex arithmetic); for n_fft 1024 the spectrum stays in
// registers between the forward and the inverse transform (paired butterflies, middle_paired).  The inverse frames
// are overlap-added into a shared accumulator in R phases: frames with equal t mod R never overlap, so plain
// read-modify-write is race free and the summation order is fixed (deterministic output).
// Twiddle tables and the window arrive by the same mbarrier-tracked TMA bulk copies.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>
#include <vector>
#include "paa_fft.cuh"
#include "paa_fft32.cuh"
#include "paa_internal.h"

namespace {

using namespace paa;

constexpr int kWarps = 8;
constexpr int kThreadsStft = kWarps * 32;

enum { SRC_TIME = 0, SRC_SPEC = 1 };
enum { SINK_TIME = 0, SINK_SPEC = 1, SINK_REDUCE = 2 };
enum { OP_NONE = 0, OP_MASK = 1, OP_PHON = 2, OP_SCALE = 3, OP_PHON_DB = 4 };

struct StftArgs {
    // time-domain source [rows, T]
    const float* x;
    const float* grad;     // PGD step fused into the tile load when non-null
    float lr;
    float* q_out;          // SINK_REDUCE: where the stepped perturbation is stored (nullable)
    int rows, T, n_frames, hop, R, Q;
    int frames_per_tile;   // FT
    int blocks_per_tile;   // S (SINK_TIME) -- owned output hop blocks
    int tiles_per_row;
    int vec_ok;            // rows are 16-byte aligned: float4 tile loads allowed
    int pf_stride;         // CTAs resident at once (SMs x CTAs per SM): CTA i prefetches the input of CTA i + pf_stride into L2
    // time-domain sink [rows, out_len]
    float* y;
    int out_len;
    // tables
    const void* blob;
    unsigned blob_bytes, off_tw, off_post, off_win;   // blob = [twiddles | split twiddles | window/2]; [0, off_win) stays resident
    // per-bin op
    float bin_hz, f_min, f_max;
    int k_lo, k_hi;        // min_max_freqs as bin indices: keep k < k_lo or k >= k_hi
    unsigned dead_slots;   // n_fft 1024, min_max_freqs: bit r set = every bin slot r of middle_paired touches is masked (r = 1..7)
    const float* spl_thresh;
    float ref_db;
    const float* scalars;  // OP_SCALE: scalars[PAA_S_SCALE]
    // spectrum source / sink, element strides in complex numbers
    const float2* spec_in;
    float2* spec_out;
    long long sb, sf, st;
    // fletcher_munson
    const float* fm_blob;  // [64 floats: phon knots][fm_np rows of paa_fm_stride(F) floats: w(knot i, f_k), out-of-band bins = fill]
    unsigned fm_blob_bytes;
    int fm_np, fm_uniform;
    float fm_fill, fm_k0, fm_inv_dk, fm_klast;
    double* partials;
    // k_stft_hw: (cos, sin)(2 pi k / n_fft) and the full true Hann window, global memory
    const void* post_table;
    const float* win_full;
};

// ---- TMA 1-D bulk copy + mbarrier (PTX) -----------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!ok);
}

// ---- per-bin operators ------------------------------------------------------------------------
// project_min_max_freqs (projections.py:74-79): keep f < f_min or f > f_max, zero the band between.
// f_k = fp32(k) * fp32(bin_hz) is monotone in k, so the host turns the two float comparisons into
// bin indices (keep k < k_lo or k >= k_hi) with the same fp32 arithmetic.
template <bool SCALED>
__device__ __forceinline__ void op_mask(const StftArgs& a, int k, float c, float& re, float& im) {
    const float keep = (k < a.k_lo || k >= a.k_hi) ? (SCALED ? c : 1.f) : 0.f;
    re *= keep;
    im *= keep;
}
// project_phon_level (projections.py:142-153) exactly as written: every bin takes the dB round trip
//   |X| -> 20 log10(|X|+1e-8) -> min(., thr) -> 10^(./20), phase kept.   thr = scaled threshold in dB.
__device__ __forceinline__ void op_phon_db(const float* thr, int k, float& re, float& im) {
    const float m = sqrtf(fmaf(re, re, im * im));
    const float db = 20.f * log10f(m + 1e-8f);
    const float t = thr[k];
    const float db2 = (db > t) ? t : db;
    const float mag = exp10f(db2 / 20.f);
    if (m > 1e-18f) {
        re = mag * (re / m);
        im = mag * (im / m);
    } else {                     // zero / denormal magnitude: the phase comes from atan2 as in torch.angle
        const float ph = atan2f(im, re);
        float s, c;
        sincosf(ph, &s, &c);
        re = mag * c;
        im = mag * s;
    }
}
// The same projection in the linear domain, used by the fused kernel: with lim[k] = 10^(thr[k]/20),
//   |X'| = min(|X| + 1e-8, lim[k])   -- the dB round trip is the identity on un-clipped bins up to
// fp32 rounding (~1e-6 relative, inside the 1e-5 parity bar) and log10 is monotone, so no log/exp per bin.
// `c` is the power-of-two factor 2/n_fft the inverse transform wants folded in; lim_c[k] = c * lim[k].
// This is the exact form (true square root; phase from atan2 like torch.angle when the magnitude is zero or
// denormal); the hot path is apply_op2's one-MUFU form, which hands frames with such bins over to this one.
// No warp-level primitives: it also runs under the lane-0 divergence of the register-resident middle.
__device__ __forceinline__ void op_phon(const float* lim_c, int k, float c, float& re, float& im) {
    const float P = fmaf(re, re, im * im);
    const float m = sqrtf(P);
    const float xc = fmaf(m, c, 1e-8f * c);           // c * (|X| + 1e-8)
    const float l = lim_c[k];
    const float mag = (xc > l) ? l : xc;              // NaN falls through like torch.where(db > thr, ...)
    if (P > 1e-36f) {
        const float g = mag * rsqrtf(P);
        re *= g;
        im *= g;
    } else {
        const float ph = atan2f(im, re);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        re = mag * cs;
        im = mag * sn;
    }
}
// compute_fm_weighted_norm_interp (projections.py:93-113): P * w(10 log10(P+1e-10), f_k)
// `tab` is the fm blob (global memory for the element-wise kernel, shared memory in the fused kernel).
template <bool FAST>
__device__ __forceinline__ float fm_term(const StftArgs& a, const float* __restrict__ tab, int F, int k, float re, float im) {
    // torch: power = abs(X)**2 (a square root and a square); FAST keeps re^2+im^2, one rounding away from it
    const float P = FAST ? fmaf(re, re, im * im) : [&] { const float m = sqrtf(fmaf(re, re, im * im)); return m * m; }();
    if (FAST && a.fm_uniform) {
        // Equally spaced phon knots (the reference's 0, 10, .., 90) in the fused kernel: branch free.  t = position of
        // spl = 10 log10(P + 1e-10) on the knot axis in cells (one MUFU.LG2: abs error ~1e-6 dB, it only positions the
        // query inside a 10 dB cell); cell index = round(t - 1/2) through the 1.5 * 2^23 magic constant (no F2I / I2F:
        // at an exact knot either neighbouring cell interpolates to the same value, ties are harmless); out-of-range
        // queries (and bins outside the frequency axis, whose table rows hold `fill`) take `fill`.
        const float t = fmaf(__log2f(P + 1e-10f), 3.0102999566398120f * a.fm_inv_dk, -a.fm_k0 * a.fm_inv_dk);
        const bool inr = !(t < 0.f) && !(t > (float)(a.fm_np - 1));
        const float tc = inr ? t : 0.f;
        const float tm = (tc - 0.5f) + 12582912.f;                 // 1.5 * 2^23: ulp 1, round to nearest even
        const int i = min(__float_as_int(tm) & 0xff, a.fm_np - 2);  // 0 <= tc <= np-1  =>  0 <= i <= np-2 already
        const float tp = tc - (tm - 12582912.f);
        const float c0 = tab[64 + i * F + k], c1 = tab[64 + (i + 1) * F + k];
        const float w = fmaf(tp, c1 - c0, c0);
        return P * (inr ? w : a.fm_fill);
    }
    // FAST: MUFU.LG2 (abs error ~1e-6 dB, it only positions the query inside a 10 dB cell)
    const float spl = FAST ? 3.0102999566398120f * __log2f(P + 1e-10f) : 10.f * log10f(P + 1e-10f);
    float w = a.fm_fill;
    if (!(spl < a.fm_k0) && !(spl > a.fm_klast)) {       // bins outside the frequency axis hold `fill` in every row
        int i;
        float tp;
        if (a.fm_uniform) {
            // equally spaced phon knots (the reference's 0,10,..,90): cell and fraction by arithmetic, no table reads
            const float t = (spl - a.fm_k0) * a.fm_inv_dk;
            i = max(0, min((int)t, a.fm_np - 2));
            tp = t - (float)i;
        } else {
            i = 0;
            for (int q = 1; q < a.fm_np - 1; ++q) i += (spl > tab[q]) ? 1 : 0;
            i = max(0, min(i, a.fm_np - 2));
            const float k0 = tab[i], k1 = tab[i + 1];
            tp = FAST ? __fdividef(spl - k0, k1 - k0) : (spl - k0) / (k1 - k0);
        }
        const float c0 = tab[64 + i * F + k], c1 = tab[64 + (i + 1) * F + k];
        w = fmaf(tp, c1 - c0, c0);
    }
    return P * w;
}

// SCALED: the fused kernel folds the inverse transform's power-of-two factor c = 2/n_fft into the operator
// (the forward window table holds w/2, so c * w/2 = w/n_fft); the element-wise spectrum kernels do not.
template <int OP, bool SCALED>
__device__ __forceinline__ void apply_op(const StftArgs& a, const float* tbl, float scale, float c, int k, float& re,
                                         float& im) {
    if (OP == OP_MASK) op_mask<SCALED>(a, k, c, re, im);
    else if (OP == OP_PHON) op_phon(tbl, k, SCALED ? c : 1.f, re, im);
    else if (OP == OP_PHON_DB) op_phon_db(tbl, k, re, im);
    else if (OP == OP_SCALE) { re *= scale; im *= scale; }          // caller pre-multiplies scale by c
    else if (SCALED) { re *= c; im *= c; }
}

// ---- packed per-bin operators of the fused kernel ---------------------------------------------------
// A real gain per bin commutes with conjugation, so the same function serves X[k] and conj(X[N-k]).
// OP_PHON: |X'| = min(|X| + 1e-8, lim[k]) with one MUFU.RSQ (|X| = P rsqrt(P), unit phasor X rsqrt(P)); `bad` is
// raised when P is zero / denormal / NaN, and the caller then redoes the whole frame with op_phon (exact magnitude,
// phase from atan2 like torch.angle) -- one vote per frame instead of one per bin.
__device__ __forceinline__ float rsqrt_ftz(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int OP, bool SCALED>
__device__ __forceinline__ cpx apply_op2(const StftArgs& a, const float* tbl, float scale, float c, int k, cpx X, bool& bad) {
    if (OP == OP_MASK) return mul2(X, bcast((k < a.k_lo || k >= a.k_hi) ? (SCALED ? c : 1.f) : 0.f));
    if (OP == OP_SCALE) return mul2(X, bcast(scale));               // caller pre-multiplies scale by c
    if (OP == OP_PHON) {
        float re, im;
        up(X, re, im);
        const float P = fmaf(re, re, im * im);
        bad = bad || !(P > 1e-36f);
        const float r = rsqrt_ftz(P);
        const float xc = fmaf(P * r, c, 1e-8f * c);                 // c * (|X| + 1e-8)
        return mul2(X, bcast(fminf(xc, tbl[k]) * r));               // a NaN bin stays NaN through r
    }
    return SCALED ? mul2(X, bcast(c)) : X;
}

// ---- the spectral middle of one frame -----------------------------------------------------------
// Works on conjugate-symmetric pairs (k, N-k) of the half-length complex FFT held in `buf` in natural, un-padded
// order (ascending and descending runs of 32 are both bank-conflict free):
//   forward split  Z -> X[k], conj X[N-k]   |   op / store / reduce   |   inverse merge  X' -> Z' (scaled by 1/n_fft)
// Lane l handles k = l + 32 i; k = 0 pairs DC with Nyquist (Z[N] = Z[0]); k = N/2 pairs with itself.
// Returns true when OP_PHON met a bin the fast operator cannot serve (SLOW = false only).
template <int NFFT, int SRC, int SINK, int OP, bool SLOW>
__device__ __forceinline__ bool spectral_middle(const StftArgs& a, cpx* buf, const float2* __restrict__ post,
                                                const float* tbl, float scale, int lane, long long spec_off, float& acc) {
    using P = Plan<NFFT>;
    constexpr int N = P::N;
    constexpr float kC = 2.f / (float)NFFT;      // folded into the operator on the way to the inverse
    constexpr bool TO_TIME = SINK == SINK_TIME;
    bool bad = false;
    // SELF: the k = N/2 bin, its own partner.  FIRST: the iteration that holds k = 0 (lane 0).
    auto pair = [&](int k, int ia, int ib, auto self_tag, auto first_tag, bool commit) {
        constexpr bool SELF = decltype(self_tag)::value, FIRST = decltype(first_tag)::value;
        const int kn = N - k;                       // partner bin; for k == 0 this is the Nyquist bin N
        cpx X, Yc;                                  // X = X[k], Yc = conj(X[N-k])
        float2 w = make_float2(0.f, 0.f);
        if (SRC == SRC_TIME || TO_TIME) w = post[k];               // (cos, sin) of 2 pi k / n_fft
        if (SRC == SRC_TIME) {
            const cpx za = buf[ia], zb = SELF ? za : buf[ib];
            // the forward window table holds w/2, so Z is already halved:
            //   E = za + conj(zb),  D = za - conj(zb),  O = -i D,  T = O conj(w) = (-i D) wx + (-D) wy
            const cpx E = add2(za, conj2(zb)), D = sub2(za, conj2(zb));
            const cpx T = fma2(neg2(D), bcast(w.y), mul2(rot90<-1>(D), bcast(w.x)));
            X = add2(E, T);
            Yc = sub2(E, T);
        } else {
            const float2 x = a.spec_in[spec_off + (long long)k * a.sf];
            const float2 y = SELF ? x : a.spec_in[spec_off + (long long)kn * a.sf];
            X = pk(x.x, x.y);
            Yc = pk(y.x, -y.y);
        }
        if (SINK == SINK_REDUCE) {
            const float t0 = fm_term<true>(a, tbl, paa_fm_stride(N + 1), k, cre(X), cim(X));
            if (commit) acc += t0;
            if (!SELF) acc += fm_term<true>(a, tbl, paa_fm_stride(N + 1), kn, cre(Yc), -cim(Yc));
            return;
        }
        if (SLOW && OP == OP_PHON) {
            float xr, xi, yr, yi;
            up(X, xr, xi);
            up(Yc, yr, yi);
            yi = -yi;
            op_phon(tbl, k, TO_TIME ? kC : 1.f, xr, xi);
            if (!SELF) op_phon(tbl, kn, TO_TIME ? kC : 1.f, yr, yi);
            X = pk(xr, xi);
            Yc = SELF ? conj2(X) : pk(yr, -yi);
        } else {
            X = apply_op2<OP, TO_TIME>(a, tbl, scale, kC, k, X, bad);
            if (!SELF) Yc = apply_op2<OP, TO_TIME>(a, tbl, scale, kC, kn, Yc, bad);
            else Yc = conj2(X);
        }
        if (SINK == SINK_SPEC) {
            if (commit) a.spec_out[spec_off + (long long)k * a.sf] = make_float2(cre(X), cim(X));
            if (!SELF) a.spec_out[spec_off + (long long)kn * a.sf] = make_float2(cre(Yc), -cim(Yc));
            return;
        }
        // inverse merge; irfft ignores the imaginary parts of the DC and Nyquist bins
        if (FIRST && k == 0) { X = pk(cre(X), 0.f); Yc = pk(cre(Yc), 0.f); }
        //   A = X + conj(Y),  B = X - conj(Y),  p = w B,  Z'[k] = A + i p,  Z'[N-k] = conj(A) + swap(p)
        const cpx A = add2(X, Yc), Bv = sub2(X, Yc);
        const cpx p = fma2(rot90<+1>(Bv), bcast(w.y), mul2(Bv, bcast(w.x)));
        if (commit) buf[ia] = add2(A, rot90<+1>(p));
        if (!SELF && !(FIRST && k == 0)) buf[ib] = add2(conj2(A), swap2(p));
    };
    // partner index N - k = (32 - lane) + (N - 32 (i + 1)): one per-lane base + a compile-time offset; only k = 0
    // (lane 0 of the first iteration) wraps to Z[0]
    const int pb = 32 - lane;
    if (SLOW) {                                 // rare path: keep it small (rolled, run-time indices)
#pragma unroll 1
        for (int k = lane; k < N / 2; k += 32) pair(k, k, (N - k) & (N - 1), std::false_type{}, std::true_type{}, true);
    } else {
        pair(lane, lane, lane == 0 ? 0 : pb + (N - 32), std::false_type{}, std::true_type{}, true);
#pragma unroll
        for (int i = 1; i < N / 64; ++i)
            pair(lane + 32 * i, lane + 32 * i, pb + (N - 32 * (i + 1)), std::false_type{}, std::false_type{}, true);
    }
    // the self-paired bin: every lane computes it, lane 0 commits
    pair(N / 2, N / 2, N / 2, std::true_type{}, std::false_type{}, lane == 0);
    return bad;
}

// ---- n_fft 1024: the spectral middle on registers ----------------------------------------------------------
// cos / sin (pi j / 16) as literals: after unrolling every use below is an immediate operand.
__host__ __device__ constexpr float cpi16(int j) {
    switch (j & 31) {
        case 0: return 1.f;
        case 1: return 0.98078528040323043f;
        case 2: return 0.92387953251128674f;
        case 3: return 0.83146961230254524f;
        case 4: return 0.70710678118654752f;
        case 5: return 0.55557023301960218f;
        case 6: return 0.38268343236508977f;
        case 7: return 0.19509032201612825f;
        case 8: return 0.f;
        case 9: return -0.19509032201612825f;
        case 10: return -0.38268343236508977f;
        case 11: return -0.55557023301960218f;
        case 12: return -0.70710678118654752f;
        case 13: return -0.83146961230254524f;
        case 14: return -0.92387953251128674f;
        case 15: return -0.98078528040323043f;
        case 16: return -1.f;
        default: return -cpi16(32 - (j & 31) < 16 ? 16 - (32 - (j & 31)) : 0);   // not used (j <= 16 everywhere)
    }
}
__host__ __device__ constexpr float spi16(int j) { return j <= 8 ? cpi16(8 - j) : cpi16(j - 8); }

// With the paired butterfly assignment (paa_fft.cuh) lane l != 0 holds z[0][r] = Z[k], k = l + 64 r, and its
// conjugate partner z[1][7-r] = Z[N-k]; so slot r = (z[0][r], z[1][7-r]) is a complete (k, N-k) pair and split ->
// per-bin op -> merge run on one thread's registers, in place.  The split twiddle e^{2 pi i k / n_fft} is
// base(l) * e^{i pi r / 8}: one per-lane register and immediates.  The only shared memory the middle touches is the
// operator's own per-bin table.
// Lane 0 holds the two self-paired butterflies 0 and 32 (bins 64 r and 32 + 64 r), whose pairs sit in other registers:
//   (Z[64 r], Z[64 (8-r)]) r = 1..3,   (Z[32 + 64 i], Z[32 + 64 (7-i)]) i = 0..3,   and the single bins Z[0], Z[N/2].
// A register permutation executed by lane 0 alone (predicated moves) puts the first kind into slots 1..3 and the
// second into slots 4..7 (with its own twiddle base and bin offset: 32 + 64 i = 64 r - 224), and slot 0 takes the two
// singles: za = Z[0] (DC + Nyquist), zb = Z[N/2].  So all 32 lanes run the same eight slots, no divergence.
__device__ __forceinline__ void lane0_permute_in(cpx (&z)[2][8]) {
    const cpx a4 = z[0][4], a5 = z[0][5], a6 = z[0][6], a7 = z[0][7];
    const cpx b0 = z[1][0], b1 = z[1][1], b2 = z[1][2], b3 = z[1][3], b4 = z[1][4], b5 = z[1][5], b6 = z[1][6], b7 = z[1][7];
    z[1][6] = a7; z[1][5] = a6; z[1][4] = a5;                   // slots 1..3: zb = Z[64 (8-r)]
    z[0][4] = b0; z[0][5] = b1; z[0][6] = b2; z[0][7] = b3;     // slots 4..7: za = Z[32 + 64 (r-4)]
    z[1][3] = b7; z[1][2] = b6; z[1][1] = b5; z[1][0] = b4;     //             zb = Z[32 + 64 (11-r)]
    z[1][7] = a4;                                               // slot 0: zb = Z[N/2]
}
__device__ __forceinline__ void lane0_permute_out(cpx (&z)[2][8]) {
    const cpx a4 = z[0][4], a5 = z[0][5], a6 = z[0][6], a7 = z[0][7];
    const cpx b0 = z[1][0], b1 = z[1][1], b2 = z[1][2], b3 = z[1][3], b4 = z[1][4], b5 = z[1][5], b6 = z[1][6], b7 = z[1][7];
    z[0][7] = b6; z[0][6] = b5; z[0][5] = b4;
    z[1][0] = a4; z[1][1] = a5; z[1][2] = a6; z[1][3] = a7;
    z[1][7] = b3; z[1][6] = b2; z[1][5] = b1; z[1][4] = b0;
    z[0][4] = b7;
}

template <int SRC, int SINK, int OP, bool SLOW>
__device__ __forceinline__ bool middle_paired(const StftArgs& a, cpx (&z)[2][8], cpx base, const float* tbl, float scale,
                                              int lane, long long spec_off, float& acc) {
    constexpr int N = 512;
    constexpr float kC = 2.f / 1024.f;           // folded into the operator on the way to the inverse
    constexpr bool TO_TIME = SINK == SINK_TIME;
    const bool l0 = lane == 0;
    bool bad = false;
    auto op = [&](cpx X, int k, bool& flag) {     // X or conj(X): a real gain commutes with conjugation
        if (SLOW && OP == OP_PHON) {
            float xr, xi;
            up(X, xr, xi);
            op_phon(tbl, k, TO_TIME ? kC : 1.f, xr, xi);
            return pk(xr, xi);
        }
        return apply_op2<OP, TO_TIME>(a, tbl, scale, kC, k, X, flag);
    };
    // za = Z[k], zb = Z[N-k] (both in/out); w = (cos, sin) of 2 pi k / n_fft.
    // dc (lane 0, slot 0 only): za = Z[0] yields the DC bin and the Nyquist bin N; zb is not part of the pair.
    auto pair = [&](cpx& za, cpx& zb, int k, cpx w, bool dc) {
        const int kn = N - k;
        float wx, wy;
        up(w, wx, wy);
        cpx X, Yc;                                  // X = X[k], Yc = conj(X[N-k])
        if (SRC == SRC_TIME) {
            const cpx zp = dc ? za : zb;
            // the forward window table holds w/2, so Z is already halved:
            //   E = za + conj(zb),  D = za - conj(zb),  O = -i D,  T = O conj(w) = (-i D) wx + (-D) wy
            const cpx E = add2(za, conj2(zp)), D = sub2(za, conj2(zp));
            const cpx T = fma2(neg2(D), bcast(wy), mul2(rot90<-1>(D), bcast(wx)));
            X = add2(E, T);
            Yc = sub2(E, T);
        } else {
            const float2 x = a.spec_in[spec_off + (long long)k * a.sf];
            const float2 y = a.spec_in[spec_off + (long long)kn * a.sf];
            X = pk(x.x, x.y);
            Yc = pk(y.x, -y.y);
        }
        if (SINK == SINK_REDUCE) {
            acc += fm_term<true>(a, tbl, paa_fm_stride(N + 1), k, cre(X), cim(X));
            acc += fm_term<true>(a, tbl, paa_fm_stride(N + 1), kn, cre(Yc), -cim(Yc));
            return;
        }
        X = op(X, k, bad);
        Yc = op(Yc, kn, bad);
        if (SINK == SINK_SPEC) {
            a.spec_out[spec_off + (long long)k * a.sf] = make_float2(cre(X), cim(X));
            a.spec_out[spec_off + (long long)kn * a.sf] = make_float2(cre(Yc), -cim(Yc));
            return;
        }
        // inverse merge; irfft ignores the imaginary parts of the DC and Nyquist bins
        X = pk(cre(X), dc ? 0.f : cim(X));
        Yc = pk(cre(Yc), dc ? 0.f : cim(Yc));
        //   A = X + conj(Y),  B = X - conj(Y),  p = w B,  Z'[k] = A + i p,  Z'[N-k] = conj(A) + swap(p)
        const cpx A = add2(X, Yc), Bv = sub2(X, Yc);
        const cpx p = fma2(rot90<+1>(Bv), bcast(wy), mul2(Bv, bcast(wx)));
        za = add2(A, rot90<+1>(p));
        zb = add2(conj2(A), swap2(p));
    };
    if (SRC == SRC_TIME && l0) lane0_permute_in(z);
    // slots 4..7 of lane 0 hold bins 32 + 64 (r-4) = 64 r - 224: own bin offset and twiddle base e^{-2 pi i 224 / 1024}
    const int kB = l0 ? -224 : lane;
    const cpx baseB = l0 ? pk(cpi16(7), -spi16(7)) : base;
    // slot 0: a normal pair for lanes 1..31; lane 0 runs the DC/Nyquist unit on za and the N/2 unit on zb
    {
        const cpx half_in = z[1][7];
        pair(z[0][0], z[1][7], lane, base, l0);
        // the self-paired bin k = N/2 (lane 0 commits):  X = 2 conj(z),  Z' = 2 conj(X')
        cpx Xh;
        if (SRC == SRC_TIME) Xh = mul2(conj2(half_in), bcast(2.f));
        else { const float2 x = a.spec_in[spec_off + (long long)(l0 ? N / 2 : lane) * a.sf]; Xh = pk(x.x, x.y); }
        if (SINK == SINK_REDUCE) {
            const float t = fm_term<true>(a, tbl, paa_fm_stride(N + 1), N / 2, cre(Xh), cim(Xh));
            if (l0) acc += t;
        } else {
            bool bad_h = false;
            Xh = op(Xh, N / 2, bad_h);
            bad = bad || (l0 && bad_h);
            if (SINK == SINK_SPEC) { if (l0) a.spec_out[spec_off + (long long)(N / 2) * a.sf] = make_float2(cre(Xh), cim(Xh)); }
            else if (l0) z[1][7] = mul2(conj2(Xh), bcast(2.f));
        }
    }
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        // min_max_freqs: a slot whose bins are masked in EVERY lane (kernel-uniform, worked out by the host: with the
        // reference's default band 120 Hz .. 20 kHz only bins 0..7 survive, i.e. slots 1..7 are dead) contributes zeros
        // to the inverse transform: no split, no merge.  (At least one slot is live, so a NaN / Inf frame still
        // reaches every output sample through it, as torch's NaN * 0 does.)
        if (OP == OP_MASK && SRC == SRC_TIME && SINK == SINK_TIME && ((a.dead_slots >> r) & 1)) {
            z[0][r] = 0;
            z[1][7 - r] = 0;
            continue;
        }
        // w = base * e^{i pi r / 8}
        const cpx bs = r < 4 ? base : baseB;
        const cpx w = fma2(rot90<+1>(bs), bcast(spi16(2 * r)), mul2(bs, bcast(cpi16(2 * r))));
        pair(z[0][r], z[1][7 - r], (r < 4 ? lane : kB) + 64 * r, w, false);
    }
    if (SINK == SINK_TIME && l0) lane0_permute_out(z);
    return bad;
}

// ---- the kernel -----------------------------------------------------------------------------------
template <int NFFT, int SRC, int SINK, int OP>
__global__ void __launch_bounds__(kThreadsStft, (NFFT == 1024 ? 2 : 3)) k_stft(StftArgs a) {
    using P = Plan<NFFT>;
    using L = TwLayout<NFFT>;
    constexpr int N = P::N;
    constexpr int F = N + 1;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ float red[kWarps];
    __shared__ int s_done[kWarps];             // overlap-add phases each warp has completed (SINK_TIME)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hop = a.hop, R = a.R, FT = a.frames_per_tile;
    const int S = a.blocks_per_tile;
    const int lin = (FT - 1) * hop + NFFT;                 // staged input span

    // shared-memory carve-up (all offsets multiples of 16 bytes)
    unsigned char* sp = smem;
    const float4* s_tw = (const float4*)(sp + a.off_tw);
    const float2* s_post = (const float2*)(sp + a.off_post);    // n_fft 512 only
    // n_fft 1024 derives the split twiddles in registers (middle_paired): only the FFT twiddles stay resident
    const unsigned tbl_bytes = (NFFT == 1024) ? a.off_post : a.off_win;
    sp += tbl_bytes;                           // the window itself only visits shared memory (see below)
    float* s_thr = (float*)sp;                 // per-bin table of the operator: phon limits, or the fletcher_munson blob
    if (OP == OP_PHON || OP == OP_PHON_DB) sp += ((F * 4 + 15) / 16) * 16;
    if (SINK == SINK_REDUCE) sp += a.fm_blob_bytes;
    float* s_renv = (float*)sp;                // reciprocal window envelope of an interior hop-block
    if (SINK == SINK_TIME) sp += ((hop * 4 + 15) / 16) * 16;
    float2* s_fft = (float2*)sp;
    sp += (size_t)kWarps * BufLayout<NFFT>::kFloat2 * sizeof(float2);
    const unsigned fft_bytes = (unsigned)(kWarps * BufLayout<NFFT>::kFloat2 * sizeof(float2));
    float* s_in = (float*)sp;
    if (SRC == SRC_TIME) sp += (size_t)lin * 4;
    float* s_ola = (float*)sp;                 // overlap-add accumulator (SINK_TIME)

    const int row = blockIdx.x / a.tiles_per_row, ti = blockIdx.x % a.tiles_per_row;
    // first frame of the tile, and the sample index (unpadded coordinates) of s_in[0]
    const int t0 = (SINK == SINK_TIME) ? ti * S - R / 2 + 1 : ti * FT;
    const int in0 = t0 * hop - NFFT / 2;

    const float* xr = (SRC == SRC_TIME) ? a.x + (size_t)row * a.T : nullptr;
    const float* gr = (SRC == SRC_TIME && a.grad) ? a.grad + (size_t)row * a.T : nullptr;
    float* qr = (SRC == SRC_TIME && SINK == SINK_REDUCE && a.q_out) ? a.q_out + (size_t)row * a.T : nullptr;
    const int own_lo = NFFT / 2, own_hi = NFFT / 2 + FT * hop;   // SINK_REDUCE: samples this tile stores to q_out
    // Interior tiles (no reflect padding, 16-byte aligned rows) are staged by two TMA bulk copies -- the perturbation
    // span straight into s_in, the gradient span into the still idle FFT buffers -- issued by one thread before
    // anything else; a shared-memory pass then applies the PGD step in place.  Row-end tiles take the per-thread path.
    const bool tma_stage = SRC == SRC_TIME && a.vec_ok && in0 >= 0 && in0 + lin <= a.T && (!gr || (unsigned)lin * 4u <= fft_bytes);

    if (tid == 0) mbar_init(&bar, 1);
    if (tid < kWarps) s_done[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        // the FFT twiddles stay resident; the window never enters shared memory (every lane keeps its 16 values in
        // registers, read once from the L2-resident table while the bulk copies are in flight)
        mbar_expect_tx(&bar, tbl_bytes + (SINK == SINK_REDUCE ? a.fm_blob_bytes : 0u) +
                                 (tma_stage ? (unsigned)lin * 4u * (gr ? 2u : 1u) : 0u));
        tma_bulk_g2s(smem, a.blob, tbl_bytes, &bar);
        if (SINK == SINK_REDUCE) tma_bulk_g2s(s_thr, a.fm_blob, a.fm_blob_bytes, &bar);
        if (tma_stage) {
            tma_bulk_g2s(s_in, xr + in0, (unsigned)lin * 4u, &bar);
            if (gr) tma_bulk_g2s(s_fft, gr + in0, (unsigned)lin * 4u, &bar);
        }
    }

    // ---- stage the input span (reflect padding at the row ends, PGD step fused) ----------------
    if (SRC == SRC_TIME && !tma_stage) {
        const int T = a.T;
        // All global loads of a chunk are issued before any of them is consumed (kStageUnroll x 2 float4 in flight
        // per thread): the staging latency is otherwise exposed once per loop trip.
        constexpr int kStageUnroll = (NFFT == 1024) ? 9 : 5;      // 1024/256: 2240 float4 per tile = 8.75 per thread -> one trip
        const int lin4 = lin / 4;
        for (int base4 = 0; base4 < lin4; base4 += kStageUnroll * kThreadsStft) {
            float4 pv[kStageUnroll], gv[kStageUnroll];
            bool fast[kStageUnroll];
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int i4 = base4 + u * kThreadsStft + tid, s = in0 + i4 * 4;
                fast[u] = i4 < lin4 && a.vec_ok && s >= 0 && s + 3 < T;
                pv[u] = gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fast[u]) {
                    pv[u] = *reinterpret_cast<const float4*>(xr + s);
                    if (gr) gv[u] = *reinterpret_cast<const float4*>(gr + s);
                }
            }
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int i4 = base4 + u * kThreadsStft + tid;
                if (i4 >= lin4) continue;
                const int i = i4 * 4, s = in0 + i;
                float4 v = pv[u];
                if (fast[u]) {
                    if (gr) {
                        const float4 g = gv[u];
                        v.x += a.lr * ((float)(g.x > 0.f) - (float)(g.x < 0.f));
                        v.y += a.lr * ((float)(g.y > 0.f) - (float)(g.y < 0.f));
                        v.z += a.lr * ((float)(g.z > 0.f) - (float)(g.z < 0.f));
                        v.w += a.lr * ((float)(g.w > 0.f) - (float)(g.w < 0.f));
                    }
                    if (qr && i >= own_lo && i < own_hi) *reinterpret_cast<float4*>(qr + s) = v;
                } else {
                    float e[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        int sc = s + c;
                        const bool inside = sc >= 0 && sc < T;
                        if (sc < 0) sc = -sc;
                        if (sc >= T) sc = 2 * (T - 1) - sc;
                        sc = min(max(sc, 0), T - 1);
                        float val = xr[sc];
                        if (gr) { const float g = gr[sc]; val += a.lr * ((float)(g > 0.f) - (float)(g < 0.f)); }
                        if (qr && inside && i + c >= own_lo && i + c < own_hi) qr[sc] = val;
                        e[c] = val;
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
                *reinterpret_cast<float4*>(s_in + i) = v;
            }
        }
    }
    // this lane's slice of the (halved) window lives in registers for the whole tile: the first forward stage and the
    // last inverse stage touch the same points m = lane + 32*j, j = 0 .. N/32-1 (saves 64 smem wavefronts per frame)
    cpx wreg[N / 32];
    {
        const float* g_w = reinterpret_cast<const float*>((const unsigned char*)a.blob + a.off_win);
        const float2* g_w2 = reinterpret_cast<const float2*>(g_w);
#pragma unroll
        for (int j = 0; j < N / 32; ++j) { const float2 w = __ldg(g_w2 + lane + 32 * j); wreg[j] = pk(w.x, w.y); }
        if (SINK == SINK_TIME) {
            for (int q = tid; q < hop; q += kThreadsStft) {
                float e = 0.f;
                for (int d = R - 1; d >= 0; --d) { const float wv = 2.f * __ldg(g_w + d * hop + q); e += wv * wv; }
                s_renv[q] = 1.f / e;
            }
            for (int i4 = tid; i4 < S * hop / 4; i4 += kThreadsStft) *reinterpret_cast<float4*>(s_ola + i4 * 4) = make_float4(0, 0, 0, 0);
        }
    }
    // The tile this slot's next CTA will stage: pull it from HBM into L2 now, so that its (latency-bound) prologue
    // finds the lines there.  Reflect-padding samples at the row ends are left to the demand loads.
    if (SRC == SRC_TIME && a.pf_stride > 0) {
        const int nb = blockIdx.x + a.pf_stride;
        if (nb < (int)gridDim.x) {
            const int nrow = nb / a.tiles_per_row, nti = nb % a.tiles_per_row;
            const int nt0 = (SINK == SINK_TIME) ? nti * S - R / 2 + 1 : nti * FT;
            const int lo = max(nt0 * hop - NFFT / 2, 0), hi = min(nt0 * hop - NFFT / 2 + lin, a.T);
            const float* px = a.x + (size_t)nrow * a.T;
            const float* pg = a.grad ? a.grad + (size_t)nrow * a.T : nullptr;
            for (int s = lo + tid * 32; s < hi; s += kThreadsStft * 32) {          // one 128-byte line per thread and trip
                asm volatile("prefetch.global.L2 [%0];" ::"l"(px + s));
                if (pg) asm volatile("prefetch.global.L2 [%0];" ::"l"(pg + s));
            }
        }
    }

    // ---- max_phon: scaled threshold  thr[k] = spl_thresh[k] - max(spl_thresh) + reference_db ------
    float scale = 1.f;
    if (OP == OP_SCALE) scale = a.scalars[PAA_S_SCALE] * (2.f / (float)NFFT);
    if (OP == OP_PHON || OP == OP_PHON_DB) {
        float mx = -INFINITY;
        for (int k = tid; k < F; k += kThreadsStft) mx = fmaxf(mx, a.spl_thresh[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) mx = fmaxf(mx, red[w]);
        // OP_PHON keeps the limit as a magnitude, 10^(thr/20); OP_PHON_DB keeps it in dB
        for (int k = tid; k < F; k += kThreadsStft) {
            const float thr = (a.spl_thresh[k] - mx) + a.ref_db;
            s_thr[k] = OP == OP_PHON ? exp10f(thr / 20.f) * (2.f / (float)NFFT) : thr;
        }
    }
    mbar_wait(&bar, 0);
    if (tma_stage && gr) {
        // PGD step on the staged span, in shared memory (the gradient sits in the FFT buffers)
        const float4* g4 = reinterpret_cast<const float4*>(s_fft);
        float4* p4 = reinterpret_cast<float4*>(s_in);
        const float lr = a.lr;
        for (int i4 = tid; i4 < lin / 4; i4 += kThreadsStft) {
            float4 v = p4[i4];
            const float4 g = g4[i4];
            v.x += lr * ((float)(g.x > 0.f) - (float)(g.x < 0.f));
            v.y += lr * ((float)(g.y > 0.f) - (float)(g.y < 0.f));
            v.z += lr * ((float)(g.z > 0.f) - (float)(g.z < 0.f));
            v.w += lr * ((float)(g.w > 0.f) - (float)(g.w < 0.f));
            p4[i4] = v;
            const int i = i4 * 4;
            if (qr && i >= own_lo && i < own_hi) *reinterpret_cast<float4*>(qr + in0 + i) = v;
        }
    }
    __syncthreads();

    // ---- frames ----------------------------------------------------------------------------------
    cpx* buf = reinterpret_cast<cpx*>(s_fft) + (size_t)warp * BufLayout<NFFT>::kFloat2;
    const LaneBase<NFFT> lb(lane);
#ifdef PAA_TW1_SMEM
    const Tw1Shared tw1{s_tw + lane};
#else
    float4 tw1[P::R1 / 2];                     // stage-1 twiddles of this lane (same for every butterfly and frame)
#pragma unroll
    for (int q = 0; q < P::R1 / 2; ++q) tw1[q] = s_tw[q * 32 + lane];
#endif
    float acc = 0.f;
    const int olim = S * hop;
    // n_fft 1024: this lane's split-twiddle base e^{2 pi i lane / n_fft} (middle_paired), straight from the global table
    cpx post_base = 0;
    if (NFFT == 1024) {
        const float2 pb = __ldg(reinterpret_cast<const float2*>((const unsigned char*)a.blob + a.off_post) + lane);
        post_base = pk(pb.x, pb.y);
    }
    for (int r = 0; r < R; ++r) {
        for (int q = 0; q < a.Q; ++q) {
            const int f = (warp * a.Q + q) * R + r;
            const int t = t0 + f;
            if (t < 0 || t >= a.n_frames) continue;              // warp-uniform
            const long long spec_off = (long long)row * a.sb + (long long)t * a.st;
            const int obase = (f - R + 1) * hop;                     // owned-region coordinate of frame sample 0
            cpx* ola = reinterpret_cast<cpx*>(s_ola + obase) + lane;   // only dereferenced inside [0, olim)
            const bool inside = obase >= 0 && obase + NFFT <= olim;  // warp-uniform: the whole frame lands in the owned region
            // Overlap-add ordering without block barriers.  Frames that touch the same accumulator samples are at most
            // R-1 apart, i.e. they belong to this warp (program order) or to a neighbouring warp and another phase;
            // adding them in phase order only requires that warp w+1 has finished phase r-1 before warp w adds phase r
            // (the mirror case, warp w-1 in a later phase, waits on this warp by the same rule).  So the warps run
            // skewed instead of in lock-step, and every sample still receives its R contributions in the same fixed
            // order.  The wait sits in front of the first accumulator access of the frame (c == 0).
            auto wait_neighbour = [&]() {
                if (r > 0 && warp + 1 < kWarps) {
                    if (lane == 0) {
                        int done;
                        do {
                            asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(done) : "r"(smem_u32(&s_done[warp + 1])) : "memory");
                        } while (done < r);
                    }
                    __syncwarp();
                }
            };
            // max_phon sits at the 128-register ceiling: there the wait goes in front of the whole inverse transform
            // (measured: the late wait costs it a spill and 1.4 %, and gains the other operators 3 %)
            constexpr bool kLateWait = OP != OP_PHON;
            auto ola_all = [&](int, int c, cpx v) {
                if (kLateWait && c == 0) wait_neighbour();
                ola[c] = fma2(v, wreg[c / 32], ola[c]);
            };
            auto ola_edge = [&](int m, int c, cpx v) {
                if (kLateWait && c == 0) wait_neighbour();
                const int o = obase + 2 * m;
                if (o >= 0 && o < olim) ola[c] = fma2(v, wreg[c / 32], ola[c]);
            };
            const cpx* x2 = reinterpret_cast<const cpx*>(s_in + f * hop) + lane;
            auto windowed = [&](int, int c) { return mul2(x2[c], wreg[c / 32]); };
            bool slow = false;
            if constexpr (NFFT == 1024) {
                // spectrum in registers between the transforms (paired butterflies, middle_paired)
                cpx z[2][8];
                for (;;) {
                    if (SRC == SRC_TIME) fft_forward_paired(buf, s_tw, tw1, lane, lb, windowed, z);
                    bool bad;
                    if (OP == OP_PHON && slow) bad = middle_paired<SRC, SINK, OP, true>(a, z, post_base, s_thr, scale, lane, spec_off, acc);
                    else bad = middle_paired<SRC, SINK, OP, false>(a, z, post_base, s_thr, scale, lane, spec_off, acc);
                    // a zero / denormal / NaN bin somewhere in the frame (rare): redo the frame with the exact operator
                    if (OP != OP_PHON || SINK == SINK_REDUCE || slow || !__any_sync(0xffffffffu, bad)) break;
                    slow = true;
                }
                if (SINK == SINK_TIME) {
                    if (!kLateWait) wait_neighbour();
                    if (inside) fft_inverse_paired(buf, s_tw, tw1, lane, lb, z, ola_all);
                    else fft_inverse_paired(buf, s_tw, tw1, lane, lb, z, ola_edge);
                }
            } else {
                // n_fft 512: the spectrum sits in `buf` in natural un-padded order between the transforms (m = lane + c)
                for (;;) {
                    if (SRC == SRC_TIME) {
                        fft_warp<NFFT, -1>(buf, s_tw, tw1, lane, lb, windowed, [&](int, int c, cpx v) { buf[lane + c] = v; });
                        __syncwarp();
                    }
                    bool bad;
                    if (OP == OP_PHON && slow)
                        bad = spectral_middle<NFFT, SRC, SINK, OP, true>(a, buf, s_post, s_thr, scale, lane, spec_off, acc);
                    else
                        bad = spectral_middle<NFFT, SRC, SINK, OP, false>(a, buf, s_post, s_thr, scale, lane, spec_off, acc);
                    if (OP != OP_PHON || SINK == SINK_REDUCE || slow || !__any_sync(0xffffffffu, bad)) break;
                    slow = true;
                    __syncwarp();
                }
                if (SINK == SINK_TIME) {
                    if (!kLateWait) wait_neighbour();
                    __syncwarp();
                    auto from_buf = [&](int, int c) { return buf[lane + c]; };
                    if (inside) fft_warp<NFFT, +1>(buf, s_tw, tw1, lane, lb, from_buf, ola_all);
                    else fft_warp<NFFT, +1>(buf, s_tw, tw1, lane, lb, from_buf, ola_edge);
                }
            }
        }
        if (SINK == SINK_TIME) {                                 // publish: this warp's phase-r contributions are in place
            __syncwarp();
            if (lane == 0) asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&s_done[warp])), "r"(r + 1) : "memory");
        }
    }
    if (SINK == SINK_TIME) __syncthreads();                      // the epilogue reads every warp's samples

    // ---- epilogue -----------------------------------------------------------------------------------
    if (SINK == SINK_TIME) {
        // y[n] = ola[n] / sum_t w^2[n + n/2 - t*hop]  (torch.istft's window envelope), zero past hop*(T'-1).
        // Interior hop-blocks share the reciprocal envelope table built at start-up; the R-1 blocks at either end
        // of a row see fewer frames and take the slow path (window from global memory).
        const float* g_win = reinterpret_cast<const float*>((const unsigned char*)a.blob + a.off_win);
        float* yr = a.y + (size_t)row * a.out_len;
        const bool vec = (a.out_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15u) == 0);
        const int hop4 = hop >> 2, n4 = S * hop4;
        const int hop4_shift = __ffs(hop4) - 1;                   // hop divides n_fft = 2^m, so it is a power of two
        constexpr int kEpi = 4;
        for (int e0 = tid; e0 < n4; e0 += kEpi * kThreadsStft) {
            float4 o[kEpi], rv[kEpi];
            int nn[kEpi];
            bool live[kEpi];
#pragma unroll
            for (int u = 0; u < kEpi; ++u) {
                const int e = e0 + u * kThreadsStft;
                const int jb = e >> hop4_shift, q = (e - (jb << hop4_shift)) * 4;
                const int gb = ti * S + jb;
                nn[u] = gb * hop + q;
                live[u] = e < n4 && nn[u] < a.out_len;
                o[u] = rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!live[u]) continue;
                const int ub = gb + R / 2;                                    // newest frame covering this block
                const bool exists = gb < a.n_frames - 1;
                const bool interior = ub - R + 1 >= 0 && ub <= a.n_frames - 1;
                if (!exists) continue;
                o[u] = *reinterpret_cast<const float4*>(s_ola + jb * hop + q);
                if (interior) {
                    rv[u] = *reinterpret_cast<const float4*>(s_renv + q);
                } else {
                    float ev[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int d = R - 1; d >= 0; --d) {
                        const int t = ub - d;
                        if (t < 0 || t >= a.n_frames) continue;
#pragma unroll
                        for (int c = 0; c < 4; ++c) { const float wv = 2.f * __ldg(g_win + d * hop + q + c); ev[c] += wv * wv; }
                    }
                    rv[u] = make_float4(1.f / ev[0], 1.f / ev[1], 1.f / ev[2], 1.f / ev[3]);
                }
            }
#pragma unroll
            for (int u = 0; u < kEpi; ++u) {
                if (!live[u]) continue;
                const float4 v = make_float4(o[u].x * rv[u].x, o[u].y * rv[u].y, o[u].z * rv[u].z, o[u].w * rv[u].w);
                const int n = nn[u];
                if (vec && n + 3 < a.out_len) {
                    *reinterpret_cast<float4*>(yr + n) = v;
                } else {
                    if (n < a.out_len) yr[n] = v.x;
                    if (n + 1 < a.out_len) yr[n + 1] = v.y;
                    if (n + 2 < a.out_len) yr[n + 2] = v.z;
                    if (n + 3 < a.out_len) yr[n + 3] = v.w;
                }
            }
        }
    } else if (SINK == SINK_REDUCE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < kWarps; ++w) s += (double)red[w];
            a.partials[blockIdx.x] = s;                  // stride 1: scratch_fm_partials, one slot per tile
        }
    }
}

// ---- n_fft 1024, hop 256: the fused kernel on half warps (paa_fft32.cuh) --------------------------------------------
// Same tile geometry, staging, overlap-add order and epilogue as k_stft<1024, SRC_TIME, SINK_TIME, OP>, but a frame
// belongs to 16 lanes with 32 points each: a CTA is 4 warps = 8 half warps, half warp h runs the frames 4h .. 4h+3 of
// the tile one after the other, and a warp instruction serves two frames that are four apart (they never touch the same
// output samples).  Conventions that differ from k_stft:
//   * Z is NOT halved by the window table (the table holds the true Hann window, which lets w[n + 512] = 1 - w[n] ride
//     the first butterfly level: only 16 window values per lane and direction, read from shared memory where both half
//     warps hit the same address).  The split therefore yields X2 = 2 X, and the operators fold 1/2048 instead of 2/1024.
//   * blob = [W_512^{l k1} as float2[32][16] | first half of the window float[512] | reciprocal envelope float[256]].
constexpr int kHwWarps = 4, kHwThreads = kHwWarps * 32, kHwHalf = 2 * kHwWarps;

// X2 = 2 X[k] -> X'[k] / 1024 * 2  (what the merge wants: see the scaling note above; tbl = limit[k] / 1024)
template <int OP>
__device__ __forceinline__ cpx apply_op_hw(const StftArgs& a, const float* tbl, float scale, int k, cpx X2, bool& bad) {
    constexpr float kC = 1.f / 2048.f;
    if (OP == OP_MASK) return mul2(X2, bcast((k < a.k_lo || k >= a.k_hi) ? kC : 0.f));
    if (OP == OP_SCALE) return mul2(X2, bcast(scale));              // caller pre-multiplies scale by 1/2048
    if (OP == OP_PHON) {
        float re, im;
        up(X2, re, im);
        const float P = fmaf(re, re, im * im);                      // 4 |X|^2
        bad = bad || !(P > 1e-36f);
        const float r = rsqrt_ftz(P);                               // 1 / (2 |X|)
        const float xc = fmaf(P * r, kC, 1e-8f * 2.f * kC);         // (|X| + 1e-8) / 1024
        return mul2(X2, bcast(fminf(xc, tbl[k]) * r));              // unit phasor X2 r; a NaN bin stays NaN through r
    }
    return mul2(X2, bcast(kC));
}

// lane 0 of a half warp holds the two self-paired butterflies k1 = 0 and 16: map their pairs onto the sixteen slots
// (za[s], zb[15-s]) every other lane uses (tools/emulate_fft32.py: lane0_in / lane0_out)
__device__ __forceinline__ void hw_lane0_in(cpx (&za)[16], cpx (&zb)[16]) {
    cpx a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = za[i]; b[i] = zb[i]; }
    zb[15] = a[8];
#pragma unroll
    for (int s = 1; s < 8; ++s) zb[15 - s] = a[16 - s];
#pragma unroll
    for (int i = 0; i < 8; ++i) { za[8 + i] = b[i]; zb[7 - i] = b[15 - i]; }
}
__device__ __forceinline__ void hw_lane0_out(cpx (&za)[16], cpx (&zb)[16]) {
    cpx a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = za[i]; b[i] = zb[i]; }
    za[8] = b[15];
#pragma unroll
    for (int s = 1; s < 8; ++s) za[16 - s] = b[15 - s];
#pragma unroll
    for (int i = 0; i < 8; ++i) { zb[i] = a[8 + i]; zb[15 - i] = b[7 - i]; }
}

// split -> per-bin operator -> merge on one thread's registers.  za[k2] = Z[lam + 32 k2], zb[k2] = Z[(32 - lam) + 32 k2]
// (lane 0: Z[32 k2] and Z[16 + 32 k2]); slot s = (za[s], zb[15-s]) is the pair (k, N-k), k = lam + 32 s, split twiddle
// e^{2 pi i k / 1024} = base(lam) e^{i pi s / 16}.  Returns true when OP_PHON met a bin the fast operator cannot serve.
template <int OP, bool SLOW>
__device__ __forceinline__ bool middle_hw(const StftArgs& a, cpx (&za)[16], cpx (&zb)[16], cpx base, const float* tbl, float scale,
                                          int lam) {
    constexpr int N = 512;
    const bool l0 = lam == 0;
    bool bad = false;
    auto op = [&](cpx X2, int k, bool& flag) {
        if (SLOW && OP == OP_PHON) {
            float xr, xi;
            up(X2, xr, xi);
            xr *= 0.5f; xi *= 0.5f;                                  // the true X[k]
            op_phon(tbl, k, 1.f / 1024.f, xr, xi);
            return pk(xr, xi);
        }
        return apply_op_hw<OP>(a, tbl, scale, k, X2, flag);
    };
    auto pair = [&](cpx& zA, cpx& zB, int k, cpx w, bool dc) {
        const int kn = N - k;
        float wx, wy;
        up(w, wx, wy);
        const cpx zp = dc ? zA : zB;
        //   E = za + conj(zb),  D = za - conj(zb),  T = (-i D) conj(w) = (-i D) wx + (-D) wy,  2 X[k] = E + T,  2 conj X[N-k] = E - T
        const cpx E = add2(zA, conj2(zp)), D = sub2(zA, conj2(zp));
        const cpx T = fma2(neg2(D), bcast(wy), mul2(rot90<-1>(D), bcast(wx)));
        cpx X = op(add2(E, T), k, bad);
        cpx Yc = op(sub2(E, T), kn, bad);
        // inverse merge; irfft ignores the imaginary parts of the DC and Nyquist bins
        X = pk(cre(X), dc ? 0.f : cim(X));
        Yc = pk(cre(Yc), dc ? 0.f : cim(Yc));
        //   A = X + conj(Y),  B = X - conj(Y),  p = w B,  Z'[k] = A + i p,  Z'[N-k] = conj(A) + swap(p)
        const cpx A = add2(X, Yc), Bv = sub2(X, Yc);
        const cpx p = fma2(rot90<+1>(Bv), bcast(wy), mul2(Bv, bcast(wx)));
        zA = add2(A, rot90<+1>(p));
        zB = add2(conj2(A), swap2(p));
    };
    if (l0) hw_lane0_in(za, zb);
    // slots 8..15 of lane 0 hold bins 16 + 32 (s - 8) = 32 s - 240: own bin offset and twiddle base e^{-2 pi i 240 / 1024}
    const int kB = l0 ? -240 : lam;
    const cpx baseB = l0 ? pk(0.09801714032956060f, -0.99518472667219689f) : base;
    {   // slot 0: a normal pair for lanes 1..15; lane 0 runs the DC/Nyquist unit on za and the N/2 unit on zb
        const cpx half_in = zb[15];
        pair(za[0], zb[15], lam, base, l0);
        bool bad_h = false;
        const cpx Xh = op(mul2(conj2(half_in), bcast(2.f)), N / 2, bad_h);       // 2 X[N/2] = 2 conj Z[N/2]
        bad = bad || (l0 && bad_h);
        if (l0) zb[15] = mul2(conj2(Xh), bcast(2.f));
    }
#pragma unroll
    for (int s = 1; s < 16; ++s) {
        const cpx bs = s < 8 ? base : baseB;
        const cpx w = fma2(rot90<+1>(bs), bcast(s32(s)), mul2(bs, bcast(c32(s))));          // base * e^{i pi s / 16}
        pair(za[s], zb[15 - s], (s < 8 ? lam : kB) + 32 * s, w, false);
    }
    if (l0) hw_lane0_out(za, zb);
    return bad;
}

template <int OP>
__global__ void __launch_bounds__(kHwThreads, 2) k_stft_hw(StftArgs a) {
    constexpr int NFFT = 1024, N = 512, F = N + 1, HOP = 256, R = 4, FT = kHwHalf * R, S = FT - R + 1;
    constexpr int LIN = (FT - 1) * HOP + NFFT;                  // staged input span (floats)
    static_assert(LIN * 4 == kHwHalf * kHwBuf * 8, "the gradient span is staged in the exchange buffers");
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ float red[kHwWarps];
    __shared__ int s_done[kHwWarps];            // overlap-add phases each warp has completed

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, hl = lane & 15, hw = tid >> 4;

    unsigned char* sp = smem;
    const cpx* s_tw = (const cpx*)sp;                          // W_512^{l k1}, [k1][l]
    const cpx* s_win = (const cpx*)(sp + 4096);                // window pairs (w[2m], w[2m+1]), m = 0..255
    const float* s_renv = (const float*)(sp + 4096 + 2048);    // reciprocal window envelope of an interior hop block
    sp += a.blob_bytes;
    float* s_thr = (float*)sp;
    if (OP == OP_PHON) sp += ((F * 4 + 15) / 16) * 16;
    cpx* s_fft = (cpx*)sp;
    sp += (size_t)kHwHalf * kHwBuf * 8;
    float* s_in = (float*)sp;
    sp += (size_t)LIN * 4;
    float* s_ola = (float*)sp;

    const int row = blockIdx.x / a.tiles_per_row, ti = blockIdx.x % a.tiles_per_row;
    const int t0 = ti * S - R / 2 + 1;
    const int in0 = t0 * HOP - NFFT / 2;
    const float* xr = a.x + (size_t)row * a.T;
    const float* gr = a.grad ? a.grad + (size_t)row * a.T : nullptr;
    const bool tma_stage = a.vec_ok && in0 >= 0 && in0 + LIN <= a.T;

    if (tid == 0) mbar_init(&bar, 1);
    if (tid < kHwWarps) s_done[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, a.blob_bytes + (tma_stage ? (unsigned)LIN * 4u * (gr ? 2u : 1u) : 0u));
        tma_bulk_g2s(smem, a.blob, a.blob_bytes, &bar);
        if (tma_stage) {
            tma_bulk_g2s(s_in, xr + in0, (unsigned)LIN * 4u, &bar);
            if (gr) tma_bulk_g2s(s_fft, gr + in0, (unsigned)LIN * 4u, &bar);
        }
    }
    if (!tma_stage) {           // row-end tiles: reflect padding, PGD step fused (as in k_stft)
        const int T = a.T;
        constexpr int kStageUnroll = 6;
        constexpr int lin4 = LIN / 4;
        for (int base4 = 0; base4 < lin4; base4 += kStageUnroll * kHwThreads) {
            float4 pv[kStageUnroll], gv[kStageUnroll];
            bool fast[kStageUnroll];
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int i4 = base4 + u * kHwThreads + tid, s = in0 + i4 * 4;
                fast[u] = i4 < lin4 && a.vec_ok && s >= 0 && s + 3 < T;
                pv[u] = gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fast[u]) {
                    pv[u] = *reinterpret_cast<const float4*>(xr + s);
                    if (gr) gv[u] = *reinterpret_cast<const float4*>(gr + s);
                }
            }
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int i4 = base4 + u * kHwThreads + tid;
                if (i4 >= lin4) continue;
                const int i = i4 * 4, s = in0 + i;
                float4 v = pv[u];
                if (fast[u]) {
                    if (gr) {
                        const float4 g = gv[u];
                        v.x += a.lr * ((float)(g.x > 0.f) - (float)(g.x < 0.f));
                        v.y += a.lr * ((float)(g.y > 0.f) - (float)(g.y < 0.f));
                        v.z += a.lr * ((float)(g.z > 0.f) - (float)(g.z < 0.f));
                        v.w += a.lr * ((float)(g.w > 0.f) - (float)(g.w < 0.f));
                    }
                } else {
                    float e[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        int sc = s + c;
                        if (sc < 0) sc = -sc;
                        if (sc >= T) sc = 2 * (T - 1) - sc;
                        sc = min(max(sc, 0), T - 1);
                        float val = xr[sc];
                        if (gr) { const float g = gr[sc]; val += a.lr * ((float)(g > 0.f) - (float)(g < 0.f)); }
                        e[c] = val;
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
                *reinterpret_cast<float4*>(s_in + i) = v;
            }
        }
    }
    for (int i4 = tid; i4 < S * HOP / 4; i4 += kHwThreads) *reinterpret_cast<float4*>(s_ola + i4 * 4) = make_float4(0, 0, 0, 0);
    if (a.pf_stride > 0) {      // the tile this slot's next CTA will stage: HBM -> L2 now
        const int nb = blockIdx.x + a.pf_stride;
        if (nb < (int)gridDim.x) {
            const int nrow = nb / a.tiles_per_row, nti = nb % a.tiles_per_row;
            const int nin0 = (nti * S - R / 2 + 1) * HOP - NFFT / 2;
            const int lo = max(nin0, 0), hi = min(nin0 + LIN, a.T);
            const float* px = a.x + (size_t)nrow * a.T;
            const float* pg = a.grad ? a.grad + (size_t)nrow * a.T : nullptr;
            for (int s = lo + tid * 32; s < hi; s += kHwThreads * 32) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(px + s));
                if (pg) asm volatile("prefetch.global.L2 [%0];" ::"l"(pg + s));
            }
        }
    }
    float scale = 1.f;
    if (OP == OP_SCALE) scale = a.scalars[PAA_S_SCALE] * (1.f / 2048.f);
    if (OP == OP_PHON) {        // thr[k] = spl_thresh[k] - max(spl_thresh) + reference_db, kept as a magnitude / 1024
        float mx = -INFINITY;
        for (int k = tid; k < F; k += kHwThreads) mx = fmaxf(mx, a.spl_thresh[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
#pragma unroll
        for (int w = 1; w < kHwWarps; ++w) mx = fmaxf(mx, red[w]);
        for (int k = tid; k < F; k += kHwThreads) {
            const float thr = (a.spl_thresh[k] - mx) + a.ref_db;
            s_thr[k] = exp10f(thr / 20.f) * (1.f / 1024.f);
        }
    }
    // this lane's split-twiddle base e^{2 pi i lam / 1024}, straight from the global table of the handle
    const float2 pbf = __ldg(reinterpret_cast<const float2*>(a.post_table) + hl);
    const cpx post_base = pk(pbf.x, pbf.y);
    mbar_wait(&bar, 0);
    if (tma_stage && gr) {      // PGD step on the staged span, in shared memory (the gradient sits in the exchange buffers)
        const float4* g4 = reinterpret_cast<const float4*>(s_fft);
        float4* p4 = reinterpret_cast<float4*>(s_in);
        const float lr = a.lr;
        for (int i4 = tid; i4 < LIN / 4; i4 += kHwThreads) {
            float4 v = p4[i4];
            const float4 g = g4[i4];
            v.x += lr * ((float)(g.x > 0.f) - (float)(g.x < 0.f));
            v.y += lr * ((float)(g.y > 0.f) - (float)(g.y < 0.f));
            v.z += lr * ((float)(g.z > 0.f) - (float)(g.z < 0.f));
            v.w += lr * ((float)(g.w > 0.f) - (float)(g.w < 0.f));
            p4[i4] = v;
        }
    }
    __syncthreads();

    // ---- frames ----------------------------------------------------------------------------------
    cpx* buf = s_fft + (size_t)hw * kHwBuf;
    cpx* bcol = buf + hl;                                   // row accesses: bcol[kHwRow * k1]
    cpx* brow0 = buf + kHwRow * hl;                         // column accesses of butterfly j0 = lam: brow0[l]
    cpx* brow1 = buf + kHwRow * hw_j1(hl);                  //                              j1
    const cpx* twl = s_tw + hl;                             // twl[16 k1] = W_512^{l k1} (forward: cos, -sin)
    const cpx* winl = s_win + hl;                           // winl[16 q], q = 0..15
    constexpr int olim = S * HOP;
    for (int r = 0; r < R; ++r) {
        const int f = hw * R + r;
        const int t = t0 + f;
        const bool live = t >= 0 && t < a.n_frames;         // per half warp
        if (__any_sync(0xffffffffu, live)) {
        const int obase = (f - R + 1) * HOP;                // owned-region coordinate of frame sample 0
        cpx* ola = reinterpret_cast<cpx*>(s_ola + obase) + hl;     // only dereferenced inside [0, olim)
        const bool inside = live && obase >= 0 && obase + NFFT <= olim;
        const cpx* x2 = reinterpret_cast<const cpx*>(s_in + f * HOP) + hl;
        cpx za[16], zb[16];
        bool slow = false;
        for (;;) {
            {   // forward: window + DFT-32 over q (the window rides the first butterfly level: w[n + 512] = 1 - w[n])
                cpx v[32];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const cpx x0 = x2[16 * b], x1 = x2[16 * (b + 8)], xh0 = x2[16 * (b + 16)], xh1 = x2[16 * (b + 24)];
                    const cpx w0 = winl[16 * b], w1 = winl[16 * (b + 8)];
                    // y_q + y_{q+16} = x_{q+16} + w (x_q - x_{q+16}),   y_q - y_{q+16} = w (x_q + x_{q+16}) - x_{q+16}
                    const cpx u0 = fma2(w0, sub2(x0, xh0), xh0), u1 = fma2(w0, add2(x0, xh0), neg2(xh0));
                    const cpx u2 = fma2(w1, sub2(x1, xh1), xh1), u3 = fma2(w1, add2(x1, xh1), neg2(xh1));
                    v[b] = add2(u0, u2);
                    v[16 + b] = sub2(u0, u2);
                    v[8 + b] = add2(u1, rot90<-1>(u3));
                    v[24 + b] = sub2(u1, rot90<-1>(u3));
                }
                dft32_tail<-1>(v);
                // column hl of the buffer belongs to this lane on either side of the forward store, and rows j0 / j1 on
                // either side of the inverse store: only the two transposes themselves need a warp barrier
#pragma unroll
                for (int k1 = 0; k1 < 32; ++k1) {
                    cpx o = v[hw_out32(k1)];
                    if (k1 > 0) { const cpx w = twl[16 * k1]; o = cmul_tw<-1>(o, cre(w), cim(w)); }
                    bcol[kHwRow * k1] = o;
                }
                __syncwarp();
#pragma unroll
                for (int l = 0; l < 16; ++l) { za[l] = brow0[l]; zb[l] = brow1[l]; }
                dft16<-1>(za);
                dft16<-1>(zb);
            }
            bool bad;
            if (OP == OP_PHON && slow) bad = middle_hw<OP, true>(a, za, zb, post_base, s_thr, scale, hl);
            else bad = middle_hw<OP, false>(a, za, zb, post_base, s_thr, scale, hl);
            // a zero / denormal / NaN bin somewhere in a live frame (rare): redo with the exact operator
            if (OP != OP_PHON || slow || !__any_sync(0xffffffffu, bad && live)) break;
            slow = true;
        }
        // overlap-add ordering: warp w adds phase r only after warp w+1 has finished phase r-1 (see k_stft)
        if (r > 0 && warp + 1 < kHwWarps) {
            if (lane == 0) {
                int done;
                do {
                    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(done) : "r"(smem_u32(&s_done[warp + 1])) : "memory");
                } while (done < r);
            }
            __syncwarp();
        }
        {   // inverse: IDFT-16 over k2 -> exchange -> conj twiddle -> IDFT-32 over k1 -> windowed overlap-add
            dft16<+1>(za);
            dft16<+1>(zb);
#pragma unroll
            for (int l = 0; l < 16; ++l) { brow0[l] = za[l]; brow1[l] = zb[l]; }
            __syncwarp();
            cpx v[32];
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) {
                cpx o = bcol[kHwRow * k1];
                if (k1 > 0) { const cpx w = twl[16 * k1]; o = cmul_tw<+1>(o, cre(w), cim(w)); }
                v[k1] = o;
            }
            dft32<+1>(v);
            const bool all_in = __all_sync(0xffffffffu, inside);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const cpx w = winl[16 * q];
                const cpx y0 = v[hw_out32(q)], y1 = v[hw_out32(q + 16)];
                const cpx c1 = fma2(neg2(w), y1, y1);               // (1 - w) y
                if (all_in) {
                    ola[16 * q] = fma2(y0, w, ola[16 * q]);
                    ola[16 * (q + 16)] = add2(ola[16 * (q + 16)], c1);
                } else if (live) {
                    const int o0 = obase + 2 * (hl + 16 * q), o1 = o0 + 512;
                    if (o0 >= 0 && o0 < olim) ola[16 * q] = fma2(y0, w, ola[16 * q]);
                    if (o1 >= 0 && o1 < olim) ola[16 * (q + 16)] = add2(ola[16 * (q + 16)], c1);
                }
            }
        }
        }
        __syncwarp();                               // publish: this warp's phase-r contributions are in place
        if (lane == 0) asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&s_done[warp])), "r"(r + 1) : "memory");
    }
    __syncthreads();

    // ---- epilogue: y[n] = ola[n] / sum_t w^2, zero past hop (T'-1), pad / crop to out_len (as k_stft) -----------------
    {
        const float* g_win = a.win_full;
        float* yr = a.y + (size_t)row * a.out_len;
        const bool vec = (a.out_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15u) == 0);
        constexpr int hop4 = HOP >> 2, n4 = S * hop4;
        constexpr int kEpi = 4;
        for (int e0 = tid; e0 < n4; e0 += kEpi * kHwThreads) {
            float4 o[kEpi], rv[kEpi];
            int nn[kEpi];
            bool livee[kEpi];
#pragma unroll
            for (int u = 0; u < kEpi; ++u) {
                const int e = e0 + u * kHwThreads;
                const int jb = e / hop4, q = (e - jb * hop4) * 4;
                const int gb = ti * S + jb;
                nn[u] = gb * HOP + q;
                livee[u] = e < n4 && nn[u] < a.out_len;
                o[u] = rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!livee[u]) continue;
                const int ub = gb + R / 2;                                    // newest frame covering this block
                const bool exists = gb < a.n_frames - 1;
                const bool interior = ub - R + 1 >= 0 && ub <= a.n_frames - 1;
                if (!exists) continue;
                o[u] = *reinterpret_cast<const float4*>(s_ola + jb * HOP + q);
                if (interior) {
                    rv[u] = *reinterpret_cast<const float4*>(s_renv + q);
                } else {
                    float ev[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int d = R - 1; d >= 0; --d) {
                        const int tt = ub - d;
                        if (tt < 0 || tt >= a.n_frames) continue;
#pragma unroll
                        for (int c = 0; c < 4; ++c) { const float wv = __ldg(g_win + d * HOP + q + c); ev[c] = fmaf(wv, wv, ev[c]); }
                    }
                    rv[u] = make_float4(1.f / ev[0], 1.f / ev[1], 1.f / ev[2], 1.f / ev[3]);
                }
            }
#pragma unroll
            for (int u = 0; u < kEpi; ++u) {
                if (!livee[u]) continue;
                const float4 v = make_float4(o[u].x * rv[u].x, o[u].y * rv[u].y, o[u].z * rv[u].z, o[u].w * rv[u].w);
                const int n = nn[u];
                if (vec && n + 3 < a.out_len) {
                    *reinterpret_cast<float4*>(yr + n) = v;
                } else {
                    if (n < a.out_len) yr[n] = v.x;
                    if (n + 1 < a.out_len) yr[n + 1] = v.y;
                    if (n + 2 < a.out_len) yr[n + 2] = v.z;
                    if (n + 3 < a.out_len) yr[n + 3] = v.w;
                }
            }
        }
    }
}

// ---- fletcher_munson finalize: norm = sqrt(sum), scale = norm <= eps ? 1 : eps / max(norm, 1e-8) ---
// projections.py:130-132.  torch's clamp(min=1e-8) propagates NaN (a NaN norm makes the whole output NaN); fmaxf
// would not, so the NaN case is kept explicit.
__device__ __forceinline__ void fm_final(double tot, float fm_eps, int apply, float& scale, float& norm) {
    norm = sqrtf((float)tot);
    scale = 1.f;
    if (apply && !(norm <= fm_eps)) {
        const float floor_n = (norm != norm) ? norm : fmaxf(norm, 1e-8f);
        scale = __frcp_rn(floor_n) * fm_eps;
    }
}
// Fixed-order sum of `n` block partials (stride `stride` doubles) by one CTA; every CTA that runs it gets the same bits.
__device__ __forceinline__ double fm_sum_partials(const double* partials, int n, int stride) {
    __shared__ double sh[32];
    __shared__ double s_tot;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[(size_t)i * stride];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) tot += sh[w];
        s_tot = tot;
    }
    __syncthreads();
    return s_tot;
}
__device__ __forceinline__ void fm_store_scalars(float* scalars, float scale, float norm, double tot) {
    scalars[PAA_S_SCALE] = scale;
    scalars[PAA_S_NORM] = norm;
    scalars[PAA_S_AUX0] = (float)tot;
    scalars[PAA_S_AUX1] = 0.f;
}
__global__ void k_fm_finalize(const double* partials, int nblocks, int stride, float* scalars, float fm_eps, int apply) {
    const double tot = fm_sum_partials(partials, nblocks, stride);
    if (threadIdx.x == 0) {
        float scale, norm;
        fm_final(tot, fm_eps, apply, scale, norm);
        fm_store_scalars(scalars, scale, norm, tot);
    }
}

// Pass B of fletcher_munson in its default form.  ISTFT(s * STFT(q)) = s * q on the samples the inverse
// reconstructs, so:  p_out[row, n] = scale * q[row, n] for n < valid, 0 up to out_len.  The kernel finalizes the norm
// itself: every CTA re-sums pass A's tile partials in the same fixed order (as k_fused does behind its grid
// barrier), so no finalize launch sits between the two passes; CTA 0 publishes the scalars.
__global__ void __launch_bounds__(256) k_fm_scale_identity(const float* __restrict__ q, float* __restrict__ out, int rows, int T,
                                                          int out_len, int valid, const double* __restrict__ partials,
                                                          int nparts, float fm_eps, float* __restrict__ scalars) {
    const double tot = fm_sum_partials(partials, nparts, 1);
    float sc, norm;
    fm_final(tot, fm_eps, 1, sc, norm);
    if (blockIdx.x == 0 && threadIdx.x == 0) fm_store_scalars(scalars, sc, norm, tot);
    const int lim = min(valid, T);
    const bool vec = (T % 4 == 0) && (out_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(q) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    const int cols4 = (out_len + 3) / 4;                       // float4 columns per row
    const long long n4 = (long long)rows * cols4;
    // (row, column) advance incrementally: no 64-bit division per element
    long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    int r = (int)(i4 / cols4), c4 = (int)(i4 - (long long)r * cols4);
    const int dr = (int)(stride / cols4), dc = (int)(stride - (long long)dr * cols4);
    for (; i4 < n4; i4 += stride) {
        const int c = c4 * 4;
        const float* src = q + (size_t)r * T + c;
        float* dst = out + (size_t)r * out_len + c;
        if (vec && c + 3 < lim) {
            float4 v = __ldcs(reinterpret_cast<const float4*>(src));
            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
            *reinterpret_cast<float4*>(dst) = v;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (c + e < out_len) dst[e] = (c + e < lim) ? src[e] * sc : 0.f;
        }
        r += dr; c4 += dc;
        if (c4 >= cols4) { c4 -= cols4; ++r; }
    }
}

// ---- element-wise spectrum kernels for the un-fused public functions -------------------------------
template <int OP>
__global__ void k_spec_op(StftArgs a, int F) {
    // OP_PHON_DB: thr[k] = spl_thresh[k] - max(spl_thresh) + reference_db, rebuilt by every CTA in shared memory
    // (F <= 513 floats): no temporary that concurrent streams would share
    __shared__ float s_thr[OP == OP_PHON_DB ? 520 : 1];
    __shared__ float s_red[32];
    if (OP == OP_PHON_DB) {
        float mx = -INFINITY;
        for (int k = threadIdx.x; k < F; k += blockDim.x) mx = fmaxf(mx, a.spl_thresh[k]);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
        __syncthreads();
        mx = s_red[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, s_red[w]);
        for (int k = threadIdx.x; k < F; k += blockDim.x) s_thr[k] = (a.spl_thresh[k] - mx) + a.ref_db;
        __syncthreads();
    }
    const long long n = (long long)a.rows * F * a.n_frames;
    const float scale = OP == OP_SCALE ? a.scalars[PAA_S_SCALE] : 1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // enumerate in the order of the smallest stride for coalescing
        int b, k, t;
        if (a.sf <= a.st) { k = (int)(i % F); t = (int)((i / F) % a.n_frames); b = (int)(i / ((long long)F * a.n_frames)); }
        else { t = (int)(i % a.n_frames); k = (int)((i / a.n_frames) % F); b = (int)(i / ((long long)F * a.n_frames)); }
        const long long off = b * a.sb + k * a.sf + t * a.st;
        float2 X = a.spec_in[off];
        apply_op<OP, false>(a, s_thr, scale, 1.f, k, X.x, X.y);
        a.spec_out[off] = X;
    }
}
__global__ void k_spec_fm_partials(StftArgs a, int F) {
    const long long n = (long long)a.rows * F * a.n_frames;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int b, k, t;
        if (a.sf <= a.st) { k = (int)(i % F); t = (int)((i / F) % a.n_frames); b = (int)(i / ((long long)F * a.n_frames)); }
        else { t = (int)(i % a.n_frames); k = (int)((i / a.n_frames) % F); b = (int)(i / ((long long)F * a.n_frames)); }
        const float2 X = a.spec_in[b * a.sb + k * a.sf + t * a.st];
        acc += fm_term<false>(a, a.fm_blob, paa_fm_stride(F), k, X.x, X.y);
    }
    __shared__ float sh[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)sh[w];
        a.partials[2 * (size_t)blockIdx.x] = s;
        a.partials[2 * (size_t)blockIdx.x + 1] = 0.0;
    }
}

// ---- host side ----------------------------------------------------------------------------------------
template <int NFFT>
size_t smem_bytes(const paa_handle* h, int src, int sink, int op, int FT, int S) {
    constexpr int N = NFFT / 2;
    size_t b = (NFFT == 1024) ? h->off_post : h->off_window;
    if (sink == SINK_TIME) b += ((h->hop * 4 + 15) / 16) * 16;
    if (op == OP_PHON || op == OP_PHON_DB) b += (((N + 1) * 4 + 15) / 16) * 16;
    if (sink == SINK_REDUCE) b += h->fm_blob_bytes;
    b += (size_t)kWarps * BufLayout<NFFT>::kFloat2 * sizeof(float2);
    if (src == SRC_TIME) b += (size_t)((FT - 1) * h->hop + NFFT) * 4;
    if (sink == SINK_TIME) b += (size_t)S * h->hop * 4;
    return b;
}

template <int NFFT, int SRC, int SINK, int OP>
int launch(paa_handle* h, StftArgs& a, int grid, cudaStream_t st) {
    size_t smem = smem_bytes<NFFT>(h, SRC, SINK, OP, a.frames_per_tile, a.blocks_per_tile);
    auto kern = k_stft<NFFT, SRC, SINK, OP>;
    // the attribute belongs to the (function, device) pair: raised once per pair (and again only if another handle
    // geometry needs more), not on every launch -- small shapes are launch-latency bound
    static int granted[64];
    const int dev = h->device & 63;
    if ((size_t)__atomic_load_n(&granted[dev], __ATOMIC_ACQUIRE) < smem) {
        PAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        __atomic_store_n(&granted[dev], (int)smem, __ATOMIC_RELEASE);
    }
    kern<<<grid, kThreadsStft, smem, st>>>(a);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

// n_fft 1024 / hop 256: the half-warp kernel (same tile geometry: 32 frames, 29 owned hop blocks)
template <int OP>
int launch_hw(paa_handle* h, StftArgs& a, int grid, cudaStream_t st) {
    a.blob = h->d_blob_hw; a.blob_bytes = (unsigned)h->blob_hw_smem;
    a.post_table = (const unsigned char*)h->d_blob + h->off_post;
    a.win_full = reinterpret_cast<const float*>((const unsigned char*)h->d_blob_hw + h->off_hw_window);
    a.pf_stride = h->num_sms * 2;
    size_t smem = h->blob_hw_smem + (OP == OP_PHON ? ((h->F * 4 + 15) / 16) * 16 : 0) + (size_t)kHwHalf * kHwBuf * 8 +
                  (size_t)((a.frames_per_tile - 1) * h->hop + 1024) * 4 + (size_t)a.blocks_per_tile * h->hop * 4;
    auto kern = k_stft_hw<OP>;
    static int granted[64];
    const int dev = h->device & 63;
    if ((size_t)__atomic_load_n(&granted[dev], __ATOMIC_ACQUIRE) < smem) {
        PAA_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        __atomic_store_n(&granted[dev], (int)smem, __ATOMIC_RELEASE);
    }
    kern<<<grid, kHwThreads, smem, st>>>(a);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

template <int SRC, int SINK, int OP>
int launch_n(paa_handle* h, StftArgs& a, int grid, cudaStream_t st) {
    if (h->n_fft == 1024) return launch<1024, SRC, SINK, OP>(h, a, grid, st);
    return launch<512, SRC, SINK, OP>(h, a, grid, st);
}

void fill_common(const paa_handle* h, StftArgs& a, int rows, int T, int n_frames) {
    a.rows = rows; a.T = T; a.n_frames = n_frames; a.hop = h->hop; a.R = h->R;
    a.Q = h->R >= 4 ? 1 : 2;
    a.pf_stride = h->num_sms * (h->n_fft == 1024 ? 2 : 3);       // CTAs per SM as in __launch_bounds__
    a.frames_per_tile = kWarps * a.R * a.Q;
    a.blocks_per_tile = a.frames_per_tile - a.R + 1;
    a.blob = h->d_blob; a.blob_bytes = (unsigned)h->blob_bytes;
    a.off_tw = (unsigned)h->off_twiddle; a.off_post = (unsigned)h->off_post; a.off_win = (unsigned)h->off_window;
    a.bin_hz = h->bin_hz;
    a.fm_blob = h->d_fm_blob; a.fm_blob_bytes = (unsigned)h->fm_blob_bytes;
    a.fm_np = h->fm_n_phon; a.fm_uniform = h->fm_uniform; a.fm_fill = h->fm_fill;
    a.fm_k0 = h->fm_k0; a.fm_klast = h->fm_klast; a.fm_inv_dk = h->fm_inv_dk;
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

// min_max_freqs: f_k = fp32(k)*fp32(bin_hz) compared in fp32 against fp32(min/max) like torch does
// (projections.py:74-76); monotone in k, so two bin indices describe the mask exactly.
void set_band(const paa_handle* h, StftArgs& a, double min_freq, double max_freq) {
    a.f_min = (float)min_freq; a.f_max = (float)max_freq;
    int lo = 0, hi = h->F;
    while (lo < h->F && (float)lo * h->bin_hz < a.f_min) ++lo;            // bins [0, lo) are below the band
    while (hi > 0 && (float)(hi - 1) * h->bin_hz > a.f_max) --hi;         // bins [hi, F) are above the band
    a.k_lo = lo; a.k_hi = hi;
    // n_fft 1024: slots of the register-resident middle (middle_paired) whose bins are all masked.  Slot r >= 1 touches
    // k = l + 64 r and 512 - k for the lanes l = 1..31, and in lane 0 the pairs (64 r, 512 - 64 r) for r = 1..3 and
    // (64 r - 224, 736 - 64 r) for r = 4..7.  Slot 0 (DC, Nyquist, N/2 and the lowest bins) always runs.
    a.dead_slots = 0;
    if (h->n_fft == 1024) {
        auto masked = [&](int k) { return k >= lo && k < hi; };
        for (int r = 1; r < 8; ++r) {
            bool dead = true;
            for (int l = 1; l < 32 && dead; ++l) dead = masked(l + 64 * r) && masked(512 - (l + 64 * r));
            const int k0 = r < 4 ? 64 * r : 64 * r - 224;
            dead = dead && masked(k0) && masked(512 - k0);
            if (dead) a.dead_slots |= 1u << r;
        }
    }
}

bool nola_check(const paa_handle* h, int n_frames);
// torch.istft refuses windows whose overlap-add envelope (after trimming n_fft/2) falls below 1e-11
bool nola_ok(const paa_handle* h, int n_frames) {
    // memoised per handle: an attack calls with the same length every step (benign race: same answer from any thread)
    if (h->nola_frames == n_frames) return h->nola_result != 0;
    const bool ok = nola_check(h, n_frames);
    h->nola_result = ok ? 1 : 0;
    h->nola_frames = n_frames;
    return ok;
}

bool nola_check(const paa_handle* h, int n_frames) {
    const int n = h->n_fft, hop = h->hop;
    const long long valid = (long long)hop * (n_frames - 1);
    auto env_at = [&](long long pos) {
        const long long u = pos + n / 2, ub = u / hop;
        const int q = (int)(u - ub * hop);
        float e = 0.f;
        for (int d = h->R - 1; d >= 0; --d) {
            long long t = ub - d;
            if (t >= 0 && t < n_frames) { float w = h->h_window[d * hop + q]; e += w * w; }
        }
        return e;
    };
    const long long edge = std::min<long long>(valid, 2LL * n);
    for (long long p = 0; p < edge; ++p)
        if (std::fabs(env_at(p)) < 1e-11f) return false;
    for (long long p = std::max<long long>(edge, valid - 2LL * n); p < valid; ++p)
        if (std::fabs(env_at(p)) < 1e-11f) return false;
    return true;
}

int check_time_shape(const paa_handle* h, int rows, int T) {
    if (rows <= 0 || T <= 0) return PAA_ERR_SHAPE;
    if (T <= h->n_fft / 2) return PAA_ERR_SHAPE;       // reflect padding needs n_fft/2 < T (torch raises too)
    return PAA_OK;
}

// The fused path shared by min_max_freqs / max_phon / fletcher_munson pass B.
template <int OP>
int run_fused(paa_handle* h, StftArgs& a, const float* src, const float* grad, float lr, float* p_out, int rows, int T,
              int out_len, cudaStream_t st) {
    const int n_frames = 1 + T / h->hop;
    fill_common(h, a, rows, T, n_frames);
    if (!nola_ok(h, n_frames)) return PAA_ERR_NOLA;
    a.x = src; a.grad = grad; a.lr = lr; a.y = p_out; a.out_len = out_len;
    a.vec_ok = aligned16(src) && (T % 4 == 0) && (!grad || aligned16(grad));
    const int out_blocks = (out_len + h->hop - 1) / h->hop;
    a.tiles_per_row = std::max(1, (out_blocks + a.blocks_per_tile - 1) / a.blocks_per_tile);
    if (h->use_hw) return launch_hw<OP>(h, a, rows * a.tiles_per_row, st);
    return launch_n<SRC_TIME, SINK_TIME, OP>(h, a, rows * a.tiles_per_row, st);
}

// Adam cannot be folded into the tile load (halo samples would be updated twice), so it runs as a
// streaming pre-pass into the scratch staging buffer; PGD is fused.
int prepare_source(paa_handle* h, const float* p_in, int rows, int T, const paa_step* step, void* scratch,
                   cudaStream_t st, const float** src, const float** grad, float* lr) {
    int mode = 0;
    StepDev sd{};
    int rc = paa_make_step(step, &mode, &sd);
    if (rc) return rc;
    *src = p_in; *grad = nullptr; *lr = 0.f;
    if (mode != PAA_STEP_NONE && sd.nparts > 1) {
        // mode U: tiles re-read halo samples, so the per-rank partial gradients are summed once into scratch
        if (!scratch) return PAA_ERR_NULL;
        float* gsum = scratch_gsum(scratch, rows, T);
        rc = paa_launch_sum_parts(h, sd, gsum, (int64_t)rows * T, st);
        if (rc) return rc;
        sd.grad = gsum; sd.nparts = 0;
    }
    if (mode == PAA_STEP_PGD) { *grad = sd.grad; *lr = sd.lr; }
    else if (mode == PAA_STEP_ADAM) {
        if (!scratch) return PAA_ERR_NULL;
        float* stage = scratch_stage(scratch);
        rc = paa_launch_adam_prepass(h, p_in, stage, (int64_t)rows * T, sd, st);
        if (rc) return rc;
        *src = stage;
    }
    return PAA_OK;
}

}  // namespace

extern "C" {

int paa_project_min_max_freqs(paa_handle* h, const float* p_in, float* p_out, int rows, int T, int out_len,
                              double min_freq, double max_freq, const paa_step* step, void* scratch, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !p_in || !p_out) return PAA_ERR_NULL;
    int rc = check_time_shape(h, rows, T);
    if (rc) return rc;
    if (out_len <= 0) return PAA_ERR_SHAPE;
    if (p_in == p_out) return PAA_ERR_ALIAS;
    cudaStream_t st = (cudaStream_t)stream;
    const float *src, *grad; float lr;
    rc = prepare_source(h, p_in, rows, T, step, scratch, st, &src, &grad, &lr);
    if (rc) return rc;
    StftArgs a{};
    set_band(h, a, min_freq, max_freq);
    return run_fused<OP_MASK>(h, a, src, grad, lr, p_out, rows, T, out_len, st);
}

int paa_project_max_phon(paa_handle* h, const float* p_in, float* p_out, int rows, int T, int out_len,
                         const float* spl_thresh_F, double phon_reference_db, const paa_step* step, void* scratch,
                         void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !p_in || !p_out || !spl_thresh_F) return PAA_ERR_NULL;
    int rc = check_time_shape(h, rows, T);
    if (rc) return rc;
    if (out_len <= 0) return PAA_ERR_SHAPE;
    if (p_in == p_out) return PAA_ERR_ALIAS;
    cudaStream_t st = (cudaStream_t)stream;
    const float *src, *grad; float lr;
    rc = prepare_source(h, p_in, rows, T, step, scratch, st, &src, &grad, &lr);
    if (rc) return rc;
    StftArgs a{};
    a.spl_thresh = spl_thresh_F; a.ref_db = (float)phon_reference_db;
    return run_fused<OP_PHON>(h, a, src, grad, lr, p_out, rows, T, out_len, st);
}

int paa_project_fletcher_munson(paa_handle* h, const float* p_in, float* p_out, int rows, int T, int out_len,
                                double fm_epsilon, int exact_roundtrip, const paa_step* step, void* scratch,
                                void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !p_in || !p_out || !scratch) return PAA_ERR_NULL;
    if (!h->d_fm_blob) return PAA_ERR_STATE;
    int rc = check_time_shape(h, rows, T);
    if (rc) return rc;
    if (out_len <= 0) return PAA_ERR_SHAPE;
    if (p_in == p_out) return PAA_ERR_ALIAS;
    cudaStream_t st = (cudaStream_t)stream;
    const float *src, *grad; float lr;
    rc = prepare_source(h, p_in, rows, T, step, scratch, st, &src, &grad, &lr);
    if (rc) return rc;
    // pass A: [PGD step ->] STFT -> weighted power partials; the stepped signal goes to the staging buffer
    const int n_frames = 1 + T / h->hop;
    if (!nola_ok(h, n_frames)) return PAA_ERR_NOLA;
    StftArgs a{};
    fill_common(h, a, rows, T, n_frames);
    a.x = src; a.grad = grad; a.lr = lr;
    a.q_out = grad ? scratch_stage(scratch) : nullptr;
    a.vec_ok = aligned16(src) && (T % 4 == 0) && (!grad || aligned16(grad));
    a.tiles_per_row = (n_frames + a.frames_per_tile - 1) / a.frames_per_tile;
    a.partials = scratch_fm_partials(scratch, rows, T);        // one slot per tile, sized by paa_scratch_bytes(rows, T)
    const int grid = rows * a.tiles_per_row;
    rc = launch_n<SRC_TIME, SINK_REDUCE, OP_NONE>(h, a, grid, st);
    if (rc) return rc;
    float* scalars = scratch_scalars(scratch);
    // pass B.  q is the stepped signal (staging buffer) or the input itself.
    const float* q = grad ? scratch_stage(scratch) : src;
    if (exact_roundtrip) {
        // the reference's literal ISTFT(scale * STFT(q)): finalize launch + a second fused transform
        k_fm_finalize<<<1, 256, 0, st>>>(a.partials, grid, 1, scalars, (float)fm_epsilon, 1);
        PAA_LAUNCH_CHECK(h);
        StftArgs b{};
        b.scalars = scalars;
        return run_fused<OP_SCALE>(h, b, q, nullptr, 0.f, p_out, rows, T, out_len, st);
    }
    const long long n = (long long)rows * ((out_len + 3) / 4);
    const int g2 = (int)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 8);
    k_fm_scale_identity<<<std::max(g2, 1), 256, 0, st>>>(q, p_out, rows, T, out_len, h->hop * (n_frames - 1), a.partials, grid,
                                                       (float)fm_epsilon, scalars);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

int paa_stft(paa_handle* h, const float* x, int rows, int T, float* spec, int64_t sb, int64_t sf, int64_t stt,
             void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !x || !spec) return PAA_ERR_NULL;
    int rc = check_time_shape(h, rows, T);
    if (rc) return rc;
    const int n_frames = 1 + T / h->hop;
    StftArgs a{};
    fill_common(h, a, rows, T, n_frames);
    a.x = x; a.vec_ok = aligned16(x) && (T % 4 == 0);
    a.spec_out = reinterpret_cast<float2*>(spec); a.sb = sb; a.sf = sf; a.st = stt;
    a.tiles_per_row = (n_frames + a.frames_per_tile - 1) / a.frames_per_tile;
    return launch_n<SRC_TIME, SINK_SPEC, OP_NONE>(h, a, rows * a.tiles_per_row, (cudaStream_t)stream);
}

int paa_istft(paa_handle* h, const float* spec, int64_t sb, int64_t sf, int64_t stt, int rows, int n_frames, float* y,
              void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !spec || !y) return PAA_ERR_NULL;
    if (rows <= 0 || n_frames < 2) return PAA_ERR_SHAPE;
    if (!nola_ok(h, n_frames)) return PAA_ERR_NOLA;
    StftArgs a{};
    fill_common(h, a, rows, 0, n_frames);
    a.spec_in = reinterpret_cast<const float2*>(spec); a.sb = sb; a.sf = sf; a.st = stt;
    a.y = y; a.out_len = h->hop * (n_frames - 1);
    const int out_blocks = n_frames - 1;
    a.tiles_per_row = std::max(1, (out_blocks + a.blocks_per_tile - 1) / a.blocks_per_tile);
    return launch_n<SRC_SPEC, SINK_TIME, OP_NONE>(h, a, rows * a.tiles_per_row, (cudaStream_t)stream);
}

static int spec_grid(const paa_handle* h, long long n) {
    return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)h->num_sms * 8));
}

int paa_spec_min_max_freqs(paa_handle* h, const float* spec_in, float* spec_out, int rows, int n_frames, int64_t sb,
                           int64_t sf, int64_t stt, double min_freq, double max_freq, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !spec_in || !spec_out) return PAA_ERR_NULL;
    if (rows <= 0 || n_frames <= 0) return PAA_ERR_SHAPE;
    StftArgs a{};
    fill_common(h, a, rows, 0, n_frames);
    a.spec_in = reinterpret_cast<const float2*>(spec_in); a.spec_out = reinterpret_cast<float2*>(spec_out);
    a.sb = sb; a.sf = sf; a.st = stt;
    set_band(h, a, min_freq, max_freq);
    k_spec_op<OP_MASK><<<spec_grid(h, (long long)rows * h->F * n_frames), 256, 0, (cudaStream_t)stream>>>(a, h->F);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

int paa_spec_phon_level(paa_handle* h, const float* spec_in, float* spec_out, int rows, int n_frames, int64_t sb,
                        int64_t sf, int64_t stt, const float* spl_thresh_F, double phon_reference_db, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !spec_in || !spec_out || !spl_thresh_F) return PAA_ERR_NULL;
    if (rows <= 0 || n_frames <= 0) return PAA_ERR_SHAPE;
    StftArgs a{};
    fill_common(h, a, rows, 0, n_frames);
    a.spec_in = reinterpret_cast<const float2*>(spec_in); a.spec_out = reinterpret_cast<float2*>(spec_out);
    a.sb = sb; a.sf = sf; a.st = stt;
    cudaStream_t st = (cudaStream_t)stream;
    a.spl_thresh = spl_thresh_F; a.ref_db = (float)phon_reference_db;
    k_spec_op<OP_PHON_DB><<<spec_grid(h, (long long)rows * h->F * n_frames), 256, 0, st>>>(a, h->F);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

static int spec_fm(paa_handle* h, const float* spec_in, float* spec_out, int rows, int n_frames, int64_t sb, int64_t sf,
                   int64_t stt, double fm_epsilon, int apply, void* scratch, cudaStream_t st) {
    if (!h || !spec_in || !scratch || (apply && !spec_out)) return PAA_ERR_NULL;
    if (!h->d_fm_blob) return PAA_ERR_STATE;
    if (rows <= 0 || n_frames <= 0) return PAA_ERR_SHAPE;
    StftArgs a{};
    fill_common(h, a, rows, 0, n_frames);
    a.spec_in = reinterpret_cast<const float2*>(spec_in); a.spec_out = reinterpret_cast<float2*>(spec_out);
    a.sb = sb; a.sf = sf; a.st = stt;
    a.partials = scratch_partials(scratch);
    a.scalars = scratch_scalars(scratch);
    const long long n = (long long)rows * h->F * n_frames;
    const int grid = std::min(spec_grid(h, n), kMaxPartialBlocks);
    k_spec_fm_partials<<<grid, 256, 0, st>>>(a, h->F);
    PAA_LAUNCH_CHECK(h);
    k_fm_finalize<<<1, 256, 0, st>>>(a.partials, grid, 2, scratch_scalars(scratch), (float)fm_epsilon, apply);
    PAA_LAUNCH_CHECK(h);
    if (apply) {
        k_spec_op<OP_SCALE><<<spec_grid(h, n), 256, 0, st>>>(a, h->F);
        PAA_LAUNCH_CHECK(h);
    }
    return PAA_OK;
}

int paa_spec_fm_norm(paa_handle* h, const float* spec_in, int rows, int n_frames, int64_t sb, int64_t sf, int64_t stt,
                     void* scratch, void* stream) {
    PaaDeviceGuard device_guard(h);
    return spec_fm(h, spec_in, nullptr, rows, n_frames, sb, sf, stt, 0.0, 0, scratch, (cudaStream_t)stream);
}

int paa_spec_fm_project(paa_handle* h, const float* spec_in, float* spec_out, int rows, int n_frames, int64_t sb,
                        int64_t sf, int64_t stt, double fm_epsilon, void* scratch, void* stream) {
    PaaDeviceGuard device_guard(h);
    return spec_fm(h, spec_in, spec_out, rows, n_frames, sb, sf, stt, fm_epsilon, 1, scratch, (cudaStream_t)stream);
}

}  // extern "C"
