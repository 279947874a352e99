// Time-domain half of the hot path for sm_100a: the PGD / Adam step on the perturbation fused
// with the linf / l2 / snr / tv projections (reference: src/training_utils/train.py:155-175,
// src/core/projections.py:11-66).  Pure HBM streaming: float4 loads, one read of every input,
// deterministic block -> grid reductions (no float atomics), device-side branch on the norm so
// the host never synchronises.
#include <cooperative_groups.h>
#include <type_traits>
#include <cstdlib>
#include <cmath>
#include "paa_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kBlocksPerSm = 8;

enum { NORM_L2 = 0, NORM_SNR = 1, NORM_TV = 2 };

// ---- the optimiser step on one element -----------------------------------------------------
// PGD  (train.py:161):  p + fp32(lr) * sign(g), sign(0) = sign(nan) = 0.
// Adam (torch/optim/adam.py:457-547): lerp, mul+addcmul, sqrt/bias2 + eps, addcdiv.
template <int STEP>
__device__ __forceinline__ float step_one(float p, float g, float& m, float& v, const StepDev& s) {
    if ((STEP & 3) == PAA_STEP_PGD) {
        float sg = (float)(g > 0.f) - (float)(g < 0.f);
        return p + s.lr * sg;
    } else if ((STEP & 3) == PAA_STEP_ADAM) {
        m = fmaf(s.w1, g - m, m);
        v = v * s.beta2;
        v = v + (s.w2 * g) * g;
        float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
        return p + (s.neg_step * m) / denom;
    }
    return p;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming loads for data touched once (clean audio): do not let it displace p in L2
__device__ __forceinline__ float4 ld4_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

// The gradient: one buffer, or (mode U, paa_parts) the sum of up to PAA_MAX_PARTS per-rank partials that live in peer
// GPUs' memory.  Plain loads reach peer-mapped addresses over NVLink; the parts are added left to right in fp32, so
// every rank that passes the same list computes the same bits -- the all-reduce is folded into this pass.
// The multi-part form is a separate instantiation (bit 2 of the STEP template value, kStepParts) so that the
// single-GPU kernels keep their instruction count and register budget.
constexpr int kStepParts = 4;
template <int STEP>
__device__ __forceinline__ float4 ldg4(const StepDev& s, int64_t i) {
    float4 g = ld4(s.grad + i);
    if (STEP & kStepParts) {
#pragma unroll
        for (int k = 1; k < PAA_MAX_PARTS; ++k)
            if (k < s.nparts) { const float4 b = ld4(s.gpart[k] + i); g.x += b.x; g.y += b.y; g.z += b.z; g.w += b.w; }
    }
    return g;
}
template <int STEP>
__device__ __forceinline__ float ldg1(const StepDev& s, int64_t i) {
    float g = s.grad[i];
    if (STEP & kStepParts) {
#pragma unroll
        for (int k = 1; k < PAA_MAX_PARTS; ++k)
            if (k < s.nparts) g += s.gpart[k][i];
    }
    return g;
}

// Loads 4 elements of p (and grad / Adam state), applies the step, stores the Adam state.
template <int STEP, bool WRITE_STATE>
__device__ __forceinline__ float4 stepped4(const float* p, int64_t i, const StepDev& s) {
    float4 x = ld4(p + i);
    if ((STEP & 3) == PAA_STEP_NONE) return x;
    float4 g = ldg4<STEP>(s, i);
    float4 m = make_float4(0, 0, 0, 0), v = m;
    if ((STEP & 3) == PAA_STEP_ADAM) { m = ld4(s.m + i); v = ld4(s.v + i); }
    x.x = step_one<STEP & 3>(x.x, g.x, m.x, v.x, s);
    x.y = step_one<STEP & 3>(x.y, g.y, m.y, v.y, s);
    x.z = step_one<STEP & 3>(x.z, g.z, m.z, v.z, s);
    x.w = step_one<STEP & 3>(x.w, g.w, m.w, v.w, s);
    if ((STEP & 3) == PAA_STEP_ADAM && WRITE_STATE) { st4(s.m + i, m); st4(s.v + i, v); }
    return x;
}
template <int STEP, bool WRITE_STATE>
__device__ __forceinline__ float stepped1(const float* p, int64_t i, const StepDev& s) {
    float x = p[i];
    if ((STEP & 3) == PAA_STEP_NONE) return x;
    float g = ldg1<STEP>(s, i), m = 0.f, v = 0.f;
    if ((STEP & 3) == PAA_STEP_ADAM) { m = s.m[i]; v = s.v[i]; }
    x = step_one<STEP & 3>(x, g, m, v, s);
    if ((STEP & 3) == PAA_STEP_ADAM && WRITE_STATE) { s.m[i] = m; s.v[i] = v; }
    return x;
}

// The same in two halves, so that the loads of the next float4 can be issued before the arithmetic on the
// current one (software pipelining in k_fused: memory-level parallelism, not instruction count, bounds it).
// np/ng: the element after the float4 (tv, last lane only); c/nc: the clean audio at the same index (snr, tv)
struct Raw4 { float4 p, g, m, v; float np, ng; float4 c; float nc; };
// STREAM: p and the gradient are read once per call; as streaming (evict-first) loads they do not displace the stepped
// values k_fused parks in L2 between its two phases (the part that does not fit registers + shared memory), so that
// part's round trip stays in L2: tv 128 x 10 s 82.3 -> 72.2 us, snr 32 x 10 s 22.1 -> 21.7 us.  Not for l2 at 512 x 10 s
// (the parked part, 280 MB, exceeds L2: 0.7 % slower) nor for the Adam forms (four input streams: 10 % slower on snr).
template <int STEP, bool STREAM>
__device__ __forceinline__ void load_raw4(Raw4& r, const float* p, int64_t i, const StepDev& s) {
    r.p = STREAM ? ld4_stream(p + i) : ld4(p + i);
    if ((STEP & 3) != PAA_STEP_NONE) r.g = (STREAM && !(STEP & kStepParts)) ? ld4_stream(s.grad + i) : ldg4<STEP>(s, i);
    if ((STEP & 3) == PAA_STEP_ADAM) { r.m = ld4(s.m + i); r.v = ld4(s.v + i); }
}
template <int STEP, bool WRITE_STATE>
__device__ __forceinline__ float4 finish4(Raw4 r, int64_t i, const StepDev& s) {
    if ((STEP & 3) == PAA_STEP_NONE) return r.p;
    float4 x = r.p;
    x.x = step_one<STEP & 3>(x.x, r.g.x, r.m.x, r.v.x, s);
    x.y = step_one<STEP & 3>(x.y, r.g.y, r.m.y, r.v.y, s);
    x.z = step_one<STEP & 3>(x.z, r.g.z, r.m.z, r.v.z, s);
    x.w = step_one<STEP & 3>(x.w, r.g.w, r.m.w, r.v.w, s);
    if ((STEP & 3) == PAA_STEP_ADAM && WRITE_STATE) { st4(s.m + i, r.m); st4(s.v + i, r.v); }
    return x;
}

// torch.clamp: NaN stays NaN, bounds applied as min(max(x, lo), hi)
__device__ __forceinline__ float clamp1(float x, float lo, float hi) {
    return (x != x) ? x : fminf(fmaxf(x, lo), hi);
}

// ---- block reduction in double, fixed order -------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
template <int NQ, int NT = kThreads>
__device__ __forceinline__ void block_sum(double (&acc)[NQ], double* out /* [NQ] or nullptr */) {
    __shared__ double sh[NQ][NT / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double s = warp_sum(acc[q]);
        if (lane == 0) sh[q][w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double s = lane < NT / 32 ? sh[q][lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0 && out) out[q] = s;
        }
    }
}

// ---- kernel: step + clamp (linf), single pass -------------------------------------------------
template <int STEP, bool VEC>
__global__ void __launch_bounds__(kThreads) k_step_clamp(const float* p_in, float* p_out,
                                                        int64_t n, float lo, float hi, bool clamp, StepDev s) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            float4 x = stepped4<STEP, true>(p_in, i * 4, s);
            if (clamp) { x.x = clamp1(x.x, lo, hi); x.y = clamp1(x.y, lo, hi); x.z = clamp1(x.z, lo, hi); x.w = clamp1(x.w, lo, hi); }
            st4(p_out + i * 4, x);
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += nth) {
            float x = stepped1<STEP, true>(p_in, i, s);
            p_out[i] = clamp ? clamp1(x, lo, hi) : x;
        }
    } else {
        for (int64_t i = tid; i < n; i += nth) {
            float x = stepped1<STEP, true>(p_in, i, s);
            p_out[i] = clamp ? clamp1(x, lo, hi) : x;
        }
    }
}

// ---- kernel: pass A of l2 / snr / tv ----------------------------------------------------------
// Reads p (+grad, Adam state) once, writes the stepped perturbation q, and leaves per-block
// partial sums of   l2: sum q^2      snr: sum q^2, sum clean^2      tv: sum |dq|, sum |dclean|.
struct ReduceArgs {
    const float* p_in;
    float* q_out;        // stepped perturbation (may alias p_in unless NORM_TV with a step)
    int64_t n;           // rows*T
    int T;               // row length of p (tv: differences never cross a row end)
    const float* clean;
    int64_t clean_n;
    int clean_T;
    double* partials;    // [gridDim.x][2]
    bool write_q;
};

template <int NORM, int STEP, bool VEC>
__global__ void __launch_bounds__(kThreads) k_reduce(ReduceArgs a, StepDev s) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
    float acc0 = 0.f, acc1 = 0.f;       // per-thread fp32 partials (a few hundred terms), widened below
    double wide[2] = {0.0, 0.0};

    if (NORM != NORM_TV) {
        if (VEC) {
            const int64_t n4 = a.n >> 2;
            for (int64_t i = tid; i < n4; i += nth) {
                float4 x = stepped4<STEP, true>(a.p_in, i * 4, s);
                if (a.write_q) st4(a.q_out + i * 4, x);
                acc0 += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
            }
            for (int64_t i = (n4 << 2) + tid; i < a.n; i += nth) {
                float x = stepped1<STEP, true>(a.p_in, i, s);
                if (a.write_q) a.q_out[i] = x;
                acc0 += x * x;
            }
            if (NORM == NORM_SNR) {
                const int64_t c4 = a.clean_n >> 2;
                for (int64_t i = tid; i < c4; i += nth) {
                    float4 c = ld4_stream(a.clean + i * 4);
                    acc1 += (c.x * c.x + c.y * c.y) + (c.z * c.z + c.w * c.w);
                }
                for (int64_t i = (c4 << 2) + tid; i < a.clean_n; i += nth) { float c = a.clean[i]; acc1 += c * c; }
            }
        } else {
            for (int64_t i = tid; i < a.n; i += nth) {
                float x = stepped1<STEP, true>(a.p_in, i, s);
                if (a.write_q) a.q_out[i] = x;
                acc0 += x * x;
            }
            if (NORM == NORM_SNR)
                for (int64_t i = tid; i < a.clean_n; i += nth) { float c = a.clean[i]; acc1 += c * c; }
        }
    } else {
        // total variation: element i pairs with i+1 unless i is the last sample of its row.
        // The neighbour's stepped value is recomputed (never read back from q_out), which is
        // why tv with a fused step needs q_out != p_in.
        if (VEC) {
            const int64_t n4 = a.n >> 2;
            for (int64_t i4 = tid; i4 < n4; i4 += nth) {
                const int64_t i = i4 * 4;
                float4 x = stepped4<STEP, false>(a.p_in, i, s);
                if (a.write_q) st4(a.q_out + i, x);
                float nx = (i + 4 < a.n) ? stepped1<STEP, false>(a.p_in, i + 4, s) : 0.f;
                const int c = (int)(i % a.T), Tp = a.T;
                float t = 0.f;
                if (c != Tp - 1) t += fabsf(x.y - x.x);
                if ((c + 1) % Tp != Tp - 1) t += fabsf(x.z - x.y);
                if ((c + 2) % Tp != Tp - 1) t += fabsf(x.w - x.z);
                if ((c + 3) % Tp != Tp - 1 && i + 4 < a.n) t += fabsf(nx - x.w);
                acc0 += t;
            }
            for (int64_t i = (n4 << 2) + tid; i < a.n; i += nth) {
                float x = stepped1<STEP, false>(a.p_in, i, s);
                if (a.write_q) a.q_out[i] = x;
                if ((int)(i % a.T) != a.T - 1 && i + 1 < a.n) acc0 += fabsf(stepped1<STEP, false>(a.p_in, i + 1, s) - x);
            }
            const int64_t c4 = a.clean_n >> 2;
            for (int64_t i4 = tid; i4 < c4; i4 += nth) {
                const int64_t i = i4 * 4;
                float4 x = ld4(a.clean + i);
                float nx = (i + 4 < a.clean_n) ? a.clean[i + 4] : 0.f;
                const int c = (int)(i % a.clean_T), Tc = a.clean_T;
                float t = 0.f;
                if (c != Tc - 1) t += fabsf(x.y - x.x);
                if ((c + 1) % Tc != Tc - 1) t += fabsf(x.z - x.y);
                if ((c + 2) % Tc != Tc - 1) t += fabsf(x.w - x.z);
                if ((c + 3) % Tc != Tc - 1 && i + 4 < a.clean_n) t += fabsf(nx - x.w);
                acc1 += t;
            }
            for (int64_t i = (c4 << 2) + tid; i < a.clean_n; i += nth)
                if ((int)(i % a.clean_T) != a.clean_T - 1 && i + 1 < a.clean_n) acc1 += fabsf(a.clean[i + 1] - a.clean[i]);
        } else {
            for (int64_t i = tid; i < a.n; i += nth) {
                float x = stepped1<STEP, false>(a.p_in, i, s);
                if (a.write_q) a.q_out[i] = x;
                if ((int)(i % a.T) != a.T - 1 && i + 1 < a.n) acc0 += fabsf(stepped1<STEP, false>(a.p_in, i + 1, s) - x);
            }
            for (int64_t i = tid; i < a.clean_n; i += nth)
                if ((int)(i % a.clean_T) != a.clean_T - 1 && i + 1 < a.clean_n) acc1 += fabsf(a.clean[i + 1] - a.clean[i]);
        }
    }
    wide[0] = (double)acc0;
    wide[1] = (double)acc1;
    block_sum<2>(wide, a.partials + 2 * (int64_t)blockIdx.x);
}

// ---- kernel: finalize -- fixed-order sum of the partials, then the reference's fp32 branch ---
struct FinalArgs {
    const double* partials;
    int nblocks;
    float* scalars;
    float eps;           // l2: epsilon; tv: tv_epsilon; snr: snr_db
    double snr_linear;   // 10**(snr_db/10), python double
    double n_p;          // numel(p)
    double n_clean;      // numel(clean)
    // mode U: the clean-audio sum (snr: index 0, tv: index 1) is the sum of the per-rank statistics, index order
    const double* cstat[PAA_MAX_PARTS];
    int ncstat, cstat_idx;
    int numel_from_stats;   // mode U with paa_parts.clean_numel == 0: numel(clean) = sum of the parts' third statistic
};
__device__ __forceinline__ double clean_total(const FinalArgs& a, double local) {
    if (a.ncstat == 0) return local;
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < PAA_MAX_PARTS; ++k)
        if (k < a.ncstat) t += a.cstat[k][a.cstat_idx];
    return t;
}

// numel of the whole clean batch: a host constant, or (mode U, uneven shards) the sum of the per-rank counts
__device__ __forceinline__ double clean_numel(const FinalArgs& a) {
    if (!a.numel_from_stats) return a.n_clean;
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < PAA_MAX_PARTS; ++k)
        if (k < a.ncstat) t += a.cstat[k][2];
    return t;
}

// The reference's data-dependent branch, in its fp32 arithmetic, from the two global sums.
template <int NORM>
__device__ __forceinline__ void final_math(double tot0, double tot1, const FinalArgs& a, float& scale, float& norm,
                                           float& aux0, float& aux1) {
    scale = 1.f; norm = 0.f; aux0 = 0.f; aux1 = 0.f;
    const double n_clean = NORM == NORM_SNR ? clean_numel(a) : 1.0;
    if (NORM == NORM_L2) {                                   // projections.py:41-46
        norm = sqrtf((float)tot0);
        if (norm > a.eps) scale = __frcp_rn(norm) * a.eps;     // python `eps / tensor` = reciprocal()*eps
    } else if (NORM == NORM_SNR) {                           // projections.py:11-35
        const float p_noise = (float)(tot0 / a.n_p);
        const float p_sig = (float)(tot1 / n_clean);
        aux0 = p_sig;
        aux1 = 10.f * log10f(p_sig / (p_noise + 1e-12f));
        norm = sqrtf((float)tot0);
        if (!(aux1 >= a.eps) && !(norm < 1e-8f)) {
            const float want = sqrtf((p_sig / (float)a.snr_linear) * (float)n_clean);
            scale = want / norm;
        }
    } else {                                                 // projections.py:56-66
        norm = (float)tot0;
        aux0 = (float)tot1;
        aux1 = a.eps * aux0;
        if (norm > aux1) scale = aux1 / norm;
    }
}

template <int NORM>
__global__ void __launch_bounds__(kThreads) k_finalize(FinalArgs a) {
    double acc[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < a.nblocks; i += kThreads) {
        acc[0] += a.partials[2 * i];
        acc[1] += a.partials[2 * i + 1];
    }
    __shared__ double tot[2];
    block_sum<2>(acc, tot);
    __syncthreads();
    if (threadIdx.x != 0) return;
    float scale, norm, aux0, aux1;
    final_math<NORM>(tot[0], clean_total(a, tot[1]), a, scale, norm, aux0, aux1);
    a.scalars[PAA_S_SCALE] = scale;
    a.scalars[PAA_S_NORM] = norm;
    a.scalars[PAA_S_AUX0] = aux0;
    a.scalars[PAA_S_AUX1] = aux1;
}

// ---- kernel: pass B -- q *= scale (skipped on the device when scale == 1 and in place) --------
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_scale(const float* q_in, float* p_out, int64_t n,
                                                   const float* __restrict__ scalars) {
    const float sc = scalars[PAA_S_SCALE];
    if (sc == 1.f && q_in == p_out) return;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            float4 x = ld4(q_in + i * 4);
            x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
            st4(p_out + i * 4, x);
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += nth) p_out[i] = q_in[i] * sc;
    } else {
        for (int64_t i = tid; i < n; i += nth) p_out[i] = q_in[i] * sc;
    }
}

// |x1-x0| + |x2-x1| + |x3-x2| + |next-x3| for the float4 whose first element sits in column `col` of a row of
// length T >= 4; a pair is dropped when its left element is the last sample of a row (projections.py:58,62).
__device__ __forceinline__ float tv_quad(float4 x, float nx, int col, int T, bool has_next) {
    if ((T & 3) == 0) {
        // rows are whole float4s (every BASELINE shape): the three inner pairs never leave the row, only the pair with the
        // next float4 can (kernel-uniform branch; 10 instructions less per float4 than the general form below)
        float t = (fabsf(x.y - x.x) + fabsf(x.z - x.y)) + fabsf(x.w - x.z);
        if (col + 4 != T && has_next) t += fabsf(nx - x.w);
        return t;
    }
    int c1 = col + 1, c2 = col + 2, c3 = col + 3;
    if (c1 >= T) c1 -= T;
    if (c2 >= T) c2 -= T;
    if (c3 >= T) c3 -= T;
    float t = 0.f;
    if (col != T - 1) t += fabsf(x.y - x.x);
    if (c1 != T - 1) t += fabsf(x.z - x.y);
    if (c2 != T - 1) t += fabsf(x.w - x.z);
    if (c3 != T - 1 && has_next) t += fabsf(nx - x.w);
    return t;
}

// ---- kernel: the whole reducing projection in ONE cooperative launch ----------------------------
// Phase A (step + reduce) -> grid barrier -> every block re-sums the partials in the same fixed
// order -> phase B (rescale).  Each thread revisits exactly the float4s it produced in phase A, and
// keeps the first kRC of them in registers and the next `sc_iters` in shared memory across the
// barrier, so up to (kRC + sc_iters) * 4 * gridDim.x * kFT elements (~12 M on a B200) never make
// the q round trip through HBM: 12 B/element of traffic instead of the 20 B the two-pass form needs.
constexpr int kFT = 512;
constexpr int kRC = 4;
constexpr int kSCMax = 12;

// PARKED: some of a thread's float4s do not fit registers + shared memory and wait in q_out (HBM / L2) between the phases.
// The lean form (PARKED = false: every BASELINE shape up to 12 M elements, and all universal shapes) compiles that path out.
template <int NORM, int STEP, bool RIDE, bool PARKED>
__global__ void __launch_bounds__(kFT, 2) k_fused(ReduceArgs a, StepDev s, FinalArgs f, int sc_iters) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ float4 cache[];                       // [sc_iters][kFT]
    // tv indexes its float4s with 32 bits (the host takes this kernel only when n/4 and clean_n/4 stay below 2^31: 8.6 G
    // elements): its loops compare, advance and branch on the index several times per float4, 64-bit integer arithmetic was
    // its largest instruction class, and narrowing it took 128 x 10 s from 91.4 to 82.2 us; addresses are widened where
    // they are formed.  l2 / snr keep 64-bit indices (one pointer increment per float4: 32-bit measured 2.8 % slower on l2).
    typedef typename std::conditional<NORM == NORM_TV, unsigned, int64_t>::type idx_t;
    const idx_t tid = (idx_t)blockIdx.x * kFT + threadIdx.x, nth = (idx_t)gridDim.x * kFT;
    const idx_t n4 = (idx_t)(a.n >> 2);
    const idx_t last4 = a.n > 0 ? (idx_t)((a.n - 1) >> 2) : 0;                    // i4 < last4  <=>  4 i4 + 4 < n
    const idx_t lastc4 = a.clean_n > 0 ? (idx_t)((a.clean_n - 1) >> 2) : 0;
    float acc0 = 0.f, acc1 = 0.f;
    constexpr int RCN = ((STEP & 3) == PAA_STEP_ADAM) ? 2 : kRC;      // Adam carries 4 streams: fewer register-resident float4s
    float4 keep[RCN];

    // total variation: row-column of each thread's current float4, advanced without 64-bit division
    const int lane = threadIdx.x & 31;
    int colp = 0, stepp = 0;
    if (NORM == NORM_TV) { colp = (int)(((int64_t)tid * 4) % a.T); stepp = (int)(((int64_t)nth * 4) % a.T); }

    // tv: the clean audio rides the same loop as the perturbation -- float4 i4 of p, grad and clean are requested
    // together, so every thread keeps three independent streams in flight (93.6 us against 96.1 us at 128 x 10 s);
    // what is left of a longer clean tensor (universal (1,T) p against a (B,T) batch) follows in a clean-only loop with
    // four float4 in flight per thread.  snr rides only when the clean tensor is longer than the perturbation (the
    // universal shape, latency bound: the first clean loads leave with the kernel's first instruction, 12.0 against
    // 14.3 us at (1,T) x 32 x 10 s); with one row per utterance riding costs registers and measured 2.4 % slower
    // (22.4 against 21.9 us), so that shape keeps the clean-only loop (RIDE = false).
    constexpr bool kRide = RIDE;
    const idx_t c4 = (NORM == NORM_L2) ? 0 : (idx_t)(a.clean_n >> 2);
    const int Tc = a.clean_T;
    int colc = 0, stepc = 0;
    if (NORM == NORM_TV) { colc = (int)(((int64_t)tid * 4) % Tc); stepc = (int)(((int64_t)nth * 4) % Tc); }

    // clean part of one float4 (every lane of the warp calls it for tv: the element after the float4 comes from the
    // next lane by shuffle, only the last lane / the last float4 took it from memory in fetch)
    auto clean_a = [&](const Raw4& raw, idx_t i4) {
        if (NORM == NORM_SNR) {
            if (i4 < c4) acc1 += (raw.c.x * raw.c.x + raw.c.y * raw.c.y) + (raw.c.z * raw.c.z + raw.c.w * raw.c.w);
        } else if (NORM == NORM_TV) {
            const bool act = i4 < c4;
            const bool has_next = act && i4 < lastc4;
            float nx = __shfl_down_sync(0xffffffffu, raw.c.x, 1);
            if (has_next && (lane == 31 || i4 + 1 >= c4)) nx = raw.nc;
            if (act) acc1 += tv_quad(raw.c, nx, colc, Tc, has_next);
            colc += stepc;
            if (colc >= Tc) colc -= Tc;
        }
    };
    // phase A on one float4 whose operands are already in registers: step, accumulate the norm, hand back the
    // stepped values.  For tv every lane of the warp calls it (act = in range): the element after the float4 comes
    // from the next lane by shuffle, only the last lane (or the last float4) recomputes it from memory.
    auto phase_a = [&](const Raw4& raw, idx_t i4, bool act) -> float4 {
        const int64_t i = (int64_t)i4 * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (NORM != NORM_TV) {
            if (act) {
                x = finish4<STEP, true>(raw, i, s);
                acc0 += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
            }
        } else {
            if (act) x = finish4<STEP, false>(raw, i, s);
            float nx = __shfl_down_sync(0xffffffffu, x.x, 1);
            const bool has_next = act && i4 < last4;
            if (has_next && (lane == 31 || i4 + 1 >= n4)) {       // its operands came with the float4 (fetch)
                float m1 = 0.f, v1 = 0.f;
                nx = step_one<(STEP & 3) == PAA_STEP_ADAM ? PAA_STEP_NONE : (STEP & 3)>(raw.np, raw.ng, m1, v1, s);
            }
            if (act) acc0 += tv_quad(x, nx, colp, a.T, has_next);
            colp += stepp;
            if (colp >= a.T) colp -= a.T;
        }
        if (kRide) clean_a(raw, i4);
        return x;
    };
    auto fetch_clean = [&](Raw4& r, idx_t i4) {
        r.c = make_float4(0.f, 0.f, 0.f, 0.f);
        r.nc = 0.f;
        if (NORM != NORM_L2 && i4 < c4) {
            r.c = ld4_stream(a.clean + (int64_t)i4 * 4);
            if (NORM == NORM_TV && (lane == 31 || i4 + 1 >= c4) && i4 < lastc4) r.nc = a.clean[(int64_t)i4 * 4 + 4];
        }
    };
    auto fetch = [&](idx_t i4, bool act) -> Raw4 {
        Raw4 r;
        r.p = r.g = r.m = r.v = make_float4(0.f, 0.f, 0.f, 0.f);
        r.np = r.ng = 0.f;
        if (act) load_raw4<STEP, (STEP & 3) != PAA_STEP_ADAM>(r, a.p_in, (int64_t)i4 * 4, s);
        if (NORM == NORM_TV && act && (lane == 31 || i4 + 1 >= n4) && i4 < last4) {
            r.np = a.p_in[(int64_t)i4 * 4 + 4];
            if ((STEP & 3) != PAA_STEP_NONE) r.ng = ldg1<STEP>(s, (int64_t)i4 * 4 + 4);
        }
        if (kRide) fetch_clean(r, i4);
        return r;
    };

    {   // the first RCN float4s, two at a time (six or more loads in flight per thread); results stay in registers
        constexpr int G = 2;
        static_assert(RCN % G == 0, "RCN is fetched in groups of G");
#pragma unroll
        for (int k0 = 0; k0 < RCN; k0 += G) {
            Raw4 raw[G];
#pragma unroll
            for (int k = 0; k < G; ++k) raw[k] = fetch(tid + (k0 + k) * nth, tid + (k0 + k) * nth < n4);
#pragma unroll
            for (int k = 0; k < G; ++k) {
                const idx_t i4 = tid + (k0 + k) * nth;
                const bool act = i4 < n4;
                const float4 x = phase_a(raw[k], i4, act);
                if (NORM == NORM_TV || act) keep[k0 + k] = x;
            }
        }
    }
    idx_t i4 = tid + RCN * nth;
    {   // the rest, software-pipelined two deep; (i4 - lane) is warp-uniform so whole warps stay for the shuffles
        int k = RCN;
        Raw4 cur = fetch(i4, i4 < n4);
        while (i4 - lane < n4) {
            const idx_t i4n = i4 + nth;
            const Raw4 nxt = fetch(i4n, i4n < n4);
            const bool act = i4 < n4;
            const float4 x = phase_a(cur, i4, act);
            if (act) {
                if (!PARKED || k < RCN + sc_iters) cache[(k - RCN) * kFT + threadIdx.x] = x;
                else if (a.write_q) st4(a.q_out + (int64_t)i4 * 4, x);
            }
            cur = nxt;
            i4 = i4n;
            ++k;
        }
        // `cur` holds the (clean-only) float4 at i4, fetched but not consumed yet
        if (kRide && i4 - lane < c4) {
            clean_a(cur, i4);
            i4 += nth;
        }
    }
    if (!kRide) i4 = tid;
    if (tid == 0) {                                         // the n % 4 trailing elements go through global memory
        for (int64_t i = (int64_t)n4 << 2; i < a.n; ++i) {
            const float x = stepped1<STEP, NORM != NORM_TV>(a.p_in, i, s);
            a.q_out[i] = x;
            if (NORM != NORM_TV) acc0 += x * x;
            else if ((int)(i % a.T) != a.T - 1 && i + 1 < a.n) acc0 += fabsf(stepped1<STEP, false>(a.p_in, i + 1, s) - x);
        }
    }
    if (NORM != NORM_L2) {
        // what is left of the clean audio behind the perturbation's range (four float4 in flight per thread)
        constexpr int U = 4;
        while (i4 - lane < c4) {
            Raw4 r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) fetch_clean(r[u], i4 + u * nth);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i4 + u * nth - lane < c4) clean_a(r[u], i4 + u * nth);
            i4 += U * nth;
        }
        if (NORM == NORM_SNR) {
            for (int64_t i = ((int64_t)c4 << 2) + tid; i < a.clean_n; i += nth) { const float c = a.clean[i]; acc1 += c * c; }
        } else {
            for (int64_t i = ((int64_t)c4 << 2) + tid; i < a.clean_n; i += nth)
                if ((int)(i % Tc) != Tc - 1 && i + 1 < a.clean_n) acc1 += fabsf(a.clean[i + 1] - a.clean[i]);
        }
    }
    double wide[2] = {(double)acc0, (double)acc1};
    block_sum<2, kFT>(wide, a.partials + 2 * (int64_t)blockIdx.x);
    __threadfence();
    grid.sync();

    // every block sums the same partials in the same order: identical scale everywhere, no second barrier
    double part[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kFT) {
        part[0] += a.partials[2 * i];
        part[1] += a.partials[2 * i + 1];
    }
    __shared__ double tot[2];
    __shared__ float s_scale;
    block_sum<2, kFT>(part, tot);
    __syncthreads();
    if (threadIdx.x == 0) {
        float scale, norm, aux0, aux1;
        final_math<NORM>(tot[0], clean_total(f, tot[1]), f, scale, norm, aux0, aux1);
        s_scale = scale;
        if (blockIdx.x == 0) {
            f.scalars[PAA_S_SCALE] = scale;
            f.scalars[PAA_S_NORM] = norm;
            f.scalars[PAA_S_AUX0] = aux0;
            f.scalars[PAA_S_AUX1] = aux1;
        }
    }
    __syncthreads();
    const float sc = s_scale;
    if (sc == 1.f && !a.write_q) return;                    // in place and already feasible: nothing to store

    auto scaled = [sc](float4 x) { x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; return x; };
#pragma unroll
    for (int k = 0; k < RCN; ++k) {
        const idx_t i4 = tid + k * nth;
        if (i4 < n4) st4(a.q_out + (int64_t)i4 * 4, scaled(keep[k]));
    }
    // What did not fit on chip was written to q_out in phase A, in index order.  The PGD forms revisit it NEWEST FIRST:
    // the most recently written part is still in the 126 MB L2 (the inputs are streamed past it), and walking the same way
    // again would evict every line just before it is needed (l2 512 x 10 s parks 280 MB: 267.5 -> 253 us).  The Adam forms
    // (four more streams through L2, nothing to gain) keep the single forward loop, and its register budget.
    constexpr bool kNewestFirst = PARKED && (STEP & 3) != PAA_STEP_ADAM;
    if (kNewestFirst) {
        int k = RCN;
        idx_t i4 = tid + RCN * nth;
        for (; i4 < n4 && k < RCN + sc_iters; i4 += nth, ++k)
            st4(a.q_out + (int64_t)i4 * 4, scaled(cache[(k - RCN) * kFT + threadIdx.x]));
        if (sc != 1.f && i4 < n4) {
            idx_t j = i4 + (idx_t)((n4 - 1 - i4) / nth) * nth;           // this thread's last float4
            for (;; j -= nth) {
                st4(a.q_out + (int64_t)j * 4, scaled(ld4(a.q_out + (int64_t)j * 4)));
                if (j == i4) break;
            }
        }
    } else {
        int k = RCN;
        for (idx_t i4 = tid + RCN * nth; i4 < n4; i4 += nth, ++k) {
            if (!PARKED || k < RCN + sc_iters) st4(a.q_out + (int64_t)i4 * 4, scaled(cache[(k - RCN) * kFT + threadIdx.x]));
            else if (sc != 1.f) st4(a.q_out + (int64_t)i4 * 4, scaled(ld4(a.q_out + (int64_t)i4 * 4)));
        }
    }
    if (tid == 0 && sc != 1.f)
        for (int64_t i = (int64_t)n4 << 2; i < a.n; ++i) a.q_out[i] *= sc;
}

// ---- mode U helpers ------------------------------------------------------------------------------------------
// Per-rank clean statistics: sum x^2 and sum_t |x[r,t+1] - x[r,t]| (no pair across a row end), block partials.
__global__ void __launch_bounds__(kThreads) k_clean_stats(const float* __restrict__ x, int64_t n, int T, double* partials) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
    float e = 0.f, tv = 0.f;
    int col = (int)(tid % T);
    const int dcol = (int)(nth % T);
    for (int64_t i = tid; i < n; i += nth) {
        const float v = x[i];
        e += v * v;
        if (col != T - 1) tv += fabsf(x[i + 1] - v);
        col += dcol;
        if (col >= T) col -= T;
    }
    double wide[2] = {(double)e, (double)tv};
    block_sum<2>(wide, partials + 2 * (int64_t)blockIdx.x);
}
__global__ void __launch_bounds__(kThreads) k_clean_stats_final(const double* partials, int nblocks, double* out3, double numel) {
    double acc[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += kThreads) { acc[0] += partials[2 * i]; acc[1] += partials[2 * i + 1]; }
    __shared__ double tot[2];
    block_sum<2>(acc, tot);
    __syncthreads();
    if (threadIdx.x == 0) { out3[0] = tot[0]; out3[1] = tot[1]; out3[2] = numel; }
}
// out = sum of the gradient parts (for the STFT-domain projections, whose tiles re-read halo samples)
__global__ void __launch_bounds__(kThreads) k_sum_parts(StepDev s, float* out, int64_t n, bool vec) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
    if (vec) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) st4(out + i * 4, ldg4<kStepParts>(s, i * 4));
        for (int64_t i = (n4 << 2) + tid; i < n; i += nth) out[i] = ldg1<kStepParts>(s, i);
    } else {
        for (int64_t i = tid; i < n; i += nth) out[i] = ldg1<kStepParts>(s, i);
    }
}

// ---- kernels: compose + clamp (train.py:136) and its backward -- the input side of the path ----
//   forward : x_adv[b,t] = clamp(clean[b,t] + p[b or 0, t], -1, 1)
//   backward: dL/dp[b or 0, t] = (sum over b of) dL/dx_adv[b,t] * 1[-1 <= clean[b,t] + p[.,t] <= 1]   (torch's clamp mask)
// grid.x covers float4 columns, grid.y strides over rows, four rows per trip and (the host sizes grid.y so) ONE trip per CTA:
// measured on B200, the finest decomposition wins (5 024 CTAs for 128 x 10 s: 32.9 us against 34.9 us with 1 256 CTAs and
// 36.9 us with exactly one resident wave); a universal (1,T) p is re-read from L2, never from HBM.
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_compose(const float* __restrict__ clean, const float* __restrict__ p,
                                                     float* __restrict__ out, int rows, int p_rows, int T) {
    const int W = VEC ? 4 : 1;
    const int cols = (T + W - 1) / W;
    for (int c = blockIdx.x * kThreads + threadIdx.x; c < cols; c += gridDim.x * kThreads) {
        if (VEC) {
            // four rows per trip, all loads issued before the first store (memory-level parallelism)
            constexpr int U = 4;
            for (int b0 = blockIdx.y; b0 < rows; b0 += U * gridDim.y) {
                float4 x[U], q[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int b = b0 + u * gridDim.y;
                    if (b < rows) {
                        x[u] = ld4_stream(clean + (size_t)b * T + (size_t)c * 4);
                        q[u] = (p_rows == 1 && u > 0) ? q[0] : ld4(p + (p_rows == 1 ? 0 : (size_t)b * T) + (size_t)c * 4);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int b = b0 + u * gridDim.y;
                    if (b < rows) {
                        float4 v = x[u];
                        v.x = clamp1(v.x + q[u].x, -1.f, 1.f); v.y = clamp1(v.y + q[u].y, -1.f, 1.f);
                        v.z = clamp1(v.z + q[u].z, -1.f, 1.f); v.w = clamp1(v.w + q[u].w, -1.f, 1.f);
                        st4(out + (size_t)b * T + (size_t)c * 4, v);
                    }
                }
            }
        } else {
            for (int b = blockIdx.y; b < rows; b += gridDim.y) {
                const size_t i = (size_t)b * T + (size_t)c * W, pi = (p_rows == 1 ? 0 : (size_t)b * T) + (size_t)c * W;
                out[i] = clamp1(clean[i] + p[pi], -1.f, 1.f);
            }
        }
    }
}

__device__ __forceinline__ float pass1(float x, float q, float g) {
    const float s = x + q;
    return (s >= -1.f && s <= 1.f) ? g : 0.f;                  // NaN input: mask false, as torch's comparison
}

// per-utterance p: element-wise, 2-D grid as the forward: two rows per trip, all loads issued before the first store, one
// trip per CTA (12.3 / 49.3 us at 32 / 128 x 10 s; one row per trip on 8 grid rows, the round-1 form: 12.6 / 53.3 us)
constexpr int kComposeBwdRows = 2;
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_compose_bwd(const float* __restrict__ clean, const float* __restrict__ p,
                                                         const float* __restrict__ gx, float* __restrict__ gp, int rows, int T) {
    const int W = VEC ? 4 : 1;
    const int cols = (T + W - 1) / W;
    for (int c = blockIdx.x * kThreads + threadIdx.x; c < cols; c += gridDim.x * kThreads) {
        if (VEC) {
            constexpr int U = kComposeBwdRows;
            for (int b0 = blockIdx.y; b0 < rows; b0 += U * gridDim.y) {
                float4 x[U], q[U], g[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int b = b0 + u * gridDim.y;
                    if (b < rows) {
                        const size_t i = (size_t)b * T + (size_t)c * 4;
                        x[u] = ld4_stream(clean + i); q[u] = ld4(p + i); g[u] = ld4_stream(gx + i);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int b = b0 + u * gridDim.y;
                    if (b < rows)
                        st4(gp + (size_t)b * T + (size_t)c * 4, make_float4(pass1(x[u].x, q[u].x, g[u].x), pass1(x[u].y, q[u].y, g[u].y),
                                                                           pass1(x[u].z, q[u].z, g[u].z), pass1(x[u].w, q[u].w, g[u].w)));
                }
            }
        } else {
            for (int b = blockIdx.y; b < rows; b += gridDim.y) {
                const size_t i = (size_t)b * T + (size_t)c * W;
                gp[i] = pass1(clean[i], p[i], gx[i]);
            }
        }
    }
}

// universal (1,T) p: dL/dp[t] = sum over the batch.  The batch is split over the CTAs of a thread-block cluster
// (cluster = (1, K, 1): K CTAs share a column range, CTA y takes the rows y, y + K, ...), each thread keeps its column's
// partial sum in registers, and the K partials meet in the leader's registers through distributed shared memory in rank
// order: one launch, no scratch, no atomics, a fixed summation order (bit-identical from run to run).
constexpr int kCuThreads = 128;
template <bool VEC>
__global__ void __launch_bounds__(kCuThreads) k_compose_bwd_u(const float* __restrict__ clean, const float* __restrict__ p,
                                                             const float* __restrict__ gx, float* __restrict__ gp, int rows, int T) {
    constexpr int W = VEC ? 4 : 1;
    __shared__ float4 part[kCuThreads];
    cg::cluster_group cl = cg::this_cluster();
    const int K = (int)gridDim.y, rank = (int)blockIdx.y;       // the cluster spans the whole y extent
    const int cols = (T + W - 1) / W;
    const int c = blockIdx.x * kCuThreads + threadIdx.x;
    const bool on = c < cols;
    const size_t col = (size_t)(on ? c : 0) * W;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
        if (VEC) {
            constexpr int U = 8;
            const float4 q = ld4(p + col);
            for (int b0 = rank; b0 < rows; b0 += U * K) {
                float4 x[U], g[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (b0 + u * K < rows) {
                        const size_t i = (size_t)(b0 + u * K) * T + col;
                        x[u] = ld4_stream(clean + i); g[u] = ld4_stream(gx + i);
                    }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (b0 + u * K < rows) {
                        acc.x += pass1(x[u].x, q.x, g[u].x); acc.y += pass1(x[u].y, q.y, g[u].y);
                        acc.z += pass1(x[u].z, q.z, g[u].z); acc.w += pass1(x[u].w, q.w, g[u].w);
                    }
            }
        } else {
            const float q = p[col];
            for (int b = rank; b < rows; b += K) acc.x += pass1(clean[(size_t)b * T + col], q, gx[(size_t)b * T + col]);
        }
    }
    if (K > 1) {
        if (rank != 0) part[threadIdx.x] = acc;
        cl.sync();
        if (rank == 0)
            for (int r = 1; r < K; ++r) {
                const float4 o = *cl.map_shared_rank(&part[threadIdx.x], r);
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
        cl.sync();                                             // the leader has read every partial: the others may exit
    }
    if (on && rank == 0) {
        if (VEC) st4(gp + col, acc); else gp[col] = acc.x;
    }
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }
inline int grid_for(const paa_handle* h, int64_t n_vec) {
    int64_t want = (n_vec + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)h->num_sms * kBlocksPerSm;
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

// template value of a launch: the step mode, plus kStepParts when the gradient has to be summed from several parts
inline int step_code(int mode, const StepDev& sd) { return mode | ((mode != PAA_STEP_NONE && sd.nparts > 1) ? kStepParts : 0); }
// every gradient buffer (the one, or all parts) 16-byte aligned
inline bool grads_aligned(const StepDev& sd) {
    if (!aligned16(sd.grad)) return false;
    for (int k = 1; k < sd.nparts; ++k) if (!aligned16(sd.gpart[k])) return false;
    return true;
}

template <int STEP>
int launch_step_clamp(paa_handle* h, const float* p_in, float* p_out, int64_t n, float lo, float hi, bool clamp,
                      const StepDev& sd, cudaStream_t st) {
    bool vec = aligned16(p_in) && aligned16(p_out) && ((STEP & 3) == PAA_STEP_NONE || grads_aligned(sd)) &&
               ((STEP & 3) != PAA_STEP_ADAM || (aligned16(sd.m) && aligned16(sd.v)));
    int grid = grid_for(h, vec ? (n + 3) / 4 : n);
    if (vec) k_step_clamp<STEP, true><<<grid, kThreads, 0, st>>>(p_in, p_out, n, lo, hi, clamp, sd);
    else k_step_clamp<STEP, false><<<grid, kThreads, 0, st>>>(p_in, p_out, n, lo, hi, clamp, sd);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

template <int NORM, int STEP>
int launch_reduce(paa_handle* h, const ReduceArgs& a, const StepDev& sd, int grid, bool vec, cudaStream_t st) {
    if (vec) k_reduce<NORM, STEP, true><<<grid, kThreads, 0, st>>>(a, sd);
    else k_reduce<NORM, STEP, false><<<grid, kThreads, 0, st>>>(a, sd);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

// The cooperative kernel of one (NORM, STEP, RIDE): its dynamic shared-memory ceiling is raised and its co-residency queried
// once per device (not per call: small shapes are launch-latency bound), the answers cached per instantiation.
constexpr int kMaxDevices = 64;
template <int NORM, int STEP, bool RIDE, bool PARKED>
cudaError_t fused_kernel_r(const paa_handle* h, void** kern, int* blocks_per_sm) {
    static int cached[kMaxDevices];          // 0 = unknown, else 1 + blocks per SM
    void* k = (void*)k_fused<NORM, STEP, RIDE, PARKED>;
    *kern = k;
    const int dev = h->device & (kMaxDevices - 1);
    int c = __atomic_load_n(&cached[dev], __ATOMIC_ACQUIRE);
    if (c == 0) {
        const size_t smem_max = (size_t)kSCMax * kFT * sizeof(float4);
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        int bps = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, kFT, smem_max);
        if (e != cudaSuccess) return e;
        c = 1 + bps;
        __atomic_store_n(&cached[dev], c, __ATOMIC_RELEASE);
    }
    *blocks_per_sm = c - 1;
    return cudaSuccess;
}

// tv always rides; snr when asked to (clean longer than the perturbation); l2 has no clean tensor
template <int NORM, int STEP, bool PARKED>
cudaError_t fused_kernel_p(const paa_handle* h, bool ride, void** kern, int* blocks_per_sm) {
    if (NORM == NORM_TV) return fused_kernel_r<NORM, STEP, NORM == NORM_TV, PARKED>(h, kern, blocks_per_sm);
    if (NORM == NORM_SNR && ride) return fused_kernel_r<NORM, STEP, NORM == NORM_SNR, PARKED>(h, kern, blocks_per_sm);
    return fused_kernel_r<NORM, STEP, false, PARKED>(h, kern, blocks_per_sm);
}
template <int NORM, int STEP>
cudaError_t fused_kernel(const paa_handle* h, bool ride, bool parked, void** kern, int* blocks_per_sm) {
    return parked ? fused_kernel_p<NORM, STEP, true>(h, ride, kern, blocks_per_sm)
                  : fused_kernel_p<NORM, STEP, false>(h, ride, kern, blocks_per_sm);
}

template <int NORM>
int project_reduce(paa_handle* h, const float* p_in, float* p_out, int rows, int T, const float* clean,
                   int64_t clean_n, int clean_T, float eps, double snr_linear, const paa_step* step, void* scratch,
                   cudaStream_t st) {
    if (!h || !p_in || !p_out || !scratch) return PAA_ERR_NULL;
    if (rows <= 0 || T <= 0) return PAA_ERR_SHAPE;
    // mode U: the clean statistics come from the per-rank parts, the clean audio itself is not read
    const paa_parts* parts = step ? step->parts : nullptr;
    const bool stats = NORM != NORM_L2 && parts && parts->n > 0 && parts->clean_stats[0];
    if (stats) {
        if (parts->n > PAA_MAX_PARTS || parts->clean_numel < 0) return PAA_ERR_SHAPE;
        for (int k = 0; k < parts->n; ++k) if (!parts->clean_stats[k]) return PAA_ERR_NULL;
        clean = nullptr; clean_n = std::max<int64_t>(parts->clean_numel, 1); clean_T = 1;
    }
    if (NORM != NORM_L2 && !clean && !stats) return PAA_ERR_NEED_CLEAN;
    if (NORM != NORM_L2 && (clean_n <= 0 || clean_T <= 0)) return PAA_ERR_SHAPE;
    int mode = 0;
    StepDev sd{};
    int rc = paa_make_step(step, &mode, &sd);
    if (rc) return rc;
    const int64_t n = (int64_t)rows * T;
    const float* src = p_in;
    if (NORM == NORM_TV && mode != PAA_STEP_NONE) {
        if (p_in == p_out) return PAA_ERR_ALIAS;
        if (mode == PAA_STEP_ADAM) {       // neighbour recomputation would need the pre-update state
            rc = paa_launch_adam_prepass(h, p_in, p_out, n, sd, st);
            if (rc) return rc;
            src = p_out;
            mode = PAA_STEP_NONE;
        }
    }
    ReduceArgs a{};
    a.p_in = src; a.q_out = p_out; a.n = n; a.T = T;
    a.clean = clean; a.clean_n = (NORM == NORM_L2 || stats) ? 0 : clean_n; a.clean_T = clean_T > 0 ? clean_T : 1;
    a.partials = scratch_partials(scratch);
    a.write_q = (mode != PAA_STEP_NONE) || (src != p_out);
    bool vec = aligned16(src) && aligned16(p_out) && (NORM == NORM_L2 || stats || aligned16(clean)) &&
               (NORM != NORM_TV || (T >= 4 && (stats || clean_T >= 4))) &&
               (mode == PAA_STEP_NONE || grads_aligned(sd)) &&
               (mode != PAA_STEP_ADAM || (aligned16(sd.m) && aligned16(sd.v)));
    int64_t work = std::max<int64_t>(n, a.clean_n);
    FinalArgs f{};
    f.partials = a.partials; f.scalars = scratch_scalars(scratch);
    f.eps = eps; f.snr_linear = snr_linear; f.n_p = (double)n; f.n_clean = (double)clean_n;
    if (stats) {
        f.ncstat = parts->n; f.cstat_idx = NORM == NORM_SNR ? 0 : 1;
        f.numel_from_stats = parts->clean_numel == 0;
        for (int k = 0; k < parts->n; ++k) f.cstat[k] = parts->clean_stats[k];
    }
    // k_fused indexes float4s with 32 bits (2^31 float4s = 8.6 G elements); beyond that the three-kernel form runs
    const bool fits32 = (work >> 2) < ((int64_t)1 << 31) - ((int64_t)1 << 22);
    if (vec && !h->no_coop && fits32) {
        // single cooperative launch; fall through to the three-kernel form only if the device refuses it
        void* kern = nullptr;
        int bps = 0;
        cudaError_t e = cudaSuccess;
        const bool ride = a.clean_n > n;
        auto pick = [&](bool parked, void** k, int* b) {
            switch (step_code(mode, sd)) {
                case PAA_STEP_NONE: return fused_kernel<NORM, PAA_STEP_NONE>(h, ride, parked, k, b);
                case PAA_STEP_PGD: return fused_kernel<NORM, PAA_STEP_PGD>(h, ride, parked, k, b);
                case PAA_STEP_ADAM: return fused_kernel<NORM, PAA_STEP_ADAM>(h, ride, parked, k, b);
                case PAA_STEP_PGD | kStepParts: return fused_kernel<NORM, PAA_STEP_PGD | kStepParts>(h, ride, parked, k, b);
                default: return fused_kernel<NORM, PAA_STEP_ADAM | kStepParts>(h, ride, parked, k, b);
            }
        };
        e = pick(true, &kern, &bps);
        if (e != cudaSuccess) return paa_cuda_fail(h, e);
        if (bps > 0) {
            const int max_grid = bps * h->num_sms;              // co-residency as the occupancy calculator reports it
            const int64_t work4 = (work + 3) / 4;
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(max_grid, (work4 + kFT - 1) / kFT));
            const int64_t iters = ((n >> 2) + (int64_t)grid * kFT - 1) / ((int64_t)grid * kFT);
            int sc_iters = (int)std::max<int64_t>(0, std::min<int64_t>(kSCMax, iters - (mode == PAA_STEP_ADAM ? 2 : kRC)));
            if (mode != PAA_STEP_ADAM && iters <= kRC + kSCMax) {          // nothing parked in HBM: the lean form, if it is as resident
                void* lean = nullptr;
                int lbps = 0;
                if (pick(false, &lean, &lbps) == cudaSuccess && lbps >= bps) kern = lean;
            }
            size_t smem = (size_t)sc_iters * kFT * sizeof(float4);
            f.nblocks = grid;
            void* params[] = {(void*)&a, (void*)&sd, (void*)&f, (void*)&sc_iters};
            e = cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kFT), params, smem, st);
            if (e == cudaSuccess) {
                __atomic_add_fetch(&g_paa_launches, 1, __ATOMIC_RELAXED);
                return PAA_OK;
            }
            (void)cudaGetLastError();
            // only "this device / configuration cannot launch cooperatively" selects the three-kernel form for good
            if (e != cudaErrorCooperativeLaunchTooLarge && e != cudaErrorNotSupported) return paa_cuda_fail(h, e);
        }
        h->last_cuda_error = (int)e;
        h->no_coop = 1;
    }
    int grid = std::min(grid_for(h, vec ? (work + 3) / 4 : work), kMaxPartialBlocks);
    switch (step_code(mode, sd)) {
        case PAA_STEP_NONE: rc = launch_reduce<NORM, PAA_STEP_NONE>(h, a, sd, grid, vec, st); break;
        case PAA_STEP_PGD: rc = launch_reduce<NORM, PAA_STEP_PGD>(h, a, sd, grid, vec, st); break;
        case PAA_STEP_ADAM: rc = launch_reduce<NORM, PAA_STEP_ADAM>(h, a, sd, grid, vec, st); break;
        case PAA_STEP_PGD | kStepParts: rc = launch_reduce<NORM, PAA_STEP_PGD | kStepParts>(h, a, sd, grid, vec, st); break;
        default: rc = launch_reduce<NORM, PAA_STEP_ADAM | kStepParts>(h, a, sd, grid, vec, st); break;
    }
    if (rc) return rc;
    f.nblocks = grid;
    k_finalize<NORM><<<1, kThreads, 0, st>>>(f);
    PAA_LAUNCH_CHECK(h);
    bool vb = aligned16(p_out);
    int gridb = grid_for(h, vb ? (n + 3) / 4 : n);
    if (vb) k_scale<true><<<gridb, kThreads, 0, st>>>(p_out, p_out, n, f.scalars);
    else k_scale<false><<<gridb, kThreads, 0, st>>>(p_out, p_out, n, f.scalars);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

}  // namespace

// Host-side folding of the Adam scalars, in double like torch/optim/adam.py:531-547, rounded to
// fp32 only where torch applies them to fp32 tensors.
int paa_make_step(const paa_step* step, int* mode, StepDev* out) {
    *mode = PAA_STEP_NONE;
    if (!step || step->mode == PAA_STEP_NONE) return PAA_OK;
    if (step->mode != PAA_STEP_PGD && step->mode != PAA_STEP_ADAM) return PAA_ERR_UNSUPPORTED;
    out->nparts = 0;
    const paa_parts* parts = step->parts;
    if (parts && parts->n > 0 && parts->grad[0]) {          // mode U: the gradient is the sum of the parts
        if (parts->n > PAA_MAX_PARTS) return PAA_ERR_SHAPE;
        for (int k = 0; k < parts->n; ++k) {
            if (!parts->grad[k]) return PAA_ERR_NULL;
            out->gpart[k] = parts->grad[k];
        }
        out->nparts = parts->n;
        out->grad = parts->grad[0];
    } else {
        if (!step->grad) return PAA_ERR_NULL;
        out->grad = step->grad;
    }
    out->lr = (float)step->lr;
    if (step->mode == PAA_STEP_ADAM) {
        if (!step->adam_m || !step->adam_v) return PAA_ERR_NULL;
        if (step->adam_t < 1) return PAA_ERR_SHAPE;
        const double b1 = step->beta1, b2 = step->beta2;
        const double bc1 = 1.0 - std::pow(b1, (double)step->adam_t);
        const double bc2 = 1.0 - std::pow(b2, (double)step->adam_t);
        out->m = step->adam_m; out->v = step->adam_v;
        out->w1 = (float)(1.0 - b1);
        out->beta2 = (float)b2;
        out->w2 = (float)(1.0 - b2);
        out->neg_step = (float)(-(step->lr / bc1));
        out->bc2_sqrt = (float)std::sqrt(bc2);
        out->eps = (float)step->eps;
    }
    *mode = step->mode;
    return PAA_OK;
}

int paa_launch_adam_prepass(paa_handle* h, const float* p_in, float* p_out, int64_t n, const StepDev& sd, cudaStream_t st) {
    PaaDeviceGuard device_guard(h);
    if (sd.nparts > 1) return launch_step_clamp<PAA_STEP_ADAM | kStepParts>(h, p_in, p_out, n, 0.f, 0.f, false, sd, st);
    return launch_step_clamp<PAA_STEP_ADAM>(h, p_in, p_out, n, 0.f, 0.f, false, sd, st);
}

int paa_launch_sum_parts(paa_handle* h, const StepDev& sd, float* out, int64_t n, cudaStream_t st) {
    PaaDeviceGuard device_guard(h);
    const bool vec = grads_aligned(sd) && aligned16(out);
    k_sum_parts<<<grid_for(h, vec ? (n + 3) / 4 : n), kThreads, 0, st>>>(sd, out, n, vec);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

extern "C" {

int paa_clean_stats(paa_handle* h, const float* clean, int rows, int T, double* out3, void* scratch, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !clean || !out3 || !scratch) return PAA_ERR_NULL;
    if (rows <= 0 || T <= 0) return PAA_ERR_SHAPE;
    const int64_t n = (int64_t)rows * T;
    const int grid = std::min(grid_for(h, n), kMaxPartialBlocks);
    cudaStream_t st = (cudaStream_t)stream;
    k_clean_stats<<<grid, kThreads, 0, st>>>(clean, n, T, scratch_partials(scratch));
    PAA_LAUNCH_CHECK(h);
    k_clean_stats_final<<<1, kThreads, 0, st>>>(scratch_partials(scratch), grid, out3, (double)n);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

int paa_step_only(paa_handle* h, const float* p_in, float* p_out, int rows, int T, const paa_step* step, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !p_in || !p_out) return PAA_ERR_NULL;
    if (rows <= 0 || T <= 0) return PAA_ERR_SHAPE;
    int mode = 0;
    StepDev sd{};
    int rc = paa_make_step(step, &mode, &sd);
    if (rc) return rc;
    const int64_t n = (int64_t)rows * T;
    cudaStream_t st = (cudaStream_t)stream;
    switch (step_code(mode, sd)) {
        case PAA_STEP_PGD: return launch_step_clamp<PAA_STEP_PGD>(h, p_in, p_out, n, 0.f, 0.f, false, sd, st);
        case PAA_STEP_ADAM: return launch_step_clamp<PAA_STEP_ADAM>(h, p_in, p_out, n, 0.f, 0.f, false, sd, st);
        case PAA_STEP_PGD | kStepParts: return launch_step_clamp<PAA_STEP_PGD | kStepParts>(h, p_in, p_out, n, 0.f, 0.f, false, sd, st);
        case PAA_STEP_ADAM | kStepParts: return launch_step_clamp<PAA_STEP_ADAM | kStepParts>(h, p_in, p_out, n, 0.f, 0.f, false, sd, st);
        default: return launch_step_clamp<PAA_STEP_NONE>(h, p_in, p_out, n, 0.f, 0.f, false, sd, st);
    }
}

int paa_project_linf(paa_handle* h, const float* p_in, float* p_out, int rows, int T, double lo, double hi,
                     const paa_step* step, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !p_in || !p_out) return PAA_ERR_NULL;
    if (rows <= 0 || T <= 0) return PAA_ERR_SHAPE;
    int mode = 0;
    StepDev sd{};
    int rc = paa_make_step(step, &mode, &sd);
    if (rc) return rc;
    const int64_t n = (int64_t)rows * T;
    cudaStream_t st = (cudaStream_t)stream;
    switch (step_code(mode, sd)) {
        case PAA_STEP_PGD: return launch_step_clamp<PAA_STEP_PGD>(h, p_in, p_out, n, (float)lo, (float)hi, true, sd, st);
        case PAA_STEP_ADAM: return launch_step_clamp<PAA_STEP_ADAM>(h, p_in, p_out, n, (float)lo, (float)hi, true, sd, st);
        case PAA_STEP_PGD | kStepParts: return launch_step_clamp<PAA_STEP_PGD | kStepParts>(h, p_in, p_out, n, (float)lo, (float)hi, true, sd, st);
        case PAA_STEP_ADAM | kStepParts: return launch_step_clamp<PAA_STEP_ADAM | kStepParts>(h, p_in, p_out, n, (float)lo, (float)hi, true, sd, st);
        default: return launch_step_clamp<PAA_STEP_NONE>(h, p_in, p_out, n, (float)lo, (float)hi, true, sd, st);
    }
}

int paa_project_l2(paa_handle* h, const float* p_in, float* p_out, int rows, int T, double epsilon,
                   const paa_step* step, void* scratch, void* stream) {
    PaaDeviceGuard device_guard(h);
    return project_reduce<NORM_L2>(h, p_in, p_out, rows, T, nullptr, 0, 1, (float)epsilon, 0.0, step, scratch,
                                   (cudaStream_t)stream);
}

int paa_project_snr(paa_handle* h, const float* p_in, float* p_out, int rows, int T, const float* clean,
                    int64_t clean_numel, double snr_db, const paa_step* step, void* scratch, void* stream) {
    PaaDeviceGuard device_guard(h);
    return project_reduce<NORM_SNR>(h, p_in, p_out, rows, T, clean, clean_numel, 1, (float)snr_db,
                                    std::pow(10.0, snr_db / 10.0), step, scratch, (cudaStream_t)stream);
}

int paa_project_tv(paa_handle* h, const float* p_in, float* p_out, int rows, int T, const float* clean,
                   int clean_rows, int clean_T, double tv_epsilon, const paa_step* step, void* scratch, void* stream) {
    PaaDeviceGuard device_guard(h);
    return project_reduce<NORM_TV>(h, p_in, p_out, rows, T, clean, (int64_t)clean_rows * clean_T, clean_T,
                                   (float)tv_epsilon, 0.0, step, scratch, (cudaStream_t)stream);
}

static dim3 compose_grid(const paa_handle* h, int cols, int rows, int rows_per_trip) {
    const int gx = std::max(1, std::min((cols + kThreads - 1) / kThreads, h->num_sms * kBlocksPerSm));
    const int gy = std::max(1, std::min(rows, rows_per_trip > 0 ? (rows + rows_per_trip - 1) / rows_per_trip
                                                                  : (h->num_sms * kBlocksPerSm + gx - 1) / gx));
    return dim3(gx, gy, 1);
}

int paa_compose_clamp(paa_handle* h, const float* clean, int clean_rows, const float* p, int p_rows, int T,
                      float* x_adv, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !clean || !p || !x_adv) return PAA_ERR_NULL;
    if (clean_rows <= 0 || T <= 0 || (p_rows != 1 && p_rows != clean_rows)) return PAA_ERR_SHAPE;
    const bool vec = aligned16(clean) && aligned16(p) && aligned16(x_adv) && (T % 4 == 0);
    const dim3 grid = compose_grid(h, vec ? T / 4 : T, clean_rows, 4);
    if (vec) k_compose<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(clean, p, x_adv, clean_rows, p_rows, T);
    else k_compose<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(clean, p, x_adv, clean_rows, p_rows, T);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

}  // extern "C"

template <bool VEC>
static int launch_compose_bwd_u(paa_handle* h, const float* clean, const float* p, const float* gx, float* gp, int rows, int T,
                                cudaStream_t st) {
    const int cols = VEC ? T / 4 : T;
    const int gxb = (cols + kCuThreads - 1) / kCuThreads;
    // measured on B200 (tools/ab_time.py, 32 x 10 s / 128 x 10 s): K = 2: 11.1 / 28.8 us, K = 4: 13.7 / 30.3, K = 8: 18.5 / 34.0
    // (the cluster barrier grows with K), the single-CTA column loop of round 1: 16.4 / 51.1
    const int K = rows >= 2 ? 2 : 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(gxb, K, 1);
    cfg.blockDim = dim3(kCuThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = K; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PAA_CUDA(h, cudaLaunchKernelEx(&cfg, k_compose_bwd_u<VEC>, clean, p, gx, gp, rows, T));
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

extern "C" {

int paa_compose_clamp_backward(paa_handle* h, const float* clean, int clean_rows, const float* p, int p_rows, int T,
                               const float* grad_x_adv, float* grad_p, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !clean || !p || !grad_x_adv || !grad_p) return PAA_ERR_NULL;
    if (clean_rows <= 0 || T <= 0 || (p_rows != 1 && p_rows != clean_rows)) return PAA_ERR_SHAPE;
    const bool vec = aligned16(clean) && aligned16(p) && aligned16(grad_x_adv) && aligned16(grad_p) && (T % 4 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (p_rows == 1 && clean_rows > 1)
        return vec ? launch_compose_bwd_u<true>(h, clean, p, grad_x_adv, grad_p, clean_rows, T, st)
                   : launch_compose_bwd_u<false>(h, clean, p, grad_x_adv, grad_p, clean_rows, T, st);
    const dim3 grid = compose_grid(h, vec ? T / 4 : T, clean_rows, kComposeBwdRows);
    if (vec) k_compose_bwd<true><<<grid, kThreads, 0, st>>>(clean, p, grad_x_adv, grad_p, clean_rows, T);
    else k_compose_bwd<false><<<grid, kThreads, 0, st>>>(clean, p, grad_x_adv, grad_p, clean_rows, T);
    PAA_LAUNCH_CHECK(h);
    return PAA_OK;
}

}  // extern "C"
