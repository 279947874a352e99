// Internal declarations shared by the translation units of libpaa.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "../../include/paa.h"

#define PAA_VERSION 200

// Host tables of one (n_fft, hop, sr) plan, mirrored on the device.
struct paa_handle {
    int device = 0;
    int n_fft = 0, hop = 0, sr = 0;
    int F = 0;               // n_fft/2 + 1
    int R = 0;               // n_fft / hop: frames overlapping one output sample
    int num_sms = 148;
    float bin_hz = 0.f;      // fp32(1/(n_fft*(1/sr))): what torch.fft.rfftfreq multiplies arange by
    mutable int last_cuda_error = 0;
    mutable int nola_frames = -1, nola_result = 0;   // last window-envelope check (paa_stft.cu: nola_ok)
    int no_coop = 0;         // set if the device refused a cooperative launch: use the three-kernel form

    // one device blob, copied into shared memory by a 1-D TMA bulk copy at kernel start:
    //   [twiddles (per-lane, stages 1..2) float4][split twiddles N/2+1 float2][window/2 n_fft f32]
    void* d_blob = nullptr;
    size_t blob_bytes = 0;
    size_t off_twiddle = 0, off_post = 0, off_window = 0;
    std::vector<float> h_window;

    // half-warp kernel (n_fft 1024, hop 256; paa_stft.cu k_stft_hw): [W_512^{l k1} float2[32][16] | window[0..511] |
    // reciprocal envelope[hop]] is the part copied into shared memory (blob_hw_smem bytes); the full window follows
    void* d_blob_hw = nullptr;
    size_t blob_hw_smem = 0, off_hw_window = 0;
    int use_hw = 0;          // 0 when the geometry has no half-warp kernel or PAA_STFT_HW=0 (A/B measurements)

    // fletcher_munson penalty grid, frequency axis pre-interpolated per rfft bin
    float* d_fm_blob = nullptr;      // [64: phon knots][n_phon rows of paa_fm_stride(F): w(knot i, f_k)], fill outside the frequency axis
    size_t fm_blob_bytes = 0;
    int fm_n_phon = 0;
    float fm_fill = 1.f;
    int fm_uniform = 0;              // knots equally spaced -> direct cell lookup
    float fm_k0 = 0.f, fm_klast = 0.f, fm_inv_dk = 0.f;
};

// Row stride of the fletcher_munson table: a multiple of 32 floats, so that the shared-memory bank of w(knot i, f_k) is
// k mod 32 = the lane whatever knot row i a lane's level selects (F = 513 itself would make it (i + k) mod 32: two-way
// conflicts wherever neighbouring bins fall into different 10-phon cells).
constexpr int paa_fm_stride(int F) { return (F + 31) / 32 * 32; }

// ---- status helpers -------------------------------------------------------------------------
static inline int paa_cuda_fail(const paa_handle* h, cudaError_t e) {
    if (h) h->last_cuda_error = (int)e;
    return PAA_ERR_CUDA;
}
#define PAA_CUDA(h, call)                                              \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return paa_cuda_fail((h), e__);        \
    } while (0)
extern long long g_paa_launches;       // bumped by every kernel launch (paa_launch_count)
#define PAA_LAUNCH_CHECK(h)                        \
    do {                                           \
        __atomic_add_fetch(&g_paa_launches, 1, __ATOMIC_RELAXED); \
        PAA_CUDA((h), cudaGetLastError());         \
    } while (0)

// Entry points launch on the handle's device whatever the caller's current device is, and put it back.
struct PaaDeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit PaaDeviceGuard(const paa_handle* h) {
        if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device) switched = cudaSetDevice(h->device) == cudaSuccess;
    }
    ~PaaDeviceGuard() { if (switched) cudaSetDevice(prev); }
    PaaDeviceGuard(const PaaDeviceGuard&) = delete;
    PaaDeviceGuard& operator=(const PaaDeviceGuard&) = delete;
};

// ---- scratch layout (bytes) -----------------------------------------------------------------
// [0,256)            float scalars[PAA_S_COUNT..]           (PAA_S_* indices)
// [256, 256+P)       double partials[kMaxBlocks][2]
// [.., +rows*T*4)    staging buffer for STFT-domain paths (stepped perturbation)
constexpr int kScalarBytes = 256;
constexpr int kMaxPartialBlocks = 8192;
constexpr size_t kPartialBytes = (size_t)kMaxPartialBlocks * 2 * sizeof(double);
static inline float* scratch_scalars(void* s) { return (float*)s; }
static inline double* scratch_partials(void* s) { return (double*)((char*)s + kScalarBytes); }
static inline float* scratch_stage(void* s) { return (float*)((char*)s + kScalarBytes + kPartialBytes); }
// mode U, STFT-domain projections: the summed gradient, after the [rows, T] staging buffer (256-byte aligned)
static inline float* scratch_gsum(void* s, int rows, int T) {
    const size_t stage = ((size_t)rows * (size_t)T * sizeof(float) + 255) / 256 * 256;
    return (float*)((char*)s + kScalarBytes + kPartialBytes + stage);
}

// fletcher_munson pass A: one fp64 partial per tile, behind the two [rows, T] buffers (no cap on rows x tiles)
static inline size_t scratch_fm_tiles(int rows, int T, int hop) {
    const size_t frames = 1 + (size_t)T / (size_t)(hop > 0 ? hop : 1);
    return (size_t)rows * ((frames + 15) / 16 + 1);            // a tile holds >= 16 frames
}
static inline double* scratch_fm_partials(void* s, int rows, int T) {
    const size_t stage = ((size_t)rows * (size_t)T * sizeof(float) + 255) / 256 * 256;
    return (double*)((char*)s + kScalarBytes + kPartialBytes + 2 * stage);
}

// ---- step parameters as the kernels see them -------------------------------------------------
struct StepDev {
    const float* grad;     // == gpart[0] in mode U
    const float* gpart[PAA_MAX_PARTS];   // mode U: per-rank partial gradients, summed in index order by the kernels
    int nparts;            // 0 or 1: `grad` alone
    float* m;
    float* v;
    float lr;          // PGD
    float w1;          // Adam: fp32(1-beta1)  (lerp weight)
    float beta2;       // Adam
    float w2;          // Adam: fp32(1-beta2)
    float neg_step;    // Adam: fp32(-(lr/bias_correction1))
    float bc2_sqrt;    // Adam: fp32(sqrt(bias_correction2))
    float eps;         // Adam
};
int paa_make_step(const paa_step* step, int* mode, StepDev* out);

// time-domain launches (paa_time.cu)
int paa_launch_adam_prepass(paa_handle* h, const float* p_in, float* p_out, int64_t n, const StepDev& sd, cudaStream_t st);
// mode U: out[i] = sum over parts of gpart[k][i] (index order), for paths that cannot sum on the fly
int paa_launch_sum_parts(paa_handle* h, const StepDev& sd, float* out, int64_t n, cudaStream_t st);
