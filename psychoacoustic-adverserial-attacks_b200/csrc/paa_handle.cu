// Handle life cycle of libpaa.so: immutable device tables for one (device, n_fft, hop, sr).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "paa_fft.cuh"
#include "paa_internal.h"

namespace {

// torch.hann_window(n, periodic=True) in fp32, step by step as ATen builds it:
// arange(n+1) * fp32(2*pi/n) -> cos -> * -0.5 -> + 0.5 (first n entries).
std::vector<float> hann_periodic(int n) {
    std::vector<float> w(n);
    const float step = (float)(M_PI * 2.0 / (double)n);
    for (int i = 0; i < n; ++i) {
        float a = (float)i * step;
        w[i] = cosf(a) * -0.5f + 0.5f;
    }
    return w;
}

// Per-lane Stockham twiddles: stage with radix R after Ns points: W_{Ns*R}^{(j mod Ns) * r},
// j = lane + 32*b, stored [b][q][lane] as float4 = forward twiddles (cos, -sin) of r = 2q and 2q+1.
void stage_twiddles(std::vector<float>& out, int N, int R, int Ns) {
    const int NB = (32 % Ns == 0) ? 1 : N / R / 32;       // k = j mod Ns is the same for every b when Ns | 32
    for (int b = 0; b < NB; ++b)
        for (int q = 0; q < 4; ++q)
            for (int lane = 0; lane < 32; ++lane)
                for (int r = 2 * q; r < 2 * q + 2; ++r) {
                    const int k = (lane + 32 * b) % Ns;
                    const double th = 2.0 * M_PI * (double)k * (double)(r % R) / (double)(Ns * R);
                    out.push_back((float)std::cos(th));
                    out.push_back((float)(-std::sin(th)));
                }
}

template <int NFFT>
void build_tables(std::vector<float>& tw, std::vector<float>& post) {
    using P = paa::Plan<NFFT>;
    stage_twiddles(tw, P::N, P::R1, P::R0);
    // stage 2, compact: W_N^{k e} for e = 1, 2, 4 as float2 [block][e][lane]; blocks: b = 0 .. NB2-1 (k = lane + 32 b)
    // and, for n_fft 1024, the paired butterfly k = j1(lane) (paa_fft.cuh)
    static_assert(P::R2 == 8, "the compact stage-2 table assumes a radix-8 last stage");
    auto block = [&](auto kof) {
        for (int e : {1, 2, 4})
            for (int lane = 0; lane < 32; ++lane) {
                const double th = 2.0 * M_PI * (double)kof(lane) * (double)e / (double)P::N;
                tw.push_back((float)std::cos(th));
                tw.push_back((float)(-std::sin(th)));
            }
    };
    for (int b = 0; b < P::N / P::R2 / 32; ++b) block([b](int lane) { return lane + 32 * b; });
    if (NFFT == 1024) block([](int lane) { return paa::paired_j1(lane); });
    for (int k = 0; k <= P::N / 2; ++k) {
        const double th = 2.0 * M_PI * (double)k / (double)NFFT;
        post.push_back((float)std::cos(th));
        post.push_back((float)std::sin(th));
    }
}

size_t round16(size_t b) { return (b + 15) / 16 * 16; }

}  // namespace

extern "C" {

int paa_create(int device, int n_fft, int hop, int sr, paa_handle** out) {
    if (!out) return PAA_ERR_NULL;
    *out = nullptr;
    if (n_fft != 512 && n_fft != 1024) return PAA_ERR_UNSUPPORTED;
    if (hop <= 0 || sr <= 0) return PAA_ERR_SHAPE;
    // frames must tile: hop divides n_fft, at least 50 % overlap, float4-aligned frame starts
    if (n_fft % hop != 0 || n_fft / hop < 2 || hop % 4 != 0 || n_fft / hop > 16) return PAA_ERR_UNSUPPORTED;
    paa_handle* h = new paa_handle();
    h->device = device; h->n_fft = n_fft; h->hop = hop; h->sr = sr;
    h->F = n_fft / 2 + 1; h->R = n_fft / hop;
    h->bin_hz = (float)(1.0 / ((double)n_fft * (1.0 / (double)sr)));
    h->h_window = hann_periodic(n_fft);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        int sms = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess && sms > 0) h->num_sms = sms;
    }
    if (e != cudaSuccess) { int rc = paa_cuda_fail(h, e); delete h; return rc; }

    std::vector<float> tw, post;
    if (n_fft == 1024) build_tables<1024>(tw, post); else build_tables<512>(tw, post);
    // blob = [twiddles | split twiddles | window/2]: the first two stay in shared memory for the whole kernel, the
    // window only passes through (paa_stft.cu)
    h->off_twiddle = 0;
    h->off_post = round16(tw.size() * 4);
    h->off_window = round16(h->off_post + post.size() * 4);
    h->blob_bytes = round16(h->off_window + (size_t)n_fft * 4);
    std::vector<unsigned char> blob(h->blob_bytes, 0);
    {   // the device table holds w/2 (exact): the real-FFT split then needs no 0.5, and 2/n_fft is folded into the per-bin op
        std::vector<float> half(h->h_window);
        for (float& v : half) v *= 0.5f;
        std::memcpy(blob.data() + h->off_window, half.data(), (size_t)n_fft * 4);
    }
    std::memcpy(blob.data() + h->off_twiddle, tw.data(), tw.size() * 4);
    std::memcpy(blob.data() + h->off_post, post.data(), post.size() * 4);
    e = cudaMalloc(&h->d_blob, h->blob_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_blob, blob.data(), h->blob_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { int rc = paa_cuda_fail(h, e); paa_destroy(h); return rc; }
    if (n_fft == 1024 && hop == 256) {
        // tables of the half-warp kernel: forward twiddles W_512^{l k1} (cos, -sin) as [k1][l], the first half of the
        // true Hann window (the second half is 1 - w), the reciprocal overlap-add envelope of an interior hop block
        std::vector<float> b;
        for (int k1 = 0; k1 < 32; ++k1)
            for (int l = 0; l < 16; ++l) {
                const double th = 2.0 * M_PI * (double)(l * k1) / 512.0;
                b.push_back((float)std::cos(th));
                b.push_back((float)(-std::sin(th)));
            }
        for (int i = 0; i < 512; ++i) b.push_back(h->h_window[i]);
        for (int q = 0; q < hop; ++q) {
            float env = 0.f;
            for (int d = h->R - 1; d >= 0; --d) env = std::fmaf(h->h_window[d * hop + q], h->h_window[d * hop + q], env);
            b.push_back(1.f / env);
        }
        h->blob_hw_smem = b.size() * 4;
        h->off_hw_window = h->blob_hw_smem;
        for (int i = 0; i < n_fft; ++i) b.push_back(h->h_window[i]);
        e = cudaMalloc(&h->d_blob_hw, b.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_blob_hw, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { int rc = paa_cuda_fail(h, e); paa_destroy(h); return rc; }
        const char* env_hw = std::getenv("PAA_STFT_HW");
        h->use_hw = (env_hw && env_hw[0] == '1') ? 1 : 0;      // opt-in: measured slower than k_stft at 8 warps/SM
    }
    *out = h;
    return PAA_OK;
}

int paa_destroy(paa_handle* h) {
    if (!h) return PAA_OK;
    cudaFree(h->d_blob);
    cudaFree(h->d_blob_hw);
    cudaFree(h->d_fm_blob);
    delete h;
    return PAA_OK;
}

int paa_last_cuda_error(const paa_handle* h) { return h ? h->last_cuda_error : 0; }
int paa_num_bins(const paa_handle* h) { return h ? h->F : 0; }
int paa_num_frames(const paa_handle* h, int T) { return (h && T >= 0) ? 1 + T / h->hop : 0; }

size_t paa_scratch_bytes(const paa_handle* h, int rows, int T) {
    if (!h || rows < 0 || T < 0) return 0;
    // scalars + block partials always; the [rows, T] staging buffer only for the STFT-domain projections
    // (twice: mode U keeps the summed gradient of an STFT-domain projection behind the staging buffer)
    // and one fp64 partial per fletcher_munson tile behind them
    const size_t fm = (rows > 0 && T > 0) ? (scratch_fm_tiles(rows, T, h->hop) * sizeof(double) + 255) / 256 * 256 : 0;
    return (size_t)kScalarBytes + kPartialBytes + 2 * (((size_t)rows * (size_t)T * sizeof(float) + 255) / 256 * 256) + fm + 256;
}

int paa_scalars(const paa_handle* h, const void* scratch, float* out8, void* stream) {
    PaaDeviceGuard device_guard(h);
    if (!h || !scratch || !out8) return PAA_ERR_NULL;
    PAA_CUDA(h, cudaMemcpyAsync(out8, scratch, PAA_S_COUNT * sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PAA_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
    return PAA_OK;
}

// Pre-interpolate the (phon x freq) penalty grid along frequency for every rfft bin, in fp64,
// then keep fp32 columns: w(phon_i, f_k) = (1-t) W[i][j] + t W[i][j+1].  Bins whose centre lies
// outside the grid's frequency range take fill_value (RegularGridInterpolator, bounds_error=False).
int paa_set_fm_grid(paa_handle* h, const double* phon_knots, int n_phon, const double* freq_knots, int n_freq,
                    const double* values, double fill_value) {
    PaaDeviceGuard device_guard(h);
    if (!h || !phon_knots || !freq_knots || !values) return PAA_ERR_NULL;
    if (n_phon < 2 || n_freq < 2 || n_phon > 32) return PAA_ERR_SHAPE;      // table must fit shared memory next to the tile
    for (int i = 1; i < n_phon; ++i) if (!(phon_knots[i] > phon_knots[i - 1])) return PAA_ERR_SHAPE;
    for (int j = 1; j < n_freq; ++j) if (!(freq_knots[j] > freq_knots[j - 1])) return PAA_ERR_SHAPE;
    // blob = [64 floats: knots][n_phon rows of paa_fm_stride(F) floats]; rows are contiguous in k so a warp's lanes (consecutive
    // bins) read neighbouring words, and the whole table rides the kernel's TMA bulk copy into shared memory
    const int FS = paa_fm_stride(h->F);
    const size_t blob_floats = ((64 + (size_t)n_phon * FS + 3) / 4) * 4;
    std::vector<float> blob(blob_floats, 0.f);
    for (int k = 0; k < h->F; ++k) {
        const double f = (double)((float)k * h->bin_hz);         // the reference queries with fp32 bin centres
        const bool inband = !(f < freq_knots[0] || f > freq_knots[n_freq - 1]);
        int j = 0;
        while (j + 1 < n_freq - 1 && freq_knots[j + 1] < f) ++j;   // searchsorted(left) - 1, clipped
        const double t = (f - freq_knots[j]) / (freq_knots[j + 1] - freq_knots[j]);
        for (int i = 0; i < n_phon; ++i)
            blob[64 + (size_t)i * FS + k] =
                inband ? (float)((1.0 - t) * values[i * n_freq + j] + t * values[i * n_freq + j + 1]) : (float)fill_value;
    }
    bool uniform = true;
    const double dk = phon_knots[1] - phon_knots[0];
    for (int i = 0; i < n_phon; ++i) {
        blob[i] = (float)phon_knots[i];
        if (i && std::fabs((phon_knots[i] - phon_knots[i - 1]) - dk) > 1e-12 * std::fabs(dk)) uniform = false;
    }
    PAA_CUDA(h, cudaSetDevice(h->device));
    cudaFree(h->d_fm_blob);
    h->d_fm_blob = nullptr;
    h->fm_blob_bytes = blob_floats * sizeof(float);
    PAA_CUDA(h, cudaMalloc((void**)&h->d_fm_blob, h->fm_blob_bytes));
    PAA_CUDA(h, cudaMemcpy(h->d_fm_blob, blob.data(), h->fm_blob_bytes, cudaMemcpyHostToDevice));
    const float* knots = blob.data();
    h->fm_n_phon = n_phon;
    h->fm_fill = (float)fill_value;
    h->fm_uniform = uniform ? 1 : 0;
    h->fm_k0 = knots[0];
    h->fm_klast = knots[n_phon - 1];
    h->fm_inv_dk = (float)(1.0 / dk);
    return PAA_OK;
}

}  // extern "C"
