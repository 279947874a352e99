// Host-side tables and counters of libpaa.so: ISO-226 equal-loudness contours, the
// fletcher_munson penalty grid, per-bin phon thresholds, and word-error counters.
// Follows src/core/iso.py:34-266 and src/training_utils/build.py:325-348 of the reference
// (fp64 throughout, cast to fp32 only where the reference does).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "paa_internal.h"

long long g_paa_launches = 0;

namespace {

constexpr int kBands = 29;
// ISO 226 third-octave parameters (iso.py:60-84)
const double kFreq[kBands] = {20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800,
                              1000, 1250, 1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000, 10000, 12500};
const double kAlpha[kBands] = {.532, .506, .480, .455, .432, .409, .387, .367, .349, .330, .315, .301, .288, .276,
                               .267, .259, .253, .250, .246, .244, .243, .243, .243, .242, .242, .245, .254, .271, .301};
const double kLu[kBands] = {-31.6, -27.2, -23.0, -19.1, -15.9, -13.0, -10.3, -8.1, -6.2, -4.5, -3.1, -2.0, -1.1, -0.4,
                            0.0, 0.3, 0.5, 0.0, -2.7, -4.1, -1.0, 1.7, 2.5, 1.2, -2.1, -7.1, -11.2, -10.7, -3.1};
const double kTf[kBands] = {78.5, 68.7, 59.5, 51.1, 44.0, 37.5, 31.5, 26.5, 22.1, 17.9, 14.4, 11.4, 8.6, 6.2, 4.4,
                            3.0, 2.2, 2.4, 3.5, 1.7, -1.3, -4.2, -6.0, -5.4, -1.5, 6.0, 12.6, 13.9, 12.3};
constexpr double kFLo = 20.0, kFHi = 20000.0;
constexpr int kKnots = kBands + 1;

inline double sgn(double v) { return (v > 0) - (v < 0); }

// Shape-preserving cubic Hermite interpolant with scipy's PCHIP slopes
// (weighted harmonic mean inside, limited three-point formula at the ends).
struct Pchip {
    double x[kKnots], y[kKnots], d[kKnots];

    static double end_slope(double h0, double h1, double m0, double m1) {
        double s = ((2.0 * h0 + h1) * m0 - h0 * m1) / (h0 + h1);
        if (sgn(s) != sgn(m0)) return 0.0;
        if (sgn(m0) != sgn(m1) && std::fabs(s) > 3.0 * std::fabs(m0)) return 3.0 * m0;
        return s;
    }

    // The reference closes the axis at 20 kHz with a copy of the 20 Hz entry (iso.py:104-124).
    explicit Pchip(const double* table) {
        for (int i = 0; i < kBands; ++i) { x[i] = kFreq[i]; y[i] = table[i]; }
        x[kBands] = kFHi;
        y[kBands] = table[0];
        double h[kKnots - 1], m[kKnots - 1];
        for (int i = 0; i < kKnots - 1; ++i) { h[i] = x[i + 1] - x[i]; m[i] = (y[i + 1] - y[i]) / h[i]; }
        for (int k = 1; k < kKnots - 1; ++k) {
            if (m[k - 1] == 0.0 || m[k] == 0.0 || sgn(m[k - 1]) != sgn(m[k])) { d[k] = 0.0; continue; }
            double w1 = 2.0 * h[k] + h[k - 1], w2 = h[k] + 2.0 * h[k - 1];
            d[k] = (w1 + w2) / (w1 / m[k - 1] + w2 / m[k]);
        }
        d[0] = end_slope(h[0], h[1], m[0], m[1]);
        d[kKnots - 1] = end_slope(h[kKnots - 2], h[kKnots - 3], m[kKnots - 2], m[kKnots - 3]);
    }

    double operator()(double q) const {
        int i = int(std::upper_bound(x, x + kKnots, q) - x) - 1;
        i = std::min(std::max(i, 0), kKnots - 2);
        double h = x[i + 1] - x[i], s = q - x[i], m = (y[i + 1] - y[i]) / h;
        double c2 = (3.0 * m - 2.0 * d[i] - d[i + 1]) / h;
        double c3 = (d[i] + d[i + 1] - 2.0 * m) / (h * h);
        return y[i] + s * (d[i] + s * (c2 + s * c3));
    }
};

struct IsoCurves {
    Pchip alpha{kAlpha}, lu{kLu}, tf{kTf};
    double spl(double phon, double f) const {                    // iso.py:163-171
        double al = alpha(f), l = lu(f), t = tf(f);
        double a = 0.00447 * (std::pow(10.0, 0.025 * phon) - 1.15);
        double b = std::pow(0.4 * std::pow(10.0, (t + l) / 10.0 - 9.0), al);
        return (10.0 / al) * std::log10(a + b) - l + 94.0;
    }
};
const IsoCurves& curves() {
    static const IsoCurves c;
    return c;
}

}  // namespace

extern "C" {

const char* paa_status_string(int s) {
    switch (s) {
        case PAA_OK: return "ok";
        case PAA_ERR_NULL: return "null pointer";
        case PAA_ERR_SHAPE: return "bad shape";
        case PAA_ERR_UNSUPPORTED: return "unsupported configuration";
        case PAA_ERR_CUDA: return "CUDA error";
        case PAA_ERR_RANGE: return "value out of the ISO-226 range";
        case PAA_ERR_NEED_CLEAN: return "projection requires clean audio";
        case PAA_ERR_ALIAS: return "output aliases input";
        case PAA_ERR_NOLA: return "window overlap add min: 1";
        case PAA_ERR_STATE: return "handle state incomplete";
        default: return "unknown status";
    }
}

int paa_version(void) { return PAA_VERSION; }
int64_t paa_launch_count(void) { return (int64_t)__atomic_load_n(&g_paa_launches, __ATOMIC_RELAXED); }

int paa_iso226_spl(double phon, const double* freqs_hz, int n, double* out) {
    if (!freqs_hz || !out) return PAA_ERR_NULL;
    if (n < 0) return PAA_ERR_SHAPE;
    if (!(phon >= 0.0 && phon <= 90.0)) return PAA_ERR_RANGE;
    for (int i = 0; i < n; ++i)
        if (freqs_hz[i] < kFLo || freqs_hz[i] > kFHi) return PAA_ERR_RANGE;
    const IsoCurves& c = curves();
    for (int i = 0; i < n; ++i) out[i] = c.spl(phon, freqs_hz[i]);
    return PAA_OK;
}

// iso.py:176-199 (grid) and :202-235 (penalty weights): w = clip((1 - spl/max spl)^2, 0, 1)
int paa_weight_grid(double* phon_knots10, double* freq_knots30, double* w) {
    if (!phon_knots10 || !freq_knots30 || !w) return PAA_ERR_NULL;
    const IsoCurves& c = curves();
    double fk[kKnots];
    for (int j = 0; j < kBands; ++j) fk[j] = kFreq[j];
    fk[kBands] = kFHi;
    double mx = -1e300;
    for (int i = 0; i < 10; ++i) {
        phon_knots10[i] = 10.0 * i;
        for (int j = 0; j < kKnots; ++j) {
            w[i * kKnots + j] = c.spl(10.0 * i, fk[j]);
            mx = std::max(mx, w[i * kKnots + j]);
        }
    }
    for (int j = 0; j < kKnots; ++j) freq_knots30[j] = fk[j];
    for (int i = 0; i < 10 * kKnots; ++i) {
        double t = 1.0 - w[i] / mx;
        w[i] = std::min(std::max(t * t, 0.0), 1.0);
    }
    return PAA_OK;
}

// build.py:325-348: ISO226(max_phon_level)(clip(rfftfreq, 20, 20000)) -> fp32
int paa_spl_thresh(int n_fft, int sr, double max_phon_level, float* out_F) {
    if (!out_F) return PAA_ERR_NULL;
    if (n_fft < 2 || sr <= 0) return PAA_ERR_SHAPE;
    if (!(max_phon_level >= 0.0 && max_phon_level <= 90.0)) return PAA_ERR_RANGE;
    const float step = (float)(1.0 / (n_fft * (1.0 / sr)));      // torch.fft.rfftfreq's fp32 multiplier
    const IsoCurves& c = curves();
    for (int k = 0; k <= n_fft / 2; ++k) {
        float f32 = (float)k * step;
        double f = std::min(std::max((double)f32, kFLo), kFHi);
        // the reference clips an fp32 array, so the ISO evaluation sees fp32-rounded frequencies
        out_F[k] = (float)c.spl(max_phon_level, (double)(float)f);
    }
    return PAA_OK;
}

int paa_interp2(const double* g0, int n0, const double* g1, int n1, const double* v, double fill,
                const double* q, int nq, double* out) {
    if (!g0 || !g1 || !v || !q || !out) return PAA_ERR_NULL;
    if (n0 < 2 || n1 < 2 || nq < 0) return PAA_ERR_SHAPE;
    for (int p = 0; p < nq; ++p) {
        double x0 = q[2 * p], x1 = q[2 * p + 1];
        if (x0 < g0[0] || x0 > g0[n0 - 1] || x1 < g1[0] || x1 > g1[n1 - 1]) { out[p] = fill; continue; }
        int i0 = int(std::lower_bound(g0, g0 + n0, x0) - g0) - 1;
        int i1 = int(std::lower_bound(g1, g1 + n1, x1) - g1) - 1;
        i0 = std::min(std::max(i0, 0), n0 - 2);
        i1 = std::min(std::max(i1, 0), n1 - 2);
        double t0 = (x0 - g0[i0]) / (g0[i0 + 1] - g0[i0]);
        double t1 = (x1 - g1[i1]) / (g1[i1 + 1] - g1[i1]);
        out[p] = v[i0 * n1 + i1] * (1 - t0) * (1 - t1) + v[i0 * n1 + i1 + 1] * (1 - t0) * t1 +
                 v[(i0 + 1) * n1 + i1] * t0 * (1 - t1) + v[(i0 + 1) * n1 + i1 + 1] * t0 * t1;
    }
    return PAA_OK;
}

// Word-level Levenshtein distance summed over the batch, and the reference word count:
// what `evaluate.load("wer")` (jiwer) divides to give the number loss_helpers.py:31 returns.
int paa_wer_counts(const char* const* refs, const char* const* hyps, int n, int64_t* errors, int64_t* ref_words) {
    if (!refs || !hyps || !errors || !ref_words) return PAA_ERR_NULL;
    if (n < 0) return PAA_ERR_SHAPE;
    auto split = [](const char* s) {
        std::vector<std::string> w;
        std::string cur;
        for (const char* c = s; c && *c; ++c) {
            if (*c == ' ' || *c == '\t' || *c == '\n' || *c == '\r' || *c == '\f' || *c == '\v') {
                if (!cur.empty()) { w.push_back(cur); cur.clear(); }
            } else {
                cur.push_back(*c);
            }
        }
        if (!cur.empty()) w.push_back(cur);
        return w;
    };
    int64_t e = 0, wsum = 0;
    for (int i = 0; i < n; ++i) {
        if (!refs[i] || !hyps[i]) return PAA_ERR_NULL;
        std::vector<std::string> a = split(refs[i]), b = split(hyps[i]);
        std::vector<int64_t> prev(b.size() + 1), cur(b.size() + 1);
        for (size_t j = 0; j <= b.size(); ++j) prev[j] = (int64_t)j;
        for (size_t r = 1; r <= a.size(); ++r) {
            cur[0] = (int64_t)r;
            for (size_t j = 1; j <= b.size(); ++j)
                cur[j] = std::min(std::min(prev[j] + 1, cur[j - 1] + 1), prev[j - 1] + (a[r - 1] != b[j - 1]));
            std::swap(prev, cur);
        }
        e += prev[b.size()];
        wsum += (int64_t)a.size();
    }
    *errors = e;
    *ref_words = wsum;
    return PAA_OK;
}

}  // extern "C"
