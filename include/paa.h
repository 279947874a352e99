/* paa.h -- C ABI of libpaa.so: the B200 (sm_100a) perturbation hot path.
 *
 * One shared library, plain pointers and sizes, no torch types.  It replaces, for the path
 * BASELINE.json names, the torch code of the reference (paths relative to the reference repo):
 *
 *   PGD / Adam step on the perturbation ........ src/training_utils/train.py:155-175
 *   norm_type dispatch .......................... src/training_utils/train.py:69-99, :38-66
 *   l2 / linf / snr / tv ........................ src/core/projections.py:11-66
 *   min_max_freqs / fletcher_munson / max_phon .. src/core/projections.py:68-159
 *   STFT / ISTFT ................................ src/core/fourier_transforms.py:4-41
 *   ISO-226 tables .............................. src/core/iso.py:34-266, src/training_utils/build.py:325-348
 *
 * Conventions
 *   - every entry point returns a paa_status code, 0 = ok; nothing throws, exits or prints.
 *   - device pointers are BORROWED; the library allocates nothing per call.  Scratch is
 *     caller-owned (size from paa_scratch_bytes) and needs no initialisation.
 *   - all device work is enqueued on the caller's stream (a cudaStream_t passed as void*);
 *     no entry point synchronises with the host, so data-dependent branches of the reference
 *     ("if norm > eps") are taken on the device and the scale lands in scratch (paa_scalars).
 *   - tensors are fp32, row-major contiguous [rows, T]; spectra are interleaved complex64
 *     addressed by element strides (batch, freq, frame), so torch.stft's layout is native.
 *   - a handle holds only immutable device tables for one (device, n_fft, hop, sr); calls on
 *     one handle are re-entrant as long as each concurrent call has its own scratch.
 */
#ifndef PAA_H_
#define PAA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum paa_status {
    PAA_OK = 0,
    PAA_ERR_NULL = 1,         /* a required pointer is NULL */
    PAA_ERR_SHAPE = 2,        /* rows/T/out_len/strides out of range (e.g. T <= n_fft/2: reflect pad impossible) */
    PAA_ERR_UNSUPPORTED = 3,  /* n_fft not in {512,1024}, hop not dividing n_fft, unknown step mode ... */
    PAA_ERR_CUDA = 4,         /* a CUDA call failed; see paa_last_cuda_error */
    PAA_ERR_RANGE = 5,        /* phon outside [0,90] or frequency outside [20,20000]  (iso.py:97-98,:152-153) */
    PAA_ERR_NEED_CLEAN = 6,   /* snr / tv without clean audio                          (train.py:90-95) */
    PAA_ERR_ALIAS = 7,        /* p_out aliases p_in where the entry point forbids it */
    PAA_ERR_NOLA = 8,         /* window overlap-add envelope < 1e-11 (torch.istft's RuntimeError) */
    PAA_ERR_STATE = 9         /* e.g. fletcher_munson before paa_set_fm_grid */
} paa_status;

typedef struct paa_handle paa_handle;

/* ---- optimiser step fused in front of the projection (train.py:161 / torch.optim.Adam) ---- */
enum { PAA_STEP_NONE = 0, PAA_STEP_PGD = 1, PAA_STEP_ADAM = 2 };

/* Universal-perturbation data parallelism (SURVEY.md 8e "mode U", next-row N4): G ranks attack disjoint utterance
 * shards with ONE shared (1,T) perturbation.  The reference's step then needs the SUM over ranks of dL/dp
 * (train.py:158 on a batch that is the union of the shards; the CTC reduction is "sum") and, for snr / tv, the clean
 * statistics of the whole batch (projections.py:17, :58).  Instead of an all-reduce followed by the step, every rank
 * hands the kernels the G partial buffers themselves -- peer-mapped device memory over NVLink (symmetric memory) --
 * and the kernels add them in index order while they step: the collective is fused into the hot-path pass, every rank
 * computes bit-identical results, and nothing but the partials crosses the links.  Making the partials visible
 * (a device-side barrier on the caller's stream) is the caller's job. */
#define PAA_MAX_PARTS 8
typedef struct paa_parts {
    int           n;                            /* ranks, 1..PAA_MAX_PARTS                                        */
    const float*  grad[PAA_MAX_PARTS];          /* rank r's partial dL/dp [rows, T]; summed left to right in fp32  */
    const double* clean_stats[PAA_MAX_PARTS];   /* rank r's {sum clean^2, sum |clean[t+1]-clean[t]|, numel(clean)}
                                                   (paa_clean_stats); NULL entries = statistics come from `clean`  */
    int64_t       clean_numel;                  /* numel of the whole clean batch (all ranks); 0 = the kernels sum
                                                   the parts' third statistic instead (uneven / changing shards:
                                                   no host-side agreement on the batch size is needed)             */
} paa_parts;

typedef struct paa_step {
    int          mode;      /* PAA_STEP_*                                                         */
    const float* grad;      /* [rows, T] dL/dp exactly as autograd left it in p.grad               */
    double       lr;        /* PGD: p += fp32(lr)*sign(grad).  Adam: learning rate of this step    */
    float*       adam_m;    /* [rows, T] exp_avg,    updated in place                              */
    float*       adam_v;    /* [rows, T] exp_avg_sq, updated in place                              */
    int64_t      adam_t;    /* 1-based step count of THIS update (state['step'] after increment)   */
    double       beta1, beta2, eps;   /* torch defaults 0.9, 0.999, 1e-8 (build.py:357)            */
    const paa_parts* parts; /* NULL, or mode U: parts->grad replaces `grad` (when non-NULL) and parts->clean_stats
                               replaces the clean-audio reductions of snr / tv; honoured with mode NONE too */
} paa_step;

/* Layout of the float scalars every reducing projection leaves at the start of scratch. */
enum {
    PAA_S_SCALE = 0,   /* factor applied to p (1.0f when the constraint already held)             */
    PAA_S_NORM  = 1,   /* l2: ||p||, snr: ||p||, tv: TV(p), fletcher_munson: weighted norm        */
    PAA_S_AUX0  = 2,   /* snr: mean(clean^2); tv: TV(clean)                                       */
    PAA_S_AUX1  = 3,   /* snr: current SNR in dB; tv: epsilon = tv_epsilon*TV(clean)              */
    PAA_S_COUNT = 8
};

const char* paa_status_string(int status);
int         paa_version(void);
/* Number of CUDA kernels this library has launched in this process (all handles, all streams). */
int64_t     paa_launch_count(void);

/* ---- handle ------------------------------------------------------------------------------ */
int paa_create(int device, int n_fft, int hop, int sr, paa_handle** out);
int paa_destroy(paa_handle* h);
int paa_last_cuda_error(const paa_handle* h);        /* cudaError_t of the last PAA_ERR_CUDA    */
int paa_num_bins(const paa_handle* h);                /* F = n_fft/2+1                            */
int paa_num_frames(const paa_handle* h, int T);       /* T' = 1 + T/hop (centre=True)             */
/* Scratch a call needs.  rows = T = 0: the reducing time-domain projections (l2, snr, tv) and the paa_spec_fm_*
 * functions (scalars + block partials, ~128 KB).  rows, T > 0: the STFT-domain projections, which add two [rows, T]
 * fp32 buffers (the staging buffer of an Adam pre-pass / of fletcher_munson's two passes, and in mode U the gradient
 * summed over the parts) and fletcher_munson's per-tile partials. */
size_t paa_scratch_bytes(const paa_handle* h, int rows, int T);
/* Copies the PAA_S_* scalars of the last reducing call on `scratch` to the host (synchronises `stream`). */
int paa_scalars(const paa_handle* h, const void* scratch, float* out8, void* stream);

/* ---- host-side ISO-226 tables (no GPU needed) -- iso.py, build.py:325-348 ------------------ */
int paa_iso226_spl(double phon, const double* freqs_hz, int n, double* out_spl);
int paa_weight_grid(double* phon_knots10, double* freq_knots30, double* weights_10x30);
int paa_spl_thresh(int n_fft, int sr, double max_phon_level, float* out_F);
/* bilinear RegularGridInterpolator(bounds_error=False, fill_value) on the host, for known-answer tests */
int paa_interp2(const double* g0, int n0, const double* g1, int n1, const double* values, double fill,
                const double* query_xy, int nq, double* out);
/* Install the (phon x freq) penalty grid used by fletcher_munson: interp.grid / interp.values /
 * interp.fill_value of the object iso.py:238-266 returns.  Synchronous, rarely called. */
int paa_set_fm_grid(paa_handle* h, const double* phon_knots, int n_phon, const double* freq_knots, int n_freq,
                    const double* values, double fill_value);

/* ---- step + projection, time domain ----------------------------------------------------------
 * p_out may equal p_in (in place) except for tv with a fused PGD step. step may be NULL. */
int paa_project_linf(paa_handle* h, const float* p_in, float* p_out, int rows, int T,
                     double lo, double hi, const paa_step* step, void* stream);                /* projections.py:37-39 */
int paa_project_l2(paa_handle* h, const float* p_in, float* p_out, int rows, int T,
                   double epsilon, const paa_step* step, void* scratch, void* stream);         /* projections.py:41-46 */
int paa_project_snr(paa_handle* h, const float* p_in, float* p_out, int rows, int T,
                    const float* clean, int64_t clean_numel, double snr_db,
                    const paa_step* step, void* scratch, void* stream);                        /* projections.py:11-35 */
int paa_project_tv(paa_handle* h, const float* p_in, float* p_out, int rows, int T,
                   const float* clean, int clean_rows, int clean_T, double tv_epsilon,
                   const paa_step* step, void* scratch, void* stream);                         /* projections.py:56-66 */

/* Mode U helper: out3[0] = sum clean^2, out3[1] = sum over rows of sum_t |clean[r,t+1]-clean[r,t]|, out3[2] = rows*T of
 * this rank's utterances (fp64 on the device, fixed summation order); goes into paa_parts.clean_stats.  Uses the
 * partials area of `scratch`. */
int paa_clean_stats(paa_handle* h, const float* clean, int rows, int T, double* out3, void* scratch, void* stream);

/* ---- step + projection, STFT domain (train.py:38-66): STFT -> per-bin op -> ISTFT in ONE kernel,
 * the spectrum never reaches HBM.  p_out must NOT alias p_in.  p_out is [rows, out_len]; samples
 * past hop*(T'-1) are zero (train.py:27-35 with clean_audio given: out_len = clean.shape[-1];
 * with clean_audio=None pass out_len = hop*(T'-1)).
 * A PGD step is applied while the input span is staged; an Adam step runs first as a streaming pass into
 * scratch (halo samples are recomputed by neighbouring tiles, optimiser state must be updated once).
 * max_phon: spl_thresh_F is the device array build.py:325-348 makes (F floats); the fused kernel clips in the
 * linear domain, |X'| = min(|X| + 1e-8, 10^(thr/20)), which equals the reference's dB round trip up to its own
 * fp32 rounding (~1e-6 relative); paa_spec_phon_level below performs the literal dB round trip.
 * fletcher_munson needs paa_set_fm_grid first (PAA_ERR_STATE otherwise) and is two passes: the weighted norm (one
 * fp64 partial per tile in scratch -- sized by paa_scratch_bytes(rows, T), no limit on rows x tiles), then
 * exact_roundtrip == 0 (what the Python drop-in passes by default): scale * q on the span the inverse transform
 *   reconstructs, zeros behind it -- ISTFT(s * STFT(q)) = s * q there, 8 B/sample, and the kernel finalizes the norm
 *   itself (two launches in all);
 * exact_roundtrip != 0: the literal ISTFT(scale * STFT(q)) of the reference (finalize launch + a second transform;
 *   differs from the identity form only by the round trip's own fp32 rounding, ~5e-7 relative). */
int paa_project_min_max_freqs(paa_handle* h, const float* p_in, float* p_out, int rows, int T, int out_len,
                              double min_freq, double max_freq,
                              const paa_step* step, void* scratch, void* stream);              /* projections.py:68-80 */
int paa_project_max_phon(paa_handle* h, const float* p_in, float* p_out, int rows, int T, int out_len,
                         const float* spl_thresh_F, double phon_reference_db,
                         const paa_step* step, void* scratch, void* stream);                   /* projections.py:138-159 */
int paa_project_fletcher_munson(paa_handle* h, const float* p_in, float* p_out, int rows, int T, int out_len,
                                double fm_epsilon, int exact_roundtrip,
                                const paa_step* step, void* scratch, void* stream);            /* projections.py:83-133 */

/* ---- the optimiser step on its own (no projection) ------------------------------------------ */
int paa_step_only(paa_handle* h, const float* p_in, float* p_out, int rows, int T,
                  const paa_step* step, void* stream);

/* ---- un-fused public functions of the reference, kept so they stay drop-in ------------------- */
/* compute_stft (fourier_transforms.py:4-29): spec[b,f,t] at spec + 2*(b*sb + f*sf + t*st) floats */
int paa_stft(paa_handle* h, const float* x, int rows, int T,
             float* spec, int64_t sb, int64_t sf, int64_t st, void* stream);
/* compute_istft (fourier_transforms.py:31-41): y is [rows, hop*(n_frames-1)] */
int paa_istft(paa_handle* h, const float* spec, int64_t sb, int64_t sf, int64_t st,
              int rows, int n_frames, float* y, void* stream);
int paa_spec_min_max_freqs(paa_handle* h, const float* spec_in, float* spec_out, int rows, int n_frames,
                           int64_t sb, int64_t sf, int64_t st, double min_freq, double max_freq, void* stream);
int paa_spec_phon_level(paa_handle* h, const float* spec_in, float* spec_out, int rows, int n_frames,
                        int64_t sb, int64_t sf, int64_t st, const float* spl_thresh_F, double phon_reference_db,
                        void* stream);
/* compute_fm_weighted_norm_interp: the norm lands in scratch[PAA_S_NORM]; project_fm_norm also scales. */
int paa_spec_fm_norm(paa_handle* h, const float* spec_in, int rows, int n_frames,
                     int64_t sb, int64_t sf, int64_t st, void* scratch, void* stream);
int paa_spec_fm_project(paa_handle* h, const float* spec_in, float* spec_out, int rows, int n_frames,
                        int64_t sb, int64_t sf, int64_t st, double fm_epsilon, void* scratch, void* stream);

/* ---- the input side of the path: x_adv = clamp(clean + p, -1, 1), train.py:136 - */
int paa_compose_clamp(paa_handle* h, const float* clean, int clean_rows, const float* p, int p_rows, int T,
                      float* x_adv, void* stream);
/* its backward: grad_p[b or 0, t] = (sum over b of) grad_x_adv[b,t] * 1[-1 <= clean[b,t]+p[.,t] <= 1]  (torch's clamp mask);
 * for a universal (1,T) p the batch is summed in a fixed order (even rows in row order + odd rows in row order: the two
 * CTAs of a thread-block cluster each take one half and meet through distributed shared memory), so the result is
 * bit-identical from run to run and within rounding of torch's own batch sum. */
int paa_compose_clamp_backward(paa_handle* h, const float* clean, int clean_rows, const float* p, int p_rows, int T,
                               const float* grad_x_adv, float* grad_p, void* stream);

/* ---- WER counters (loss_helpers.py:25-32; evaluate/jiwer semantics): sum(S+D+I), sum(ref words) */
int paa_wer_counts(const char* const* references, const char* const* hypotheses, int n,
                   int64_t* errors, int64_t* ref_words);

#ifdef __cplusplus
}
#endif
#endif /* PAA_H_ */
