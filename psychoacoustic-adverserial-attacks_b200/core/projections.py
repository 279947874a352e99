"""The perceptual-constraint projections with the reference's signatures
(src/core/projections.py:11-159), executed by libpaa.so.

Differences a caller can observe: results are always fresh tensors (the reference returns the
*same object* when a constraint already holds, which needs a host synchronisation to decide; here
the branch is taken on the device and the scale factor is 1.0f), and inputs must be CUDA fp32."""
import torch

try:
    from .. import paa_lib as L
except ImportError:
    import paa_lib as L


def _rows(p):
    T = p.shape[-1]
    return p.numel() // T, T


def _empty(p):
    """An empty perturbation has nothing to project: torch returns it unchanged (clamp / `norm > eps` is False)."""
    return p.numel() == 0


def project_snr(clean, perturbation, snr_db, _step=None):
    """Rescale so that SNR(clean, perturbation) >= snr_db (projections.py:11-35; the target norm uses
    clean.numel() whatever the shape of the perturbation)."""
    L.need_cuda(clean, perturbation)
    if _empty(perturbation):                         # mean of nothing is NaN: `snr >= snr_db` is False, the norm is 0 < 1e-8
        return perturbation.detach().clone()
    p, c = L.f32c(perturbation.detach()), L.f32c(clean.detach())
    plan = L.plan_plain(p)
    rows, T = _rows(p)
    out = torch.empty_like(p)
    L.check(L.lib.paa_project_snr(plan.h, p.data_ptr(), out.data_ptr(), rows, T, c.data_ptr(), c.numel(), float(snr_db),
                                  L.step_ref(_step), plan.scratch(0, 0), L.stream_ptr(p.device)), plan.h)
    return out


def project_linf(p, min_val, max_val, _step=None):
    """clamp(p, min_val, max_val) (projections.py:37-39)."""
    L.need_cuda(p)
    if _empty(p):
        return p.detach().clone()
    x = L.f32c(p.detach())
    plan = L.plan_plain(x)
    rows, T = _rows(x)
    out = torch.empty_like(x)
    L.check(L.lib.paa_project_linf(plan.h, x.data_ptr(), out.data_ptr(), rows, T, float(min_val), float(max_val),
                                   L.step_ref(_step), L.stream_ptr(x.device)), plan.h)
    return out


def project_l2(p, epsilon, _step=None):
    """p * (epsilon/||p||) when ||p|| > epsilon, the norm taken over the whole tensor (projections.py:41-46)."""
    L.need_cuda(p)
    if _empty(p):
        return p.detach().clone()
    x = L.f32c(p.detach())
    plan = L.plan_plain(x)
    rows, T = _rows(x)
    out = torch.empty_like(x)
    L.check(L.lib.paa_project_l2(plan.h, x.data_ptr(), out.data_ptr(), rows, T, float(epsilon), L.step_ref(_step),
                                 plan.scratch(0, 0), L.stream_ptr(x.device)), plan.h)
    return out


def project_tv(p, args, clean_audio, _step=None):
    """Total variation of p at most args.tv_epsilon * TV(clean_audio) (projections.py:56-66); 2-D inputs."""
    L.need_cuda(p, clean_audio)
    if p.dim() != 2 or clean_audio.dim() != 2:
        raise IndexError("project_tv expects 2-D (batch, time) tensors")
    x, c = L.f32c(p.detach()), L.f32c(clean_audio.detach())
    plan = L.plan_plain(x)
    rows, T = x.shape
    out = torch.empty_like(x)
    L.check(L.lib.paa_project_tv(plan.h, x.data_ptr(), out.data_ptr(), rows, T, c.data_ptr(), c.shape[0], c.shape[1],
                                 float(args.tv_epsilon), L.step_ref(_step), plan.scratch(0, 0),
                                 L.stream_ptr(x.device)), plan.h)
    return out


def _spec(stft_p):
    L.need_cuda(stft_p)
    if stft_p.dtype != torch.complex64 or stft_p.dim() != 3:
        raise TypeError("expected a (batch, freq, frames) complex64 spectrum")
    return stft_p.detach()


def project_min_max_freqs(args, stft_p, min_freq, max_freq):
    """Zero every bin with min_freq <= f <= max_freq, keep the rest (projections.py:68-80)."""
    s = _spec(stft_p)
    plan = L.plan_for(s, args)
    out = torch.empty_like(s)                      # preserve_format keeps torch.stft's frame-major strides
    if out.stride() != s.stride():
        s = s.contiguous(); out = torch.empty_like(s)
    B, F, Tn = s.shape
    sb, sf, st = s.stride()
    L.check(L.lib.paa_spec_min_max_freqs(plan.h, s.data_ptr(), out.data_ptr(), B, Tn, sb, sf, st, float(min_freq),
                                         float(max_freq), L.stream_ptr(s.device)), plan.h)
    return out


def compute_fm_weighted_norm_interp(stft_p, interp, args):
    """sqrt(sum |X|^2 * w(10 log10(|X|^2+1e-10), f)) with the bilinear ISO-226 penalty grid (projections.py:83-113).
    Returns a 0-dim tensor on the device; the bilinear lookup runs on the GPU, not in scipy."""
    s = _spec(stft_p)
    plan = L.plan_for(s, args)
    plan.set_fm_grid(interp)
    B, F, Tn = s.shape
    sb, sf, st = s.stride()
    scratch = plan.scratch(0, 0)
    L.check(L.lib.paa_spec_fm_norm(plan.h, s.data_ptr(), B, Tn, sb, sf, st, scratch, L.stream_ptr(s.device)), plan.h)
    return plan._scratch[:16].view(torch.float32)[L.S_NORM].clone()


def project_fm_norm(stft_p, args, interp):
    """Scale the spectrum so that its Fletcher-Munson weighted norm is at most args.fm_epsilon (projections.py:116-133)."""
    s = _spec(stft_p)
    plan = L.plan_for(s, args)
    plan.set_fm_grid(interp)
    out = torch.empty_like(s)
    if out.stride() != s.stride():
        s = s.contiguous(); out = torch.empty_like(s)
    B, F, Tn = s.shape
    sb, sf, st = s.stride()
    scratch = plan.scratch(0, 0)
    L.check(L.lib.paa_spec_fm_project(plan.h, s.data_ptr(), out.data_ptr(), B, Tn, sb, sf, st, float(args.fm_epsilon),
                                      scratch, L.stream_ptr(s.device)), plan.h)
    return out


def project_phon_level(stft_p, args, spl_thresh, plot_debug=False, tag=""):
    """Clip every bin's level 20 log10(|X|+1e-8) to spl_thresh - max(spl_thresh) + args.phon_reference_db and
    rebuild the bin with its phase (projections.py:138-159).  plot_debug is accepted and ignored (no matplotlib)."""
    s = _spec(stft_p)
    L.need_cuda(spl_thresh)
    plan = L.plan_for(s, args)
    thr = L.f32c(spl_thresh.detach().reshape(-1))
    if thr.numel() != plan.n_fft // 2 + 1:
        raise RuntimeError(f"spl_thresh has {thr.numel()} bins, expected {plan.n_fft // 2 + 1}")
    out = torch.empty_like(s)
    if out.stride() != s.stride():
        s = s.contiguous(); out = torch.empty_like(s)
    B, F, Tn = s.shape
    sb, sf, st = s.stride()
    L.check(L.lib.paa_spec_phon_level(plan.h, s.data_ptr(), out.data_ptr(), B, Tn, sb, sf, st, thr.data_ptr(),
                                      float(args.phon_reference_db), L.stream_ptr(s.device)), plan.h)
    return out
