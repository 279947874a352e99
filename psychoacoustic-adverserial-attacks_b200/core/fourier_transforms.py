"""compute_stft / compute_istft with the reference's signatures (src/core/fourier_transforms.py:4-41),
computed by the warp-per-frame shared-memory FFT of libpaa.so instead of torch.stft / cuFFT."""
import torch

try:
    from .. import paa_lib as L
except ImportError:          # the package directory itself is on sys.path (drop-in for the reference's src/)
    import paa_lib as L


def compute_stft(p, args):
    """(batch, length) fp32 -> (batch, n_fft//2+1, 1+length//hop) complex64.
    Periodic Hann(n_fft), centre=True with reflect padding, one-sided, unnormalised; the result has
    torch.stft's memory layout (frame-major, strides (F*T', 1, F))."""
    L.need_cuda(p)
    squeeze = p.dim() == 1
    x = L.f32c(p.detach().reshape(-1, p.shape[-1]))
    plan = L.plan_for(x, args)
    rows, T = x.shape
    F, frames = plan.n_fft // 2 + 1, 1 + T // plan.hop
    buf = torch.empty((rows, frames, F), dtype=torch.complex64, device=x.device)
    L.check(L.lib.paa_stft(plan.h, x.data_ptr(), rows, T, buf.data_ptr(), frames * F, 1, F, L.stream_ptr(x.device)), plan.h)
    spec = buf.transpose(1, 2)
    return spec[0] if squeeze else spec


def compute_istft(stft_p, args):
    """(batch, F, T') complex64 -> (batch, hop*(T'-1)) fp32: irfft, Hann, overlap-add, divide by the
    window envelope, trim n_fft/2 from both ends (torch.istft with centre=True)."""
    L.need_cuda(stft_p)
    if stft_p.dtype != torch.complex64:
        raise TypeError(f"expected complex64, got {stft_p.dtype}")
    squeeze = stft_p.dim() == 2
    s = stft_p.detach()
    if squeeze:
        s = s.unsqueeze(0)
    plan = L.plan_for(s, args)
    rows, F, frames = s.shape
    if F != plan.n_fft // 2 + 1:
        raise RuntimeError(f"expected {plan.n_fft // 2 + 1} frequency bins, got {F}")
    y = torch.empty((rows, plan.hop * (frames - 1)), dtype=torch.float32, device=s.device)
    sb, sf, st = s.stride()
    L.check(L.lib.paa_istft(plan.h, s.data_ptr(), sb, sf, st, rows, frames, y.data_ptr(), L.stream_ptr(s.device)), plan.h)
    return y[0] if squeeze else y
